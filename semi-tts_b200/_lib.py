"""ctypes binding of libvqb200.so (include/vqb.h).  Loading fails loudly: there is no fallback."""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvqb200.so")
ABI_VERSION = 3

# flags (include/vqb.h)
SCORE_L2 = 0x0001
SCORE_LINEAR = 0x0002
STOP_GRAD = 0x0004
SKIP = 0x0008
TEMP_GRAD = 0x0010
TENSOR_CORES = 0x0020
SEARCH_TENSOR = TENSOR_CORES
AFTER_ASSEMBLE = 0x0040

_p = ctypes.c_void_p


class FwdArgs(ctypes.Structure):
    _fields_ = [("struct_size", ctypes.c_uint32), ("flags", ctypes.c_uint32),
                ("n_rows", ctypes.c_int64), ("dim", ctypes.c_int64), ("n_codes", ctypes.c_int64),
                ("x", _p), ("score_w", _p), ("score_b", _p), ("score_w_bf16", _p), ("gather_table", _p),
                ("temp", _p), ("p_code", _p), ("idx", _p), ("new_latent", _p), ("hist", _p),
                ("sq_err_sum", _p), ("search_stats", _p), ("operand_cache", _p), ("workspace", _p), ("workspace_bytes", ctypes.c_size_t),
                ("row_lengths", _p), ("frames_per_utt", ctypes.c_int64), ("ctc_logp", _p), ("ctc_eps", ctypes.c_float)]


class BwdTail(ctypes.Structure):
    _fields_ = [("phn_attr", _p), ("n_attr", ctypes.c_int64), ("dim_attr", ctypes.c_int64), ("d_flat", _p),
                ("counter", _p), ("world", ctypes.c_int32), ("rank", ctypes.c_int32), ("peer_bufs", _p),
                ("timeout_ms", ctypes.c_uint32), ("reserved", ctypes.c_uint32)]


class BwdArgs(ctypes.Structure):
    _fields_ = [("struct_size", ctypes.c_uint32), ("flags", ctypes.c_uint32),
                ("n_rows", ctypes.c_int64), ("dim", ctypes.c_int64), ("n_codes", ctypes.c_int64),
                ("n_real_rows", ctypes.c_int64),
                ("x", _p), ("score_w", _p), ("score_b", _p), ("gather_table", _p), ("temp", _p),
                ("p_code", _p), ("idx", _p), ("g_p", _p), ("g_q", _p),
                ("dx", _p), ("d_score_w", _p), ("colsum", _p), ("d_gather", _p), ("d_temp", _p),
                ("operand_cache", _p), ("workspace", _p), ("workspace_bytes", ctypes.c_size_t),
                ("row_lengths", _p), ("frames_per_utt", ctypes.c_int64), ("g_logp", _p), ("ctc_eps", ctypes.c_float),
                ("tail", ctypes.POINTER(BwdTail))]


EXPORTS = ["vqb_abi_version", "vqb_last_error", "vqb_device_count", "vqb_operand_cache_bytes", "vqb_assemble_table",
           "vqb_table_backward", "vqb_forward_workspace", "vqb_forward", "vqb_backward_workspace",
           "vqb_backward", "vqb_forward_kernel_name", "vqb_backward_kernel_name", "vqb_launch_count", "vqb_exchange_bytes", "vqb_exchange_finish", "vqb_inference_gather", "vqb_scatter_add", "vqb_scatter_workspace", "vqb_loss_backward",
           "vqb_row_argmax", "vqb_segment_plan", "vqb_segment_mean", "vqb_segment_mean_backward", "vqb_ctc_logp", "vqb_ctc_logp_backward"]

_lib = None
_lock = threading.Lock()


def load():
    """Returns the loaded library; raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                "semi-tts_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C semi-tts_b200/csrc` (there is no CPU / PyTorch fallback)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        lib.vqb_abi_version.restype = ctypes.c_int
        lib.vqb_last_error.restype = ctypes.c_char_p
        lib.vqb_device_count.restype = ctypes.c_int
        i64, sz = ctypes.c_int64, ctypes.c_size_t
        lib.vqb_assemble_table.argtypes = [_p, _p, _p, _p, i64, i64, i64, i64, _p, _p, _p, _p, _p]
        lib.vqb_operand_cache_bytes.argtypes = [i64, i64]
        lib.vqb_table_backward.argtypes = [_p, _p, _p, _p, i64, i64, i64, i64, _p, _p, _p, _p]
        lib.vqb_forward_workspace.argtypes = [ctypes.POINTER(FwdArgs), ctypes.POINTER(sz)]
        lib.vqb_forward.argtypes = [ctypes.POINTER(FwdArgs), _p]
        lib.vqb_backward_workspace.argtypes = [ctypes.POINTER(BwdArgs), ctypes.POINTER(sz)]
        lib.vqb_backward.argtypes = [ctypes.POINTER(BwdArgs), _p]
        lib.vqb_inference_gather.argtypes = [_p, i64, _p, i64, i64, _p, _p]
        lib.vqb_scatter_add.argtypes = [_p, i64, _p, i64, i64, _p, _p, _p, sz, _p]
        lib.vqb_scatter_workspace.argtypes = [i64, i64, i64, ctypes.POINTER(sz)]
        lib.vqb_loss_backward.argtypes = [_p, _p, _p, i64, i64, i64, _p, _p, _p, ctypes.c_int, _p, _p]
        lib.vqb_row_argmax.argtypes = [_p, i64, i64, _p, _p]
        lib.vqb_segment_plan.argtypes = [_p, i64, i64, i64, _p, _p, _p, _p, _p]
        lib.vqb_segment_mean.argtypes = [_p, _p, _p, _p, i64, i64, i64, i64, _p, _p]
        lib.vqb_segment_mean_backward.argtypes = [_p, _p, _p, i64, i64, i64, i64, _p, _p]
        lib.vqb_ctc_logp.argtypes = [_p, i64, i64, i64, ctypes.c_float, _p, _p]
        lib.vqb_ctc_logp_backward.argtypes = [_p, _p, i64, i64, i64, ctypes.c_float, _p, ctypes.c_int, _p]
        lib.vqb_forward_kernel_name.argtypes = [ctypes.POINTER(FwdArgs)]
        lib.vqb_backward_kernel_name.argtypes = [ctypes.POINTER(BwdArgs)]
        for name in EXPORTS:
            getattr(lib, name).restype = ctypes.c_int
        for name in ("vqb_last_error", "vqb_forward_kernel_name", "vqb_backward_kernel_name"):
            getattr(lib, name).restype = ctypes.c_char_p
        lib.vqb_operand_cache_bytes.restype = ctypes.c_size_t
        lib.vqb_launch_count.restype = ctypes.c_uint64
        lib.vqb_exchange_bytes.argtypes = [i64, ctypes.c_int32]
        lib.vqb_exchange_finish.argtypes = [ctypes.POINTER(BwdTail), i64, _p]
        lib.vqb_exchange_bytes.restype = ctypes.c_size_t
        if lib.vqb_abi_version() != ABI_VERSION:
            raise RuntimeError("semi-tts_b200: libvqb200.so ABI %d != expected %d -- rebuild"
                               % (lib.vqb_abi_version(), ABI_VERSION))
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        msg = load().vqb_last_error()
        raise RuntimeError((msg.decode() if msg else "libvqb200 error") + " [code %d]" % rc)


def ptr(t):
    """device pointer of a tensor (None -> NULL)"""
    return None if t is None else t.data_ptr()
