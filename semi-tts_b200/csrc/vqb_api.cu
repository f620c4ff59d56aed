// C-ABI entry points of libvqb200.so (declared in include/vqb.h): argument validation, kernel
// selection, thread-local error string.  No CPU fallback exists anywhere in this library.
#include <atomic>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "vqb_common.cuh"

namespace vqb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int invalid(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return VQB_ERR_INVALID;
}

int cuda_fail(cudaError_t e, const char* what) {
    // keep torch's wording for OOM so the reference's substring check still matches
    // (bin/train_vqvae.py:321 looks for 'out of memory')
    if (e == cudaErrorMemoryAllocation)
        set_error("CUDA out of memory in libvqb200 (%s)", what);
    else
        set_error("CUDA error in libvqb200: %s (%s) at %s", cudaGetErrorName(e), cudaGetErrorString(e), what);
    cudaGetLastError();   // clear the sticky-free error state
    return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? VQB_ERR_NO_DEVICE : VQB_ERR_CUDA;
}

struct DevInfo { int sms; int smem; int major; };
static DevInfo dev_info() {
    static thread_local int cached_dev = -1;
    static thread_local DevInfo info = {148, 227 * 1024, 10};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return info;
    if (dev != cached_dev) {
        cudaDeviceGetAttribute(&info.sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&info.smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cudaDeviceGetAttribute(&info.major, cudaDevAttrComputeCapabilityMajor, dev);
        cached_dev = dev;
    }
    return info;
}
int sm_count() { return dev_info().sms; }
int max_optin_smem() { return dev_info().smem; }

static int require_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_error("libvqb200: no CUDA device visible (this library has no CPU fallback)");
        return VQB_ERR_NO_DEVICE;
    }
    if (dev_info().major != 10) {
        set_error("libvqb200: built for sm_100a only; current device has compute capability %d.x", dev_info().major);
        return VQB_ERR_NO_DEVICE;
    }
    return VQB_OK;
}

static int check_common(const char* fn, int64_t N, int64_t D, int64_t K) {
    if (N < 0 || D <= 0 || K <= 0) return invalid("%s: bad shape N=%lld D=%lld K=%lld", fn, (long long)N, (long long)D, (long long)K);
    if (N >= (1ll << 31) || K >= (1ll << 24)) return invalid("%s: N must be < 2^31 and K < 2^24", fn);
    if (D % 4 != 0 || D > 512) return invalid("%s: D must be a multiple of 4 and <= 512 (got %lld)", fn, (long long)D);
    return VQB_OK;
}

}  // namespace vqb

using namespace vqb;

namespace vqb { void set_debug_timeline(void* p); }
// undocumented developer hook: device buffer of 128 u64 that CTA 0 of the tensor-core forward fills with
// (tag << 56 | globaltimer ns) marks; pass NULL to disable
extern "C" __attribute__((visibility("default"))) void vqb_debug_set_timeline(void* dev_ptr) { vqb::set_debug_timeline(dev_ptr); }

// undocumented developer hook (A/B): force the software-pipelined x_lo of the streamed 3xTF32 search on (1) / off (0);
// -1 = default (on unless VQB_SEARCH_NOPIPE is set)
namespace vqb { void set_debug_search_pipe(int v); }
extern "C" __attribute__((visibility("default"))) void vqb_debug_set_search_pipe(int v) { vqb::set_debug_search_pipe(v); }

namespace vqb {
static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}
namespace vqb {
// measurement hook: a pair of CUDA events recorded on the launch stream immediately before / after the DOMINANT kernel
// of the next forward / backward / scatter call (not its helper kernels), so that a caller can time that kernel alone
static cudaEvent_t g_kev0 = nullptr, g_kev1 = nullptr;
void kernel_event_begin(cudaStream_t s) { if (g_kev0) cudaEventRecord(g_kev0, s); }
void kernel_event_end(cudaStream_t s) { if (g_kev1) cudaEventRecord(g_kev1, s); }
}
extern "C" __attribute__((visibility("default"))) void vqb_debug_set_kernel_events(void* ev_start, void* ev_stop) {
    vqb::g_kev0 = (cudaEvent_t)ev_start; vqb::g_kev1 = (cudaEvent_t)ev_stop;
}

extern "C" uint64_t vqb_launch_count(void) { return vqb::g_launches.load(std::memory_order_relaxed); }

extern "C" int vqb_abi_version(void) { return VQB_ABI_VERSION; }
extern "C" const char* vqb_last_error(void) { return g_err; }

extern "C" int vqb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int d = 0; d < n; ++d) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ++ok;
    }
    return ok;
}

static int validate_fwd(const vqb_fwd_args* a) {
    if (!a || a->struct_size != sizeof(vqb_fwd_args)) return invalid("vqb_forward: struct_size mismatch (ABI %d)", VQB_ABI_VERSION);
    const bool l2 = a->flags & VQB_SCORE_L2, lin = a->flags & VQB_SCORE_LINEAR;
    if (l2 == lin) return invalid("vqb_forward: exactly one of VQB_SCORE_L2 / VQB_SCORE_LINEAR must be set");
    int rc = check_common("vqb_forward", a->n_rows, a->dim, a->n_codes);
    if (rc) return rc;
    if (a->n_rows == 0) return VQB_OK;
    if (!a->x || !a->score_w || !a->score_b || !a->gather_table || !a->idx || !a->new_latent)
        return invalid("vqb_forward: x, score_w, score_b, gather_table, idx and new_latent are required");
    if (l2 && !a->temp) return invalid("vqb_forward: temp is required for the L2 score");
    if (!aligned16(a->x) || !aligned16(a->score_w) || !aligned16(a->gather_table) || !aligned16(a->new_latent) ||
        (a->p_code && !aligned16(a->p_code)))
        return invalid("vqb_forward: tensor pointers must be 16-byte aligned");
    if (a->row_lengths || a->ctc_logp) {
        if (a->frames_per_utt <= 0 || a->n_rows % a->frames_per_utt != 0)
            return invalid("vqb_forward: row_lengths / ctc_logp need frames_per_utt > 0 dividing n_rows");
        if (!((a->flags & VQB_TENSOR_CORES) && forward_pcode_supported(a)))
            return invalid("vqb_forward: row_lengths / ctc_logp are served by the parity-mode tensor-core kernel only (p_code, K <= 64, D in {32, 64})");
    }
    return VQB_OK;
}

extern "C" int vqb_forward_workspace(const vqb_fwd_args* a, size_t* bytes) {
    if (!bytes) return invalid("vqb_forward_workspace: bytes is NULL");
    *bytes = 0;
    int rc = validate_fwd(a);
    if (rc) return rc;
    if ((a->flags & VQB_TENSOR_CORES) && a->n_rows > 0) {
        if (forward_pcode_supported(a)) { *bytes = forward_pcode_workspace(a); return VQB_OK; }
        return forward_tensor_workspace(a, bytes);
    }
    return VQB_OK;
}

extern "C" int vqb_forward(const vqb_fwd_args* a, void* stream) {
    int rc = validate_fwd(a);
    if (rc) return rc;
    if ((rc = require_device())) return rc;
    if (a->n_rows == 0) return VQB_OK;
    if ((a->flags & VQB_TENSOR_CORES) && forward_pcode_supported(a)) return launch_forward_pcode(a, (cudaStream_t)stream);
    if ((a->flags & VQB_TENSOR_CORES) && forward_tensor_supported(a)) return launch_forward_tensor(a, (cudaStream_t)stream);
    if (a->dim > 256) return invalid("vqb_forward: the exact-fp32 path supports D <= 256 (got %lld)", (long long)a->dim);
    return launch_forward_simt(a, (cudaStream_t)stream);
}

extern "C" const char* vqb_forward_kernel_name(const vqb_fwd_args* a) {
    if (validate_fwd(a) != VQB_OK) return "invalid";
    if ((a->flags & VQB_TENSOR_CORES) && forward_pcode_supported(a)) return "vqb_fwd_pcode_kernel";
    if ((a->flags & VQB_TENSOR_CORES) && forward_tensor_supported(a)) return "vqb_fwd_tc_kernel";
    return a->n_codes <= 64 ? "vqb_fwd_simt_small_kernel" : "vqb_fwd_simt_generic_kernel";
}

static int validate_bwd(const vqb_bwd_args* a) {
    if (!a || a->struct_size != sizeof(vqb_bwd_args)) return invalid("vqb_backward: struct_size mismatch (ABI %d)", VQB_ABI_VERSION);
    const bool l2 = a->flags & VQB_SCORE_L2, lin = a->flags & VQB_SCORE_LINEAR;
    if (l2 == lin) return invalid("vqb_backward: exactly one of VQB_SCORE_L2 / VQB_SCORE_LINEAR must be set");
    int rc = check_common("vqb_backward", a->n_rows, a->dim, a->n_codes);
    if (rc) return rc;
    if (a->n_rows == 0) return VQB_OK;
    if (!a->idx) return invalid("vqb_backward: idx is required");
    if (!a->g_p && !a->g_q && !a->g_logp) return invalid("vqb_backward: at least one of g_p / g_q / g_logp is required");
    if (a->g_p && a->g_logp) return invalid("vqb_backward: give either g_p or g_logp (add the two upstream with vqb_ctc_logp_backward)");
    if (a->row_lengths || a->g_logp) {
        if (a->frames_per_utt <= 0 || a->n_rows % a->frames_per_utt != 0)
            return invalid("vqb_backward: row_lengths / g_logp need frames_per_utt > 0 dividing n_rows");
        if (!backward_pcode_supported(a))
            return invalid("vqb_backward: row_lengths / g_logp are served by the vqb_bwd_pcode_kernel route only");
    }
    return VQB_OK;
}

extern "C" int vqb_backward_workspace(const vqb_bwd_args* a, size_t* bytes) {
    if (!bytes) return invalid("vqb_backward_workspace: bytes is NULL");
    *bytes = 0;
    int rc = validate_bwd(a);
    if (rc) return rc;
    if (a->n_rows > 0) {
        *bytes = backward_pcode_workspace(a);
        if (!a->g_p && !a->g_logp && (a->flags & VQB_STOP_GRAD)) *bytes = scatter_workspace_bytes(a->n_rows, a->n_codes, a->dim);
        else if (backward_generic_needed(a)) *bytes = backward_generic_workspace(a);
    }
    return VQB_OK;
}

extern "C" int vqb_backward(const vqb_bwd_args* a, void* stream) {
    int rc = validate_bwd(a);
    if (rc) return rc;
    if ((rc = require_device())) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const bool l2 = a->flags & VQB_SCORE_L2;
    if (a->n_rows == 0) {
        // an empty shard of a data-parallel run: the fused tail still runs (over zero partial records), so that this rank
        // pushes its zeros and its peers' exchange completes
        if (a->tail && l2) {
            if (!a->tail->d_flat || !a->tail->counter || !a->gather_table) return invalid("vqb_backward: tail.d_flat, tail.counter and gather_table are required");
            return launch_bwd_reduce(a, nullptr, 0, s, nullptr);
        }
        return VQB_OK;
    }
    const bool stop_grad = a->flags & VQB_STOP_GRAD;
    const bool skip = a->flags & VQB_SKIP;
    const size_t nd_bytes = (size_t)a->n_rows * a->dim * sizeof(float);

    if (!a->g_p && !a->g_logp && stop_grad) {
        // scatter-only backward: nothing reaches the softmax route.
        //   L2:     dx = g_q (straight-through identity; zero bytes when the caller aliases it)
        //   LINEAR: dx = 0   (no path from new_latent to x with stop_grad, src/embed.py:194-197)
        if (a->dx) {
            if (l2) {
                if (a->dx != a->g_q) VQB_CUDA(cudaMemcpyAsync(a->dx, a->g_q, nd_bytes, cudaMemcpyDeviceToDevice, s));
            } else {
                VQB_CUDA(cudaMemsetAsync(a->dx, 0, nd_bytes, s));
            }
        }
        if (skip && l2) return VQB_OK;
        float* dst = l2 ? a->d_score_w : a->d_gather;
        if (!dst) return invalid("vqb_backward: the scatter destination (d_score_w for L2, d_gather for LINEAR) is NULL");
        if (!aligned16(a->g_q) || !aligned16(dst)) return invalid("vqb_backward: tensor pointers must be 16-byte aligned");
        return launch_scatter_add(a->idx, a->n_rows, a->g_q, a->n_codes, a->dim, dst, nullptr, a->workspace, a->workspace_bytes, s);
    }

    if (!a->x || !a->score_w || !a->gather_table || !a->p_code || !a->dx || (!a->tail && (!a->d_score_w || !a->colsum)))
        return invalid("vqb_backward: x, score_w, gather_table, p_code, dx, d_score_w and colsum are required for the p_code route");
    if (l2 && !a->temp) return invalid("vqb_backward: temp is required for the L2 score");
    if ((a->flags & VQB_TEMP_GRAD) && (!a->d_temp || !a->score_b)) return invalid("vqb_backward: VQB_TEMP_GRAD needs d_temp and score_b");
    if (!l2 && a->g_q && !a->d_gather) return invalid("vqb_backward: d_gather is required for the LINEAR score when g_q is given");
    if (!aligned16(a->x) || !aligned16(a->score_w) || !aligned16(a->gather_table) || !aligned16(a->dx) ||
        (a->d_score_w && !aligned16(a->d_score_w)) || (a->g_q && !aligned16(a->g_q)) || (a->d_gather && !aligned16(a->d_gather)))
        return invalid("vqb_backward: tensor pointers must be 16-byte aligned");
    if (a->tail) {
        const vqb_bwd_tail* tl = a->tail;
        if (!l2 || !backward_pcode_supported(a))
            return invalid("vqb_backward: the fused tail needs the L2 score on the vqb_bwd_pcode_kernel route (see vqb_backward_kernel_name)");
        if (!tl->d_flat || !tl->counter) return invalid("vqb_backward: tail.d_flat and tail.counter are required");
        if (tl->phn_attr ? (tl->n_attr <= 0 || tl->n_attr > 63 || tl->dim_attr <= 0 || tl->dim_attr >= a->dim) : (tl->n_attr != 0 || tl->dim_attr != 0))
            return invalid("vqb_backward: tail.phn_attr / n_attr / dim_attr are inconsistent");
        if (tl->world > 1 && (!tl->peer_bufs || tl->rank < 0 || tl->rank >= tl->world || tl->world > VQB_MAX_WORLD))
            return invalid("vqb_backward: tail.world=%d rank=%d needs peer_bufs and world <= %d", tl->world, tl->rank, VQB_MAX_WORLD);
    }
    if (backward_pcode_supported(a)) return launch_backward_pcode(a, s);
    if (backward_generic_needed(a)) return launch_backward_generic(a, s);
    return launch_backward_simt(a, s);
}

extern "C" const char* vqb_backward_kernel_name(const vqb_bwd_args* a) {
    if (validate_bwd(a) != VQB_OK) return "invalid";
    if (!a->g_p && !a->g_logp && (a->flags & VQB_STOP_GRAD)) return "scatter_hist_kernel";
    if (backward_pcode_supported(a)) return "vqb_bwd_pcode_kernel";
    if (backward_generic_needed(a)) return "bwdg_dx_kernel";
    return "vqb_bwd_simt_kernel";
}

extern "C" size_t vqb_exchange_bytes(int64_t n_flat, int32_t world) { return vqb::exchange_bytes(n_flat, world); }

extern "C" int vqb_exchange_finish(const vqb_bwd_tail* tail, int64_t n_flat, void* stream) {
    if (!tail || !tail->d_flat || !tail->counter || n_flat <= 0) return invalid("vqb_exchange_finish: tail.d_flat, tail.counter and n_flat are required");
    if (tail->world > 1 && (!tail->peer_bufs || tail->rank < 0 || tail->rank >= tail->world || tail->world > VQB_MAX_WORLD))
        return invalid("vqb_exchange_finish: tail.world=%d rank=%d needs peer_bufs and world <= %d", tail->world, tail->rank, VQB_MAX_WORLD);
    int rc = require_device();
    if (rc) return rc;
    return launch_exchange_finish(tail, n_flat, (cudaStream_t)stream);
}
