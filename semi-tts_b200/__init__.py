"""semi-tts_b200 -- the semi-tts vector-quantisation bottleneck, B200-native.

Drop-in replacements for the reference's quantizer modules (src/embed.py: L2Embedding :57-147,
SeperateEmbedding :150-205) whose forward/backward run in hand-written sm_100a CUDA kernels behind a
C ABI (include/vqb.h, libvqb200.so).  There is no CPU path: every op raises if the library or a
B200 is missing.
"""
from . import _lib
from .functional import (vq_l2, vq_linear, codebook_lookup, assemble_table, vq_search)
from .embed import L2Embedding, SeperateEmbedding, read_phn_attr
from .usage import UsageHistogram
from .segment import mean_forward, row_argmax, ctc_log_probs
from .patch import install_into_reference, uninstall_from_reference
from . import dist

__all__ = ["L2Embedding", "SeperateEmbedding", "read_phn_attr", "vq_l2", "vq_linear", "vq_search",
           "codebook_lookup", "assemble_table", "UsageHistogram", "mean_forward", "row_argmax", "ctc_log_probs", "install_into_reference", "uninstall_from_reference", "dist", "_lib"]
