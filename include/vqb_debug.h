/* libvqb200 -- developer / measurement hooks.  NOT part of the drop-in ABI of include/vqb.h: nothing the reference's call
 * sites need is declared here, no parity claim rests on these, and they may change without an ABI version bump.  They are
 * exported so that bench.py, tools/ and tests/ can time single kernels and read in-kernel timelines without a debugger.
 * All of them set process-global state and are not thread-safe. */
#ifndef VQB_DEBUG_H
#define VQB_DEBUG_H
#include "vqb.h"

#ifdef __cplusplus
extern "C" {
#endif

/* A pair of CUDA events (cudaEvent_t) that the NEXT vqb_forward / vqb_backward / vqb_scatter_add call records on its launch
 * stream immediately before and after its dominant kernel (not around its helper kernels): bench.py's roofline record times
 * that kernel alone with them.  Pass NULL, NULL to switch the hook off. */
VQB_API void vqb_debug_set_kernel_events(void* ev_start, void* ev_stop);

/* Device buffer of at least 1024 uint64 words that the tcgen05 kernels fill with (tag << 56 | globaltimer ns) marks of their
 * phases (CTA 0: slots 0..59; tail kernel: slots 100..; per-CTA entry / exit: slots 128..): tools/timeline_pc.py.  NULL = off. */
VQB_API void vqb_debug_set_timeline(void* dev_ptr);

/* A/B switch of the streamed 3xTF32 search: force the software-pipelined x_lo on (1) or off (0); -1 = default. */
VQB_API void vqb_debug_set_search_pipe(int v);

#ifdef __cplusplus
}
#endif
#endif
