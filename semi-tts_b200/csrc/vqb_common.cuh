// Shared helpers for libvqb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include "vqb.h"

namespace vqb {

// thread-local error string behind vqb_last_error()
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
int invalid(const char* fmt, ...);

#define VQB_CUDA(call)                                              \
    do {                                                            \
        cudaError_t e_ = (call);                                    \
        if (e_ != cudaSuccess) return ::vqb::cuda_fail(e_, #call);  \
    } while (0)

void count_launch();
void kernel_event_begin(cudaStream_t s);   // developer hook (vqb_debug_set_kernel_events): events around the dominant kernel
void kernel_event_end(cudaStream_t s);                  // host-side tally behind vqb_launch_count()

#define VQB_CHECK_LAUNCH(name)                                      \
    do {                                                            \
        ::vqb::count_launch();                                      \
        cudaError_t e_ = cudaGetLastError();                        \
        if (e_ != cudaSuccess) return ::vqb::cuda_fail(e_, name);   \
    } while (0)

int sm_count();                       // SMs of the current device (148 on B200)
int max_optin_smem();                 // bytes of dynamic smem a CTA may opt in to (227 KB)

// launch with the programmatic-stream-serialization attribute (see pdl_wait / pdl_launch below)
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool off = getenv("VQB_NO_PDL") != nullptr;       // developer switch: plain stream-ordered launches
    cfg.attrs = at; cfg.numAttrs = off ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize, PreferredSharedMemoryCarveout) once per (kernel instantiation, device,
// size) instead of on every launch: the attribute calls cost more than the launch itself on the eager encode path
template <typename Kern>
static inline int ensure_smem(Kern kern, size_t smem, bool max_carveout) {
    static thread_local struct { Kern k; int dev; size_t smem; } memo[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    for (auto& m : memo)
        if (m.k == kern && m.dev == dev && m.smem >= smem) return VQB_OK;
    VQB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (max_carveout) VQB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    for (auto& m : memo)
        if (m.k == nullptr || (m.k == kern && m.dev == dev)) { m.k = kern; m.dev = dev; m.smem = smem; break; }
    return VQB_OK;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device helpers ------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// streaming (read-once) 128-bit load: do not allocate in L1
__device__ __forceinline__ float4 ldg4_stream(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg4_stream(float* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, const float4& v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// programmatic dependent launch (PDL): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may
// start while its predecessor in the stream is still running; pdl_wait() blocks until that predecessor has completed
// and its memory is visible (a no-op without the attribute), pdl_launch() lets the successor start early.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float tf32_rn(float v) {      // round to tf32 (10-bit mantissa), ties away
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}
__device__ __forceinline__ float tf32_trunc(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

// entry points implemented in the individual .cu files (host side, return VQB_* codes)
int launch_forward_simt(const vqb_fwd_args* a, cudaStream_t s);
int launch_backward_simt(const vqb_bwd_args* a, cudaStream_t s);
int launch_scatter_add(const int64_t* idx, int64_t n, const float* g, int64_t K, int64_t D,
                       float* dtable, int64_t* hist, void* workspace, size_t workspace_bytes, cudaStream_t s);
size_t scatter_workspace_bytes(int64_t n, int64_t K, int64_t D);
bool forward_tensor_supported(const vqb_fwd_args* a);
int forward_tensor_workspace(const vqb_fwd_args* a, size_t* bytes);
int launch_forward_tensor(const vqb_fwd_args* a, cudaStream_t s);
size_t exchange_bytes(int64_t n_flat, int world);
int launch_exchange_finish(const vqb_bwd_tail* tl, int64_t n_flat, cudaStream_t s);
int launch_bwd_reduce(const vqb_bwd_args* a, const float* partial, int grid, cudaStream_t s, unsigned long long* dbg);
// third generation of the parity-mode kernels (vqb_fwd_pc.cu, vqb_bwd_pc.cu): one tile per CTA at a time, several CTAs per SM
bool forward_pcode_supported(const vqb_fwd_args* a);
size_t forward_pcode_workspace(const vqb_fwd_args* a);
int launch_forward_pcode(const vqb_fwd_args* a, cudaStream_t s);
bool backward_pcode_supported(const vqb_bwd_args* a);
size_t backward_pcode_workspace(const vqb_bwd_args* a);
int launch_backward_pcode(const vqb_bwd_args* a, cudaStream_t s);
// any-K p_code-route backward through a coefficient matrix in the workspace (vqb_bwd_generic.cu)
bool backward_generic_needed(const vqb_bwd_args* a);
size_t backward_generic_workspace(const vqb_bwd_args* a);
int launch_backward_generic(const vqb_bwd_args* a, cudaStream_t s);

}  // namespace vqb
