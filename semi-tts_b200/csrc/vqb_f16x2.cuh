// "fp16x2" operands for the parity-mode kernels (K <= 64, D <= 64): a value is carried as two fp16 pieces after an exact
// power-of-two rescale, v * 2^s = hi + lo with hi = the top 11 significant bits (exact in fp16) and lo = the rounded
// remainder (22 significant bits in all), so one burst of kind::f16 tcgen05 MMAs over the four piece products gives
// fp32-level accuracy.  A 16-bit operand tile [rows][64 values] is 128 bytes per row, and with the 128-byte swizzle the
// K-major and the MN-major layouts of such a tile are the SAME bytes -- one image of the score table serves the forward
// (B operand, contraction over d) and GEMM 1 of the backward (B operand, contraction over codes).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>
#include "vqb_tc.cuh"

namespace vqb {

// Operand image of a score table [K <= 64][D <= 64] in global memory, written once per step by the table assembly
// (or by build_image_kernel for a table that is not assembled here):
//   [0, 8192)      hi pieces   [64 code rows][64 halves], 128-byte swizzle (chunk j of row k at sw128_offset(k, j))
//   [8192, 16384)  lo pieces   same layout
//   [16384, +64)   header: int32 gE (table scale: image = table * 2^-gE), float emax (max_k |w_k|_2)
//   [16448, +256)  the score bias of the 64 codes: |e_k|^2 (L2) or b_k (LINEAR); entries >= K are 0
// Rows >= K and columns >= D are zero.
constexpr int IMG_PIECE = 64 * 128;
constexpr int IMG_HDR = 2 * IMG_PIECE;
constexpr int IMG_BIAS = IMG_HDR + 64;
constexpr int IMG_BYTES = IMG_BIAS + 256;

// floor(log2(|v|)) of a positive normal float; subnormals and zero give -127, inf/nan give 128
__device__ __forceinline__ int exp_of(float v) { return (int)((__float_as_uint(v) >> 23) & 0xFFu) - 127; }
// 2^n for n in [-126, 127] (clamped)
__device__ __forceinline__ float pow2i(int n) {
    n = n < -126 ? -126 : (n > 127 ? 127 : n);
    return __uint_as_float((uint32_t)(n + 127) << 23);
}
__device__ __forceinline__ float f16_trunc(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
// eight scaled values -> their fp16 hi pieces (exact: 11 significant bits) and lo pieces (rounded remainder)
__device__ __forceinline__ void split8(const float* v, float s, uint4& hi, uint4& lo) {
    float h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float w = v[i] * s;
        h[i] = f16_trunc(w);
        l[i] = w - h[i];
    }
    hi = make_uint4(pack_h2(h[0], h[1]), pack_h2(h[2], h[3]), pack_h2(h[4], h[5]), pack_h2(h[6], h[7]));
    lo = make_uint4(pack_h2(l[0], l[1]), pack_h2(l[2], l[3]), pack_h2(l[4], l[5]), pack_h2(l[6], l[7]));
}
// scale exponent that brings a positive maximum into [2^14, 2^15): v * 2^-e; 0 for a zero / non-finite maximum
__device__ __forceinline__ int scale_exp(float mx) {
    if (!(mx > 0.f) || !(mx < INFINITY)) return 0;
    int e = exp_of(mx) - 14;
    return e < -126 ? -126 : (e > 126 ? 126 : e);
}

// device part of the image build, called by every thread of ONE CTA: `tab` is the table in shared or global memory
// ([K][ld] floats), gmax = max |tab|, emax = max row norm (both block-uniform)
__device__ __forceinline__ void write_image(const float* tab, int ld, int K, int D, float gmax, float emax, const float* bias,
                                            uint8_t* img) {
    const int gE = scale_exp(gmax);
    const float sE = pow2i(-gE);
    for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) {
        const int k = i >> 3, j = i & 7;
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (k < K && 8 * j + u < D) ? tab[k * ld + 8 * j + u] : 0.f;
        uint4 hi, lo;
        split8(v, sE, hi, lo);
        *reinterpret_cast<uint4*>(img + tc::sw128_offset(k, j)) = hi;
        *reinterpret_cast<uint4*>(img + IMG_PIECE + tc::sw128_offset(k, j)) = lo;
    }
    if (threadIdx.x == 0) {
        *reinterpret_cast<int*>(img + IMG_HDR) = gE;
        *reinterpret_cast<float*>(img + IMG_HDR + 4) = emax;
    }
    if (threadIdx.x < 64) reinterpret_cast<float*>(img + IMG_BIAS)[threadIdx.x] = ((int)threadIdx.x < K && bias) ? bias[threadIdx.x] : 0.f;
}

// host: enqueue the image build for a table that was not assembled by vqb_assemble_table (LINEAR score, raw C-ABI calls)
int launch_build_image(const float* w, const float* bias, int K, int D, void* img, cudaStream_t s);

}  // namespace vqb
