// e4m3 row quantisation for the fp8 decode path (BASELINE configs[4]: "fp8-weight tcgen05 path"):
//
//     scale[r] = max_k |x[r,k]| / 448        (1.0 for an all-zero row)
//     q[r,k]   = e4m3_rn_satfinite(x[r,k] * (448 / max_k |x[r,k]|))
//
// used (a) once per weight matrix at b2_model_enable_fp8_decode (per-OUTPUT-CHANNEL scales: a weight row is a "row"
// here) and (b) per decode step on the activations (per-TOKEN scales), fused with RMSNorm where the GEMM input is a
// normed hidden state. The reference has no fp8 path: the test suite carries a CPU restatement of exactly these two
// formulas (torch.float8_e4m3fn, round-to-nearest-even) and derives the tolerance of the whole path against the bf16
// path from it.
#include <cuda_fp8.h>

#include "common.cuh"
#include "kernels.h"

namespace b2 {
namespace {

constexpr float kE4M3Max = 448.0f;

__device__ __forceinline__ float block_max_256(float v, float* s_part) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
    __syncthreads();
    float m = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) m = fmaxf(m, s_part[w]);
    __syncthreads();  // s_part is reused
    return m;
}

__device__ __forceinline__ uint2 pack8_e4m3(const float (&f)[8], float inv) {
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const __nv_fp8x2_storage_t a = __nv_cvt_float2_to_fp8x2(make_float2(f[4 * e] * inv, f[4 * e + 1] * inv), __NV_SATFINITE, __NV_E4M3);
        const __nv_fp8x2_storage_t b = __nv_cvt_float2_to_fp8x2(make_float2(f[4 * e + 2] * inv, f[4 * e + 3] * inv), __NV_SATFINITE, __NV_E4M3);
        const uint32_t w = (uint32_t)a | ((uint32_t)b << 16);
        if (e == 0) lo = w; else hi = w;
    }
    return make_uint2(lo, hi);
}

__device__ __forceinline__ void unpack8f(const uint4& u, float (&f)[8]) {
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
    f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}

// One CTA (256 threads) per row; K % 8 == 0. Two passes over the row (the second one hits L1/L2).
__global__ void __launch_bounds__(256)
quantize_rows_e4m3_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, int K, uint8_t* __restrict__ q, int64_t ldq,
                          float* __restrict__ scale) {
    __shared__ float s_part[8];
    const int row = blockIdx.x, tid = threadIdx.x;
    const __nv_bfloat16* xr = x + (size_t)row * ldx;
    const int nvec = K >> 3;
    float amax = 0.f;
    for (int i = tid; i < nvec; i += 256) {
        float f[8];
        unpack8f(*reinterpret_cast<const uint4*>(xr + i * 8), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) amax = fmaxf(amax, fabsf(f[e]));
    }
    amax = block_max_256(amax, s_part);
    const float inv = amax > 0.f ? kE4M3Max / amax : 1.0f;
    if (tid == 0) scale[row] = amax > 0.f ? amax / kE4M3Max : 1.0f;
    uint8_t* qr = q + (size_t)row * ldq;
    for (int i = tid; i < nvec; i += 256) {
        float f[8];
        unpack8f(*reinterpret_cast<const uint4*>(xr + i * 8), f);
        *reinterpret_cast<uint2*>(qr + i * 8) = pack8_e4m3(f, inv);
    }
}

// RMSNorm (HF semantics, as rmsnorm_row_kernel: y = gamma * bf16(x * rstd), y rounded to bf16) followed by the per-token
// quantisation of y, row in registers; cols % 8 == 0, cols <= 8192.
__global__ void __launch_bounds__(256)
rmsnorm_quant_row_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_row_stride, const __nv_bfloat16* __restrict__ gamma,
                         uint8_t* __restrict__ q, int64_t ldq, float* __restrict__ scale, int cols, float eps) {
    __shared__ float s_part[8];
    const int row = blockIdx.x, tid = threadIdx.x;
    const __nv_bfloat16* xr = x + (size_t)row * x_row_stride;
    const int nvec = cols >> 3;
    float y[4][8];
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = j * 256 + tid;
        const uint4 u = i < nvec ? *reinterpret_cast<const uint4*>(xr + i * 8) : make_uint4(0, 0, 0, 0);
        unpack8f(u, y[j]);
#pragma unroll
        for (int e = 0; e < 8; ++e) ss += y[j][e] * y[j][e];
    }
    ss = warp_sum(ss);
    if ((tid & 31) == 0) s_part[tid >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += s_part[w];
    __syncthreads();
    const float rstd = rsqrtf(tot / cols + eps);
    float amax = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = j * 256 + tid;
        if (i < nvec) {
            float g[8];
            unpack8f(*reinterpret_cast<const uint4*>(gamma + i * 8), g);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                y[j][e] = round_bf16(g[e] * round_bf16(y[j][e] * rstd));
                amax = fmaxf(amax, fabsf(y[j][e]));
            }
        }
    }
    amax = block_max_256(amax, s_part);
    const float inv = amax > 0.f ? kE4M3Max / amax : 1.0f;
    if (tid == 0) scale[row] = amax > 0.f ? amax / kE4M3Max : 1.0f;
    uint8_t* qr = q + (size_t)row * ldq;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = j * 256 + tid;
        if (i < nvec) *reinterpret_cast<uint2*>(qr + i * 8) = pack8_e4m3(y[j], inv);
    }
}

}  // namespace

int quantize_rows_e4m3(const void* x, int64_t ldx, int rows, int K, void* q, int64_t ldq, float* scale, cudaStream_t stream) {
    B2_CHECK_ARG(rows > 0 && K > 0 && K % 8 == 0 && ldx % 8 == 0 && ldq % 8 == 0,
                 "quantize_rows_e4m3: rows=%d K=%d must be positive, K and the row pitches multiples of 8", rows, K);
    B2_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) & 15) | (reinterpret_cast<uintptr_t>(q) & 7)) == 0,
                 "quantize_rows_e4m3: x must be 16-byte and q 8-byte aligned");
    quantize_rows_e4m3_kernel<<<rows, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), ldx, K,
                                                         reinterpret_cast<uint8_t*>(q), ldq, scale);
    B2_LAUNCH_CHECK();
    return 0;
}

int rmsnorm_quant_e4m3(const void* x, int64_t x_row_stride, const void* gamma, void* q, int64_t ldq, float* scale, int rows,
                       int cols, float eps, cudaStream_t stream) {
    B2_CHECK_ARG(rows > 0 && cols > 0 && cols % 8 == 0 && cols <= 8192 && ldq % 8 == 0,
                 "rmsnorm_quant_e4m3: cols must be a multiple of 8, <= 8192 (cols=%d rows=%d)", cols, rows);
    rmsnorm_quant_row_kernel<<<rows, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), x_row_stride,
                                                        reinterpret_cast<const __nv_bfloat16*>(gamma),
                                                        reinterpret_cast<uint8_t*>(q), ldq, scale, cols, eps);
    B2_LAUNCH_CHECK();
    return 0;
}

}  // namespace b2
