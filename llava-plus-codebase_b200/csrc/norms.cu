// LayerNorm (CLIP, transformers modeling_clip.py:359-384 layer_norm1/2, :677 pre_layrnorm) and RMSNorm
// (LLaMA, transformers modeling_llama.py:62-67). One warp per row, 16-byte vector loads, fp32 statistics.
// HBM-bound elementwise work: rows are read once and written once.
#include "common.cuh"
#include "kernels.h"

namespace b2 {
namespace {

constexpr int kWarpsPerBlock = 4;

// cols % 256 == 0 (each lane handles cols/32 elements in 16 B vectors); cols <= 8192
template <int MAX_VEC>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
layernorm_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ gamma,
                 const __nv_bfloat16* __restrict__ beta, __nv_bfloat16* __restrict__ y, int rows, int cols,
                 float eps) {
    pdl_trigger();
    pdl_wait();  // inputs are outputs of the upstream kernel (programmatic dependent launch)
    const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int nvec = cols / 256;
    const __nv_bfloat16* xr = x + (size_t)row * cols;
    float v[MAX_VEC][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_VEC; ++i) {
        if (i < nvec) {
            const uint4 u = *reinterpret_cast<const uint4*>(xr + i * 256 + lane * 8);
            v[i][0] = bf16_lo(u.x); v[i][1] = bf16_hi(u.x); v[i][2] = bf16_lo(u.y); v[i][3] = bf16_hi(u.y);
            v[i][4] = bf16_lo(u.z); v[i][5] = bf16_hi(u.z); v[i][6] = bf16_lo(u.w); v[i][7] = bf16_hi(u.w);
#pragma unroll
            for (int e = 0; e < 8; ++e) sum += v[i][e];
        }
    }
    const float mean = warp_sum(sum) / cols;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_VEC; ++i) {
        if (i < nvec) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float d = v[i][e] - mean;
                var += d * d;
            }
        }
    }
    const float rstd = rsqrtf(warp_sum(var) / cols + eps);
    __nv_bfloat16* yr = y + (size_t)row * cols;
#pragma unroll
    for (int i = 0; i < MAX_VEC; ++i) {
        if (i < nvec) {
            const int c = i * 256 + lane * 8;
            const uint4 g = *reinterpret_cast<const uint4*>(gamma + c);
            const uint4 b = *reinterpret_cast<const uint4*>(beta + c);
            const float gg[8] = {bf16_lo(g.x), bf16_hi(g.x), bf16_lo(g.y), bf16_hi(g.y),
                                 bf16_lo(g.z), bf16_hi(g.z), bf16_lo(g.w), bf16_hi(g.w)};
            const float bb[8] = {bf16_lo(b.x), bf16_hi(b.x), bf16_lo(b.y), bf16_hi(b.y),
                                 bf16_lo(b.z), bf16_hi(b.z), bf16_lo(b.w), bf16_hi(b.w)};
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = (v[i][e] - mean) * rstd * gg[e] + bb[e];
            *reinterpret_cast<uint4*>(yr + c) = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]),
                                                           pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
        }
    }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
rmsnorm_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_row_stride, const int32_t* __restrict__ row_index,
               const __nv_bfloat16* __restrict__ gamma, __nv_bfloat16* __restrict__ y, int rows, int cols,
               float eps) {
    pdl_trigger();
    pdl_wait();  // inputs are outputs of the upstream kernel (programmatic dependent launch)
    const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int nvec = cols / 256;
    const int64_t src_row = row_index != nullptr ? (int64_t)row_index[row] : (int64_t)row;
    const __nv_bfloat16* xr = x + src_row * x_row_stride;
    // pass 1: sum of squares (the row is re-read from L1/L2 in pass 2; keeps registers small for h=5120)
    float ss = 0.f;
    for (int i = 0; i < nvec; ++i) {
        const uint4 u = *reinterpret_cast<const uint4*>(xr + i * 256 + lane * 8);
        const float v[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y),
                            bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
        for (int e = 0; e < 8; ++e) ss += v[e] * v[e];
    }
    const float rstd = rsqrtf(warp_sum(ss) / cols + eps);
    __nv_bfloat16* yr = y + (size_t)row * cols;
    for (int i = 0; i < nvec; ++i) {
        const int c = i * 256 + lane * 8;
        const uint4 u = *reinterpret_cast<const uint4*>(xr + c);
        const uint4 g = *reinterpret_cast<const uint4*>(gamma + c);
        const float v[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y),
                            bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
        const float gg[8] = {bf16_lo(g.x), bf16_hi(g.x), bf16_lo(g.y), bf16_hi(g.y),
                             bf16_lo(g.z), bf16_hi(g.z), bf16_lo(g.w), bf16_hi(g.w)};
        float o[8];
        // modeling_llama.py:62-67: weight * hidden_states.to(input_dtype)  -> two bf16 roundings
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = gg[e] * round_bf16(v[e] * rstd);
        *reinterpret_cast<uint4*>(yr + c) = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]),
                                                       pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
    }
}

// One CTA (256 threads) per row, single pass with the row in registers: used when there are too few rows for the
// warp-per-row kernel to fill the machine (a B=1 prefill has 704 rows; one warp walking a 4096-wide row is a chain of
// dependent L2 round trips: 11 us measured for 11 MB of traffic). Same arithmetic and rounding points as rmsnorm_kernel;
// cols % 8 == 0, cols <= 8192.
__global__ void __launch_bounds__(256)
rmsnorm_row_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_row_stride, const int32_t* __restrict__ row_index,
                   const __nv_bfloat16* __restrict__ gamma, __nv_bfloat16* __restrict__ y, int cols, float eps) {
    __shared__ float s_part[8];
    const int row = blockIdx.x, tid = threadIdx.x;
    pdl_trigger();
    pdl_wait();  // x is the previous kernel's output
    const int64_t src_row = row_index != nullptr ? (int64_t)row_index[row] : (int64_t)row;
    const __nv_bfloat16* xr = x + src_row * x_row_stride;
    const int nvec = cols >> 3;
    uint4 u[4];
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = j * 256 + tid;
        u[j] = i < nvec ? *reinterpret_cast<const uint4*>(xr + i * 8) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float v[8] = {bf16_lo(u[j].x), bf16_hi(u[j].x), bf16_lo(u[j].y), bf16_hi(u[j].y),
                            bf16_lo(u[j].z), bf16_hi(u[j].z), bf16_lo(u[j].w), bf16_hi(u[j].w)};
#pragma unroll
        for (int e = 0; e < 8; ++e) ss += v[e] * v[e];
    }
    ss = warp_sum(ss);
    if ((tid & 31) == 0) s_part[tid >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += s_part[w];
    const float rstd = rsqrtf(tot / cols + eps);
    __nv_bfloat16* yr = y + (size_t)row * cols;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = j * 256 + tid;
        if (i < nvec) {
            const uint4 g = *reinterpret_cast<const uint4*>(gamma + i * 8);
            const float v[8] = {bf16_lo(u[j].x), bf16_hi(u[j].x), bf16_lo(u[j].y), bf16_hi(u[j].y),
                                bf16_lo(u[j].z), bf16_hi(u[j].z), bf16_lo(u[j].w), bf16_hi(u[j].w)};
            const float gg[8] = {bf16_lo(g.x), bf16_hi(g.x), bf16_lo(g.y), bf16_hi(g.y),
                                 bf16_lo(g.z), bf16_hi(g.z), bf16_lo(g.w), bf16_hi(g.w)};
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = gg[e] * round_bf16(v[e] * rstd);
            *reinterpret_cast<uint4*>(yr + i * 8) = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]),
                                                               pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
        }
    }
}

}  // namespace

int layernorm_bf16(const void* x, const void* gamma, const void* beta, void* y, int rows, int cols, float eps,
                   cudaStream_t stream) {
    B2_CHECK_ARG(rows > 0 && cols > 0 && cols % 256 == 0 && cols <= 2048,
                 "layernorm: cols must be a multiple of 256 and <= 2048 (cols=%d rows=%d)", cols, rows);
    const int grid = (rows + kWarpsPerBlock - 1) / kWarpsPerBlock;
    B2_CUDA_CHECK(launch_pdl(layernorm_kernel<8>, dim3(grid), dim3(kWarpsPerBlock * 32), 0, stream,
        reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(gamma),
        reinterpret_cast<const __nv_bfloat16*>(beta), reinterpret_cast<__nv_bfloat16*>(y), rows, cols, eps));
    B2_LAUNCH_CHECK();
    return 0;
}

static int rmsnorm_launch(const void* x, int64_t stride, const int32_t* row_index, const void* gamma, void* y,
                          int rows, int cols, float eps, cudaStream_t stream) {
    B2_CHECK_ARG(rows > 0 && cols > 0 && cols % 256 == 0,
                 "rmsnorm: cols must be a multiple of 256 (cols=%d rows=%d)", cols, rows);
    if (rows < 4096 && cols <= 8192) {  // few rows: a CTA per row keeps every SM busy and the row in registers
        B2_CUDA_CHECK(launch_pdl(rmsnorm_row_kernel, dim3(rows), dim3(256), 0, stream, reinterpret_cast<const __nv_bfloat16*>(x),
                                 stride, row_index, reinterpret_cast<const __nv_bfloat16*>(gamma),
                                 reinterpret_cast<__nv_bfloat16*>(y), cols, eps));
        B2_LAUNCH_CHECK();
        return 0;
    }
    const int grid = (rows + kWarpsPerBlock - 1) / kWarpsPerBlock;
    B2_CUDA_CHECK(launch_pdl(rmsnorm_kernel, dim3(grid), dim3(kWarpsPerBlock * 32), 0, stream,
        reinterpret_cast<const __nv_bfloat16*>(x), stride, row_index,
        reinterpret_cast<const __nv_bfloat16*>(gamma), reinterpret_cast<__nv_bfloat16*>(y), rows, cols, eps));
    B2_LAUNCH_CHECK();
    return 0;
}

int rmsnorm_bf16(const void* x, int64_t x_row_stride, const void* gamma, void* y, int rows, int cols, float eps,
                 cudaStream_t stream) {
    return rmsnorm_launch(x, x_row_stride, nullptr, gamma, y, rows, cols, eps, stream);
}

int rmsnorm_gather_bf16(const void* x, const int32_t* row_index, const void* gamma, void* y, int rows, int cols,
                        float eps, cudaStream_t stream) {
    return rmsnorm_launch(x, cols, row_index, gamma, y, rows, cols, eps, stream);
}

}  // namespace b2
