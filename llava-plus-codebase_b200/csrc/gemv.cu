// Weight-streaming GEMV for the one-token decode step (batch <= 8):
//     out[B, N] = (RMSNorm(x) | x)[B, K] · W[N, K]^T  (+ residual)      or the fused SwiGLU variant.
// This is the HBM-bound hot loop of LLaVA decode: every weight byte is read
// exactly once per step with 16-byte coalesced, L1-bypassing loads; a persistent grid of one 512-thread CTA
// per SM keeps >= 8 independent 16 B loads in flight per lane through a register double buffer, and the
// first weight loads are issued BEFORE the activation prologue (x -> smem, fused RMSNorm), since weights do
// not depend on the previous kernel's output. Tensor cores are deliberately not used: at B <= 8 the op is
// >20x below the tensor roofline and purely bandwidth-limited.
// Reference math replaced: transformers modeling_llama.py:62-67 (RMSNorm), :182-184 (LlamaMLP),
// :251-289 (q/k/v/o projections), llava_llama.py:48 (lm_head).
#include "common.cuh"
#include "kernels.h"

namespace b2 {
namespace {

constexpr int GV_THREADS = 512;
constexpr int GV_WARPS = GV_THREADS / 32;
constexpr int GV_R = 2;  // weight rows per work item
constexpr int GV_U = 4;  // 256-element K-chunks per pipeline step

struct GemvParams {
    const __nv_bfloat16* x; int64_t ldx;
    const __nv_bfloat16* W; int ldw;
    const __nv_bfloat16* gamma; float eps;
    const __nv_bfloat16* residual; int ld_res;
    void* out; int ld_out; int out_fp32;
    int B, N, K, act;
};

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
    f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}

template <int NB>
__global__ void __launch_bounds__(GV_THREADS, 1) gemv_kernel(GemvParams p) {
    extern __shared__ __align__(16) uint8_t gv_smem[];
    __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(gv_smem);  // [NB][K]
    __shared__ float s_red[GV_WARPS][NB];
    __shared__ float s_rstd[NB];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gw = blockIdx.x * GV_WARPS + warp;
    const int total_warps = gridDim.x * GV_WARPS;
    const int K = p.K;
    const int nchunks = K >> 8;                       // K % 256 == 0
    const int G = (nchunks + GV_U - 1) / GV_U;        // pipeline steps per item
    const int n_items = p.N / GV_R;
    const int n_my = gw < n_items ? (n_items - gw + total_warps - 1) / total_warps : 0;
    const int total_steps = n_my * G;
    const bool swiglu = p.act == ACT_SWIGLU;

    auto item_rows = [&](int item, int& r0, int& r1) {
        if (swiglu) {  // item = output channel; gate/up rows are block-64 interleaved
            r0 = (item >> 6) * 128 + (item & 63);
            r1 = r0 + 64;
        } else {
            r0 = item * 2;
            r1 = r0 + 1;
        }
    };
    auto issue = [&](int s, uint4 (&buf)[GV_R][GV_U]) {
        if (s < total_steps) {
            const int item = gw + (s / G) * total_warps;
            const int g = s % G;
            int rr[2];
            item_rows(item, rr[0], rr[1]);
#pragma unroll
            for (int r = 0; r < GV_R; ++r) {
                const __nv_bfloat16* wr = p.W + (size_t)rr[r] * p.ldw + lane * 8;
#pragma unroll
                for (int u = 0; u < GV_U; ++u) {
                    const int ch = g * GV_U + u;
                    buf[r][u] = (ch < nchunks) ? ld_stream_16(wr + ch * 256) : make_uint4(0, 0, 0, 0);
                }
            }
        }
    };

    uint4 bufA[GV_R][GV_U], bufB[GV_R][GV_U];
    issue(0, bufA);  // weights do not depend on x: get HBM requests in flight before the prologue

    // ---------------- prologue: x -> smem (bf16), optional fused RMSNorm ----------------
    {
        const int nvec = K >> 3;
        float ss[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) ss[b] = 0.f;
        for (int i = tid; i < nvec; i += GV_THREADS) {
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                uint4 u = make_uint4(0, 0, 0, 0);
                if (b < p.B) u = *reinterpret_cast<const uint4*>(p.x + (size_t)b * p.ldx + i * 8);
                *reinterpret_cast<uint4*>(xs + (size_t)b * K + i * 8) = u;
                if (p.gamma != nullptr) {
                    float f[8];
                    unpack8(u, f);
#pragma unroll
                    for (int e = 0; e < 8; ++e) ss[b] += f[e] * f[e];
                }
            }
        }
        if (p.gamma != nullptr) {
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const float v = warp_sum(ss[b]);
                if (lane == 0) s_red[warp][b] = v;
            }
            __syncthreads();
            if (tid < NB) {
                float t = 0.f;
                for (int w = 0; w < GV_WARPS; ++w) t += s_red[w][tid];
                s_rstd[tid] = rsqrtf(t / K + p.eps);
            }
            __syncthreads();
            for (int i = tid; i < nvec; i += GV_THREADS) {
                const uint4 gq = *reinterpret_cast<const uint4*>(p.gamma + i * 8);
                float gf[8];
                unpack8(gq, gf);
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    uint4* px = reinterpret_cast<uint4*>(xs + (size_t)b * K + i * 8);
                    float f[8];
                    unpack8(*px, f);
                    const float rstd = s_rstd[b];
                    float o[8];
                    // HF LlamaRMSNorm: weight * (x * rstd).to(bf16), result in bf16
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] = gf[e] * round_bf16(f[e] * rstd);
                    *px = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]),
                                     pack_bf16(o[6], o[7]));
                }
            }
        }
        __syncthreads();
    }

    // ---------------- main loop ----------------
    float acc[GV_R][NB];
    auto compute = [&](int s, const uint4 (&buf)[GV_R][GV_U]) {
        if (s >= total_steps) return;
        const int item = gw + (s / G) * total_warps;
        const int g = s % G;
        if (g == 0) {
#pragma unroll
            for (int r = 0; r < GV_R; ++r)
#pragma unroll
                for (int b = 0; b < NB; ++b) acc[r][b] = 0.f;
        }
#pragma unroll
        for (int u = 0; u < GV_U; ++u) {
            const int ch = g * GV_U + u;
            if (ch < nchunks) {
                float w0[8], w1[8];
                unpack8(buf[0][u], w0);
                unpack8(buf[1][u], w1);
                const int koff = ch * 256 + lane * 8;
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    float xf[8];
                    unpack8(*reinterpret_cast<const uint4*>(xs + (size_t)b * K + koff), xf);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        acc[0][b] = fmaf(w0[e], xf[e], acc[0][b]);
                        acc[1][b] = fmaf(w1[e], xf[e], acc[1][b]);
                    }
                }
            }
        }
        if (g == G - 1) {
#pragma unroll
            for (int r = 0; r < GV_R; ++r)
#pragma unroll
                for (int b = 0; b < NB; ++b) acc[r][b] = warp_sum(acc[r][b]);
            int r0, r1;
            item_rows(item, r0, r1);
            if (swiglu) {
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    if (lane == b && b < p.B) {
                        const float gte = acc[0][b], up = acc[1][b];
                        const float y = gte / (1.0f + __expf(-gte)) * up;
                        reinterpret_cast<__nv_bfloat16*>(p.out)[(size_t)b * p.ld_out + item] =
                            __float2bfloat16_rn(y);
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < GV_R; ++r) {
#pragma unroll
                    for (int b = 0; b < NB; ++b) {
                        if (lane == r * NB + b && b < p.B) {
                            const int row = r == 0 ? r0 : r1;
                            float y = acc[r][b];
                            if (p.residual != nullptr)
                                y += __bfloat162float(p.residual[(size_t)b * p.ld_res + row]);
                            if (p.out_fp32)
                                reinterpret_cast<float*>(p.out)[(size_t)b * p.ld_out + row] = y;
                            else
                                reinterpret_cast<__nv_bfloat16*>(p.out)[(size_t)b * p.ld_out + row] =
                                    __float2bfloat16_rn(y);
                        }
                    }
                }
            }
        }
    };

    for (int s = 0; s < total_steps; s += 2) {
        issue(s + 1, bufB);
        compute(s, bufA);
        issue(s + 2, bufA);
        compute(s + 1, bufB);
    }
}

template <int NB>
int launch_gemv(const GemvParams& p, cudaStream_t stream) {
    const size_t smem = (size_t)NB * p.K * 2;
    B2_CHECK_ARG(smem <= 200 * 1024 + 24 * 1024, "gemv: activation tile does not fit shared memory (B=%d K=%d)",
                 p.B, p.K);
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        B2_CUDA_CHECK(cudaFuncSetAttribute(gemv_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)(smem > 48 * 1024 ? smem : 48 * 1024)));
        attr_smem = smem > 48 * 1024 ? smem : 48 * 1024;
    }
    gemv_kernel<NB><<<num_sms(), GV_THREADS, smem, stream>>>(p);
    B2_LAUNCH_CHECK();
    return 0;
}

}  // namespace

int gemv_bf16(const GemvArgs& g, cudaStream_t stream) {
    B2_CHECK_ARG(g.B >= 1 && g.B <= 8, "gemv: batch must be 1..8 (got %d)", g.B);
    B2_CHECK_ARG(g.K % 256 == 0 && g.K > 0, "gemv: K must be a positive multiple of 256 (K=%d)", g.K);
    B2_CHECK_ARG(g.N % 2 == 0 && g.N > 0, "gemv: N must be even (N=%d)", g.N);
    B2_CHECK_ARG(g.act == ACT_NONE || g.act == ACT_SWIGLU, "gemv: unsupported activation %d", g.act);
    B2_CHECK_ARG(g.act != ACT_SWIGLU || (g.N % 128 == 0 && g.residual == nullptr && !g.out_fp32),
                 "gemv: swiglu needs N %% 128 == 0, no residual, bf16 output");
    B2_CHECK_ARG((reinterpret_cast<uintptr_t>(g.x) & 15) == 0 && g.ldx % 8 == 0 && g.ldw % 8 == 0 &&
                     (reinterpret_cast<uintptr_t>(g.W) & 15) == 0,
                 "gemv: x/W must be 16B aligned with 16B-multiple row pitches");
    GemvParams p;
    p.x = reinterpret_cast<const __nv_bfloat16*>(g.x); p.ldx = g.ldx;
    p.W = reinterpret_cast<const __nv_bfloat16*>(g.W); p.ldw = g.ldw;
    p.gamma = reinterpret_cast<const __nv_bfloat16*>(g.norm_gamma); p.eps = g.eps;
    p.residual = reinterpret_cast<const __nv_bfloat16*>(g.residual); p.ld_res = g.ld_res;
    p.out = g.out; p.ld_out = g.ld_out; p.out_fp32 = g.out_fp32;
    p.B = g.B; p.N = g.N; p.K = g.K; p.act = g.act;
    if (g.B == 1) return launch_gemv<1>(p, stream);
    if (g.B == 2) return launch_gemv<2>(p, stream);
    if (g.B <= 4) return launch_gemv<4>(p, stream);
    return launch_gemv<8>(p, stream);
}

}  // namespace b2
