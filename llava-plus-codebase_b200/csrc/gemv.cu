// Weight-streaming GEMV for the one-token decode step outside the megakernel (batch 3..8, prefill's last-position
// lm_head, B2_DECODE_MEGA=0):
//     out[B, N] = (RMSNorm(x) | x)[B, K] · W[N, K]^T  (+ residual)      or the fused SwiGLU variant.
// HBM-bound: every weight byte is read exactly once with 16-byte coalesced,
// L1-bypassing loads from a persistent grid (one 512-thread CTA per SM, 16 loads in flight per lane through a
// register double buffer; the first loads are issued BEFORE the activation prologue since weights do not depend
// on the previous kernel's output).
// The arithmetic runs on the tensor cores: mma.sync.m16n8k16 (bf16 x bf16 -> fp32) with the activations as the
// A operand (rows = batch, zero padded to 16) and 8 weight rows as the B operand. One 16-byte load per lane
// (weight row g = lane/4, 8 consecutive k at (lane%4)*8) feeds TWO MMAs with no unpacking, because the k index
// is permuted identically on both operands (a dot product does not care). ~5 instructions per 16 B of weights for
// any batch <= 8, against ~40 (B=1) .. ~250 (B=8) for the scalar bf16->fp32 FMA loop this replaces, which ncu
// showed to be issue-bound (profiles/r1a_prof_gemv_ncu_full.txt).
// CTA c owns a contiguous range of output rows (SwiGLU: channels) that differs by at most one unit between CTAs,
// walked in 8-row blocks (last one partial: missing rows are not fetched); its 16 warps split K, the 16 partial
// sums per output are reduced through shared memory in a fixed order (deterministic).
// Reference math replaced: transformers modeling_llama.py:62-67 (RMSNorm), :182-184 (LlamaMLP),
// :251-289 (q/k/v/o projections), llava_llama.py:48 (lm_head).
#include "common.cuh"
#include "kernels.h"

namespace b2 {
namespace {

constexpr int GV_THREADS = 512;
constexpr int GV_WARPS = GV_THREADS / 32;
constexpr int GV_UB = 8;  // 16-byte loads per lane per pipeline batch

struct GemvParams {
    const __nv_bfloat16* x; int64_t ldx;
    const __nv_bfloat16* W; int ldw;
    const __nv_bfloat16* gamma; float eps;
    const __nv_bfloat16* residual; int ld_res;
    void* out; int ld_out; int out_fp32;
    int B, N, K, act;
    int nb_max;  // blocks per CTA bound used for the smem partial table
};

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
    f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}

__device__ __forceinline__ void gv_mma(float (&c)[4], uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
    // A rows 8..15 (a1, a3) are the zero padding of the batch dimension
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b0), "r"(b1));
}

struct Ctx {
    int u_lo, nu, nb;    // output units (rows / SwiGLU channels) of this CTA, blocks of 8 rows
    int ks_lo, ks_len;   // this warp's K slice in 32-element blocks
};

// physical weight row of lane-group g of local block rb; valid = the row belongs to this CTA's range
__device__ __forceinline__ int gv_phys_row(const GemvParams& p, const Ctx& c, int rb, int g, bool& valid) {
    if (p.act == ACT_SWIGLU) {  // block-64 interleaved gate/up: g<4 -> gate of local channel 4*rb+g, g>=4 -> its up row
        const int lc = rb * 4 + (g & 3);
        valid = lc < c.nu;
        const int ch = c.u_lo + lc;
        return (ch >> 6) * 128 + (ch & 63) + ((g >> 2) ? 64 : 0);
    }
    const int lr = rb * 8 + g;
    valid = lr < c.nu;
    return c.u_lo + lr;
}

// issue the loads of batch bt (units 8*bt ..; unit u = (block u / ks_len, k32 ks_lo + u % ks_len)); buf fully defined
__device__ __forceinline__ void gv_issue(const GemvParams& p, const Ctx& c, int bt, int lane, uint4 (&buf)[GV_UB]) {
    const int u0 = bt * GV_UB;
    const int U = c.nb * c.ks_len;
    if (u0 < U) {
        const int g = lane >> 2, t = lane & 3;
        int rb = u0 / c.ks_len;
        int kk = u0 - rb * c.ks_len;
        bool valid;
        int row = gv_phys_row(p, c, rb, g, valid);
        const __nv_bfloat16* wr = p.W + (size_t)row * p.ldw + (size_t)c.ks_lo * 32 + t * 8;
#pragma unroll
        for (int j = 0; j < GV_UB; ++j) {
            buf[j] = (u0 + j < U && valid) ? ld_stream_16(wr + kk * 32) : make_uint4(0, 0, 0, 0);
            if (++kk == c.ks_len) {
                kk = 0;
                ++rb;
                row = gv_phys_row(p, c, rb, g, valid);
                wr = p.W + (size_t)row * p.ldw + (size_t)c.ks_lo * 32 + t * 8;
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < GV_UB; ++j) buf[j] = make_uint4(0, 0, 0, 0);
    }
}

template <int NB>
__global__ void __launch_bounds__(GV_THREADS, 1) gemv_kernel(GemvParams p) {
    extern __shared__ __align__(16) uint8_t gv_smem[];
    __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(gv_smem);                               // [NB][K]
    float* s_part = reinterpret_cast<float*>(gv_smem + (size_t)NB * p.K * 2);                    // [16][nb_max][NB][8]
    __shared__ float s_red[GV_WARPS][NB];
    __shared__ float s_rstd[NB];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = p.K;
    Ctx c;
    {
        const long long units = p.act == ACT_SWIGLU ? (p.N >> 1) : p.N;
        c.u_lo = (int)((units * blockIdx.x) / gridDim.x);
        c.nu = (int)((units * (blockIdx.x + 1)) / gridDim.x) - c.u_lo;
        c.nb = p.act == ACT_SWIGLU ? (c.nu + 3) >> 2 : (c.nu + 7) >> 3;
        const int nk32 = K >> 5;
        c.ks_lo = (nk32 * warp) / GV_WARPS;
        c.ks_len = (nk32 * (warp + 1)) / GV_WARPS - c.ks_lo;
    }

    uint4 bufA[GV_UB], bufB[GV_UB];
    gv_issue(p, c, 0, lane, bufA);  // weights do not depend on x: get HBM requests in flight before the prologue
    gv_issue(p, c, 1, lane, bufB);

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
                float gf[8];
                unpack8(*reinterpret_cast<const uint4*>(p.gamma + i * 8), gf);
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    uint4* px = reinterpret_cast<uint4*>(xs + (size_t)b * K + i * 8);
                    float f[8], o[8];
                    unpack8(*px, f);
                    const float rstd = s_rstd[b];
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

    // ---------------- main loop: tensor-core dot products over this warp's K slice of every block ----------------
    {
        const int g = lane >> 2, t = lane & 3;
        const int U = c.nb * c.ks_len;
        const int n_batches = (U + GV_UB - 1) / GV_UB;
        float acc[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f};
        int rb = 0, kk = 0;
        auto compute = [&](int bt, const uint4 (&buf)[GV_UB]) {
            const int u0 = bt * GV_UB;
            if (u0 >= U) return;
#pragma unroll
            for (int j = 0; j < GV_UB; ++j) {
                if (u0 + j < U) {  // warp-uniform
                    uint4 xv = make_uint4(0, 0, 0, 0);
                    if (g < p.B) xv = *reinterpret_cast<const uint4*>(xs + (size_t)g * K + (size_t)(c.ks_lo + kk) * 32 + t * 8);
                    gv_mma(acc, xv.x, xv.y, buf[j].x, buf[j].y);
                    gv_mma(acc2, xv.z, xv.w, buf[j].z, buf[j].w);
                    if (++kk == c.ks_len) {
                        // acc[0..1] = D[batch g][weight rows 2t, 2t+1] of block rb over this warp's K slice
                        if (g < NB)
                            *reinterpret_cast<float2*>(s_part + (((size_t)warp * p.nb_max + rb) * NB + g) * 8 + t * 2) =
                                make_float2(acc[0] + acc2[0], acc[1] + acc2[1]);
                        acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
                        acc2[0] = acc2[1] = acc2[2] = acc2[3] = 0.f;
                        kk = 0;
                        ++rb;
                    }
                }
            }
        };
#pragma unroll 1
        for (int bt = 0; bt < n_batches; bt += 2) {
            compute(bt, bufA);
            gv_issue(p, c, bt + 2, lane, bufA);
            compute(bt + 1, bufB);
            gv_issue(p, c, bt + 3, lane, bufB);
        }
    }
    __syncthreads();

    // ---------------- reduce the 16 K slices (fixed order), epilogue, coalesced writes ----------------
    {
        const int nk32 = K >> 5;
        auto row_value = [&](int rbl, int b, int r) {
            float v = 0.f;
#pragma unroll
            for (int ww = 0; ww < GV_WARPS; ++ww) {
                const bool has = (nk32 * (ww + 1)) / GV_WARPS > (nk32 * ww) / GV_WARPS;  // K < 512: empty slices
                if (has) v += s_part[(((size_t)ww * p.nb_max + rbl) * NB + b) * 8 + r];
            }
            return v;
        };
        if (p.act == ACT_SWIGLU) {
            for (int idx = tid; idx < c.nu * p.B; idx += GV_THREADS) {
                const int b = idx / c.nu, r = idx - b * c.nu;
                const float gt = row_value(r >> 2, b, r & 3);
                const float up = row_value(r >> 2, b, (r & 3) + 4);
                reinterpret_cast<__nv_bfloat16*>(p.out)[(size_t)b * p.ld_out + c.u_lo + r] =
                    __float2bfloat16_rn(gt / (1.0f + __expf(-gt)) * up);
            }
        } else {
            for (int idx = tid; idx < c.nu * p.B; idx += GV_THREADS) {
                const int b = idx / c.nu, r = idx - b * c.nu;
                float y = row_value(r >> 3, b, r & 7);
                const int row = c.u_lo + r;
                if (p.residual != nullptr) y += __bfloat162float(p.residual[(size_t)b * p.ld_res + row]);
                if (p.out_fp32) reinterpret_cast<float*>(p.out)[(size_t)b * p.ld_out + row] = y;
                else reinterpret_cast<__nv_bfloat16*>(p.out)[(size_t)b * p.ld_out + row] = __float2bfloat16_rn(y);
            }
        }
    }
}

template <int NB>
int launch_gemv(GemvParams p, cudaStream_t stream) {
    const int grid = num_sms();
    const long long units = p.act == ACT_SWIGLU ? (p.N >> 1) : p.N;
    const int nu_max = (int)((units + grid - 1) / grid);
    p.nb_max = p.act == ACT_SWIGLU ? (nu_max + 3) / 4 : (nu_max + 7) / 8;
    const size_t smem = (size_t)NB * p.K * 2 + (size_t)GV_WARPS * p.nb_max * NB * 8 * 4;
    B2_CHECK_ARG(smem <= 226 * 1024, "gemv: activation tile + partial table do not fit shared memory (B=%d K=%d N=%d)",
                 p.B, p.K, p.N);
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        const size_t want = smem > 48 * 1024 ? smem : 48 * 1024;
        B2_CUDA_CHECK(cudaFuncSetAttribute(gemv_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want));
        attr_smem = want;
    }
    gemv_kernel<NB><<<grid, GV_THREADS, smem, stream>>>(p);
    B2_LAUNCH_CHECK();
    return 0;
}

size_t gemv_smem_bytes(int B, int N, int K, int act) {
    const int NB = B == 1 ? 1 : (B == 2 ? 2 : (B <= 4 ? 4 : 8));
    const int grid = num_sms();
    const long long units = act == ACT_SWIGLU ? (N >> 1) : N;
    const int nu_max = (int)((units + grid - 1) / grid);
    const int nb_max = act == ACT_SWIGLU ? (nu_max + 3) / 4 : (nu_max + 7) / 8;
    return (size_t)NB * K * 2 + (size_t)GV_WARPS * nb_max * NB * 8 * 4;
}

}  // namespace

bool gemv_fits(int B, int N, int K, int act) { return B >= 1 && B <= 8 && gemv_smem_bytes(B, N, K, act) <= 226 * 1024; }

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
    p.B = g.B; p.N = g.N; p.K = g.K; p.act = g.act; p.nb_max = 0;
    if (g.B == 1) return launch_gemv<1>(p, stream);
    if (g.B == 2) return launch_gemv<2>(p, stream);
    if (g.B <= 4) return launch_gemv<4>(p, stream);
    return launch_gemv<8>(p, stream);
}

}  // namespace b2
