// Persistent decode-step megakernel (batch <= 8): ONE launch per generated token.
//
// The one-token decode step of LLaVA/LLaMA (reference: the HF one-token forward behind
// llava/model/llava_arch.py:103-112; transformers modeling_llama.py:303-332 per layer) is a pure weight
// stream: 13.2 GB (7B) read once per token. Run as separate kernels, every Linear pays a launch gap, a
// pipeline ramp and a tail; here one cooperative grid (one 512-thread CTA per SM) walks all phases
//     embed -> L x { QKV gemv (+RMSNorm) | RoPE + KV append + split-KV attention | o_proj gemv (+res) |
//                    gate/up gemv (+RMSNorm, SwiGLU) | down gemv (+res) } -> lm_head gemv (+RMSNorm) -> argmax
// separated by grid-wide barriers, and — because weights do not depend on activations — every warp issues the
// first 16-byte weight loads of the NEXT phase before it arrives at the barrier, so the HBM pipe stays busy
// across the dependency. Weights use L1-bypassing non-coherent loads; activations that cross a barrier are read
// with ld.global.cg (L2) because L1 is not coherent between SMs.
#include <limits.h>
#include <math.h>

#include "common.cuh"
#include "kernels.h"

namespace b2 {
namespace {

constexpr int MK_THREADS = 512;
constexpr int MK_WARPS = MK_THREADS / 32;
constexpr int MK_U = 4;  // 256-element K chunks per pipeline step (x 2 rows = 8 loads in flight per lane)
constexpr int MK_D = 128;

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
    f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 ldcg16(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }

// Grid barrier (all CTAs co-resident: cooperative launch, grid = #SMs). One monotonically increasing counter:
// every CTA does a fire-and-forget red.release (+1) and polls with ld.acquire until the counter reaches the
// target the host-provided launch sequence number implies — one L2 round trip on the critical path instead of
// the four (load generation, fence, atomic, poll) of a sense-reversing barrier. Signed difference compare
// survives 32-bit wrap-around.
__device__ __forceinline__ void grid_sync(unsigned int* counter, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        unsigned int cur, spins = 0;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(cur) : "l"(counter) : "memory");
            if (++spins > (1u << 26)) asm volatile("trap;");  // protocol bug -> CUDA error, not a hang
        } while ((int)(cur - target) < 0);
    }
    __syncthreads();
}

// hot per-phase state (lives in registers through the weight-streaming loop).
// Work decomposition of one GEMV phase:
//   * output rows are grouped into logical 8-row blocks (SwiGLU: 4 gate rows + the 4 up rows of the same
//     channels); CTA c owns the contiguous block range [rb_lo, rb_lo + nb);
//   * inside the CTA the 16 warps split K (in 32-element slices), so every warp streams exactly the same number
//     of bytes for every block, and the 16 partial sums per output are reduced through shared memory in a fixed
//     order (deterministic);
//   * the arithmetic runs on the tensor cores as mma.sync.m16n8k16 (bf16 x bf16 -> fp32) with the activations
//     as the A operand (rows = batch, zero padded to 16) and 8 weight rows as the B operand. One 16-byte load per
//     lane (row g = lane/4, 8 consecutive k at (lane%4)*8) feeds TWO MMAs with no unpacking: the k index is
//     permuted consistently on both operands, which a dot product does not care about. ~5 instructions per
//     16 B of weights instead of ~40 for the scalar bf16->fp32 FMA loop (the round-1 kernel was issue-bound).
struct GemvCtx {
    const __nv_bfloat16* W;
    int K, act;
    int rb_lo, nb;       // logical 8-row blocks of this CTA
    int ks_lo, ks_len;   // this warp's K slice, in 32-element blocks
};
// cold per-phase I/O, recomputed from the phase index where needed (prologue, epilogue)
struct PhaseIO {
    const __nv_bfloat16* xin;
    const __nv_bfloat16* gamma;
    const __nv_bfloat16* residual;
    void* out;
    int ld_out, out_fp32;
};

struct WarpId { int tid, lane, warp, gw, total_warps; };
constexpr int MK_MAXL = 48;  // decoder layers whose weight-pointer table is cached in shared memory
__device__ __forceinline__ PhaseIO mk_phase_io(const MegaParams& p, const MegaLayer* layers, int ph);

constexpr int MK_UB = 8;      // units (16 B loads per lane) per pipeline batch
constexpr int MK_MAXNB = 32;  // max 8-row blocks per CTA per phase (host-checked)

// physical weight row streamed by lane-group g (0..7) of logical block `blk`
__device__ __forceinline__ int mk_phys_row(int act, int blk, int g) {
    if (act == ACT_SWIGLU) {  // block-64 interleaved gate/up: g<4 -> gate of channel 4*blk+g, g>=4 -> up of 4*blk+g-4
        const int ch = blk * 4 + (g & 3);
        return (ch >> 6) * 128 + (ch & 63) + ((g >> 2) ? 64 : 0);
    }
    return blk * 8 + g;
}

// issue the loads of batch `bt` (units 8*bt ...; unit u = (block u / ks_len, k-slice element u % ks_len)) —
// a full definition of buf on every path
__device__ __forceinline__ void mk_issue(const GemvCtx& c, int bt, const WarpId& w, uint4 (&buf)[MK_UB]) {
    const int u0 = bt * MK_UB;
    const int U = c.nb * c.ks_len;
    if (u0 < U) {
        const int g = w.lane >> 2, t = w.lane & 3;
        int rb = u0 / c.ks_len;
        int kk = u0 - rb * c.ks_len;
        const __nv_bfloat16* wr = c.W + (size_t)mk_phys_row(c.act, c.rb_lo + rb, g) * c.K + (size_t)c.ks_lo * 32 + t * 8;
#pragma unroll
        for (int j = 0; j < MK_UB; ++j) {
            buf[j] = (u0 + j < U) ? ld_stream_16(wr + kk * 32) : make_uint4(0, 0, 0, 0);
            if (++kk == c.ks_len) {
                kk = 0;
                ++rb;
                wr = c.W + (size_t)mk_phys_row(c.act, c.rb_lo + rb, g) * c.K + (size_t)c.ks_lo * 32 + t * 8;
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < MK_UB; ++j) buf[j] = make_uint4(0, 0, 0, 0);
    }
}

// x -> smem (bf16), optional RMSNorm (HF semantics: gamma * bf16(x * rstd))
template <int NB>
__device__ __forceinline__ void mk_prologue(const PhaseIO& c, int K, int B, float eps, const WarpId& w,
                                            __nv_bfloat16* xs, float (*s_red)[NB], float* s_rstd) {
    const int nvec = K >> 3;
    float ss[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) ss[b] = 0.f;
    for (int i = w.tid; i < nvec; i += MK_THREADS) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            uint4 u = make_uint4(0, 0, 0, 0);
            if (b < B) u = ldcg16(c.xin + (size_t)b * K + i * 8);
            *reinterpret_cast<uint4*>(xs + (size_t)b * K + i * 8) = u;
            if (c.gamma != nullptr) {
                float f[8];
                unpack8(u, f);
#pragma unroll
                for (int e = 0; e < 8; ++e) ss[b] += f[e] * f[e];
            }
        }
    }
    if (c.gamma != nullptr) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const float v = warp_sum(ss[b]);
            if (w.lane == 0) s_red[w.warp][b] = v;
        }
        __syncthreads();
        if (w.tid < NB) {
            float t = 0.f;
            for (int i = 0; i < MK_WARPS; ++i) t += s_red[i][w.tid];
            s_rstd[w.tid] = rsqrtf(t / K + eps);
        }
        __syncthreads();
        for (int i = w.tid; i < nvec; i += MK_THREADS) {
            float gf[8];
            unpack8(*reinterpret_cast<const uint4*>(c.gamma + i * 8), gf);
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                uint4* px = reinterpret_cast<uint4*>(xs + (size_t)b * K + i * 8);
                float f[8], o[8];
                unpack8(*px, f);
                const float rstd = s_rstd[b];
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = gf[e] * round_bf16(f[e] * rstd);
                *px = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]),
                                 pack_bf16(o[6], o[7]));
            }
        }
    }
    __syncthreads();
}

// running state of a warp inside a phase
struct RowState { int rb, kk; };

__device__ __forceinline__ void mk_mma(float (&c)[4], uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
    // A rows 8..15 (a1, a3) are the zero padding of the batch dimension
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b0), "r"(b1));
}

// s_part layout: [warp][block][b][8 rows] fp32
template <int NB>
__device__ __forceinline__ void mk_compute(const GemvCtx& c, int bt, int B, const WarpId& w,
                                           const uint4 (&buf)[MK_UB], const __nv_bfloat16* xs, RowState& st,
                                           float (&acc)[4], float* s_part) {
    const int u0 = bt * MK_UB;
    const int U = c.nb * c.ks_len;
    if (u0 >= U) return;
    const int g = w.lane >> 2, t = w.lane & 3;
#pragma unroll
    for (int j = 0; j < MK_UB; ++j) {
        if (u0 + j < U) {  // warp-uniform
            uint4 xv = make_uint4(0, 0, 0, 0);
            if (g < B) xv = *reinterpret_cast<const uint4*>(xs + (size_t)g * c.K + (size_t)(c.ks_lo + st.kk) * 32 + t * 8);
            mk_mma(acc, xv.x, xv.y, buf[j].x, buf[j].y);
            mk_mma(acc, xv.z, xv.w, buf[j].z, buf[j].w);
            if (++st.kk == c.ks_len) {
                // acc[0], acc[1] = D[batch g][weight rows 2t, 2t+1] over this warp's K slice
                if (g < NB) {
                    float2* dst = reinterpret_cast<float2*>(s_part + (((size_t)w.warp * MK_MAXNB + st.rb) * NB + g) * 8 + t * 2);
                    *dst = make_float2(acc[0], acc[1]);
                }
                acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
                st.kk = 0;
                ++st.rb;
            }
        }
    }
}

template <int NB>
__device__ __forceinline__ float mk_row_value(const float* s_part, int nk32, int rb, int b, int r) {
    float v = 0.f;
#pragma unroll
    for (int ww = 0; ww < MK_WARPS; ++ww) {  // fixed order; warps whose K slice is empty (K < 512) wrote nothing
        const bool has = (nk32 * (ww + 1)) / MK_WARPS > (nk32 * ww) / MK_WARPS;
        if (has) v += s_part[(((size_t)ww * MK_MAXNB + rb) * NB + b) * 8 + r];
    }
    return v;
}

// after the streaming loop: reduce the 16 K-slices, apply the epilogue, coalesced global writes
template <int NB>
__device__ __forceinline__ void mk_epilogue(const MegaParams& p, const MegaLayer* layers, int ph, const GemvCtx& c,
                                            int B, const WarpId& w, const float* s_part) {
    const PhaseIO io = mk_phase_io(p, layers, ph);
    if (c.act == ACT_SWIGLU) {
        const int nch = c.nb * 4;
        for (int idx = w.tid; idx < nch * B; idx += MK_THREADS) {
            const int b = idx / nch, r = idx - b * nch;
            const int rb = r >> 2, q = r & 3;
            const float gt = mk_row_value<NB>(s_part, c.K >> 5, rb, b, q);
            const float up = mk_row_value<NB>(s_part, c.K >> 5, rb, b, q + 4);
            reinterpret_cast<__nv_bfloat16*>(io.out)[(size_t)b * io.ld_out + (size_t)c.rb_lo * 4 + r] =
                __float2bfloat16_rn(gt / (1.0f + __expf(-gt)) * up);
        }
    } else {
        const int nr = c.nb * 8;
        for (int idx = w.tid; idx < nr * B; idx += MK_THREADS) {
            const int b = idx / nr, r = idx - b * nr;
            float y = mk_row_value<NB>(s_part, c.K >> 5, r >> 3, b, r & 7);
            const size_t o = (size_t)b * io.ld_out + (size_t)c.rb_lo * 8 + r;
            if (io.residual != nullptr) y += __bfloat162float(__ldcg(io.residual + o));
            if (io.out_fp32) reinterpret_cast<float*>(io.out)[o] = y;
            else reinterpret_cast<__nv_bfloat16*>(io.out)[o] = __float2bfloat16_rn(y);
        }
    }
}

// RoPE + cache append + split-KV attention.
// (b, head) pairs are spread over the grid: with pairs <= #CTAs, G = #CTAs / pairs CTAs share one pair (split-KV
// across CTAs, merged by the last CTA to arrive — only G partials, G = 4 at B = 1); otherwise each CTA walks
// pairs one after the other. Inside a CTA the 16 warps split the key range (half-warp per 256 B K/V row,
// online softmax in registers) and merge through shared memory, so no long serial merge sits on the critical path.
__device__ __forceinline__ void mk_attention(const MegaParams& p, const MegaLayer& Lw, const WarpId& w,
                                             float (*s_part)[MK_D + 2], int* s_flag, const float* s_cos,
                                             const float* s_sin) {
    const int h = p.h, H = p.H;
    const int pairs = p.B * H;
    const int grid = gridDim.x;
    const int G = pairs <= grid ? grid / pairs : 1;
    const int lane = w.lane;
    const int hw = lane >> 4, c = lane & 15;
    const int split = pairs <= grid ? (int)blockIdx.x % G : 0;
    for (int pair = pairs <= grid ? (int)blockIdx.x / G : (int)blockIdx.x; pair < pairs;
         pair += (pairs <= grid ? pairs : grid)) {  // CTA-uniform loop
        const int head = pair % H, b = pair / H;
        const int pos = p.cur_len[b];
        const int total = pos + 1;
        const int cchunk = (total + G - 1) / G;
        const int cb = split * cchunk;
        const int ce = min(cb + cchunk, total);
        const int wchunk = (max(ce - cb, 0) + MK_WARPS - 1) / MK_WARPS;
        const int k_begin = cb + w.warp * wchunk;
        const int k_end = min(k_begin + wchunk, ce);
        const __nv_bfloat16* qrow = p.qkv + (size_t)b * 3 * h + head * MK_D;
        float qreg[8], knew[8], vnew[8];
        {
            // RoPE: element i pairs with i+64 -> chunk c pairs with chunk c^8
            float qa[8], qb[8], ka[8], kb[8];
            unpack8(ldcg16(qrow + c * 8), qa);
            unpack8(ldcg16(qrow + (c ^ 8) * 8), qb);
            unpack8(ldcg16(qrow + h + c * 8), ka);
            unpack8(ldcg16(qrow + h + (c ^ 8) * 8), kb);
            unpack8(ldcg16(qrow + 2 * h + c * 8), vnew);
            const float sign = (c < 8) ? -1.f : 1.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int i = ((c & 7) * 8 + e);  // frequency index 0..63; cos/sin(pos_b * inv_freq_i) tabulated once per step
                const float cbf = s_cos[b * 64 + i], sbf = s_sin[b * 64 + i];
                qreg[e] = round_bf16(round_bf16(qa[e] * cbf) + round_bf16(sign * qb[e] * sbf));
                knew[e] = round_bf16(round_bf16(ka[e] * cbf) + round_bf16(sign * kb[e] * sbf));
            }
        }
        const size_t cbase = ((size_t)b * H + head) * p.Smax * MK_D;
        if (pos >= k_begin && pos < k_end && hw == 0) {  // append (one half-warp writes the 256 B rows)
            *reinterpret_cast<uint4*>(Lw.kcache + cbase + (size_t)pos * MK_D + c * 8) =
                make_uint4(pack_bf16(knew[0], knew[1]), pack_bf16(knew[2], knew[3]),
                           pack_bf16(knew[4], knew[5]), pack_bf16(knew[6], knew[7]));
            *reinterpret_cast<uint4*>(Lw.vcache + cbase + (size_t)pos * MK_D + c * 8) =
                make_uint4(pack_bf16(vnew[0], vnew[1]), pack_bf16(vnew[2], vnew[3]),
                           pack_bf16(vnew[4], vnew[5]), pack_bf16(vnew[6], vnew[7]));
        }
        float m_run = -INFINITY, l_run = 0.f, av[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) av[e] = 0.f;
        const __nv_bfloat16* kb_ = Lw.kcache + cbase + c * 8;
        const __nv_bfloat16* vb_ = Lw.vcache + cbase + c * 8;
        for (int kbase = k_begin; kbase < k_end; kbase += 2 * MK_U) {  // warp-uniform trip count
            uint4 kraw[MK_U], vraw[MK_U];
#pragma unroll
            for (int u = 0; u < MK_U; ++u) {
                const int key = kbase + hw + 2 * u;
                if (key < k_end && key != pos) {
                    kraw[u] = ld_stream_16(kb_ + (size_t)key * MK_D);
                    vraw[u] = ld_stream_16(vb_ + (size_t)key * MK_D);
                } else {
                    kraw[u] = make_uint4(0, 0, 0, 0);
                    vraw[u] = make_uint4(0, 0, 0, 0);
                }
            }
#pragma unroll
            for (int u = 0; u < MK_U; ++u) {
                const int key = kbase + hw + 2 * u;
                float kf[8], vf[8];
                unpack8(kraw[u], kf);
                unpack8(vraw[u], vf);
                if (key == pos) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) { kf[e] = knew[e]; vf[e] = vnew[e]; }
                }
                float dot = 0.f;
#pragma unroll
                for (int e = 0; e < 8; ++e) dot += qreg[e] * kf[e];
                dot += __shfl_xor_sync(0xffffffffu, dot, 8);
                dot += __shfl_xor_sync(0xffffffffu, dot, 4);
                dot += __shfl_xor_sync(0xffffffffu, dot, 2);
                dot += __shfl_xor_sync(0xffffffffu, dot, 1);
                if (key < k_end) {
                    const float sc = dot * p.scale_log2;
                    const float m_new = fmaxf(m_run, sc);
                    const float corr = exp2f(m_run - m_new);
                    const float pr = exp2f(sc - m_new);
                    l_run = l_run * corr + pr;
#pragma unroll
                    for (int e = 0; e < 8; ++e) av[e] = av[e] * corr + pr * vf[e];
                    m_run = m_new;
                }
            }
        }
        {   // merge the two half-warps
            const float m_o = __shfl_xor_sync(0xffffffffu, m_run, 16);
            const float l_o = __shfl_xor_sync(0xffffffffu, l_run, 16);
            const float m_c = fmaxf(m_run, m_o);
            const float w_s = (m_run == -INFINITY) ? 0.f : exp2f(m_run - m_c);
            const float w_o = (m_o == -INFINITY) ? 0.f : exp2f(m_o - m_c);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float a_o = __shfl_xor_sync(0xffffffffu, av[e], 16);
                av[e] = av[e] * w_s + a_o * w_o;
            }
            l_run = l_run * w_s + l_o * w_o;
            m_run = m_c;
        }
        if (hw == 0) {
#pragma unroll
            for (int e = 0; e < 8; ++e) s_part[w.warp][c * 8 + e] = av[e];
            if (c == 0) { s_part[w.warp][MK_D] = m_run; s_part[w.warp][MK_D + 1] = l_run; }
        }
        __syncthreads();
        // ---- merge the 16 warps of this CTA (thread d owns output element d) ----
        float m_cta = -INFINITY, l_cta = 0.f, o_cta = 0.f;
        if (w.tid < MK_D) {
#pragma unroll
            for (int i = 0; i < MK_WARPS; ++i) m_cta = fmaxf(m_cta, s_part[i][MK_D]);
#pragma unroll
            for (int i = 0; i < MK_WARPS; ++i) {
                const float ms = s_part[i][MK_D];
                const float wgt = (ms == -INFINITY) ? 0.f : exp2f(ms - m_cta);
                l_cta += s_part[i][MK_D + 1] * wgt;
                o_cta += s_part[i][w.tid] * wgt;
            }
        }
        if (G == 1) {
            if (w.tid < MK_D)
                p.attn[(size_t)b * h + head * MK_D + w.tid] = __float2bfloat16_rn(o_cta / l_cta);
        } else {
            float* part = p.attn_partial + ((size_t)pair * G + split) * (MK_D + 2);
            if (w.tid < MK_D) {
                part[w.tid] = o_cta;
                if (w.tid == 0) { part[MK_D] = m_cta; part[MK_D + 1] = l_cta; }
            }
            __threadfence();
            __syncthreads();
            if (w.tid == 0) *s_flag = (atomicAdd(&p.attn_counters[pair], 1) == G - 1) ? 1 : 0;
            __syncthreads();
            if (*s_flag) {  // last CTA of this (b, head): merge the G partials
                __threadfence();
                if (w.tid < MK_D) {
                    const float* pb = p.attn_partial + (size_t)pair * G * (MK_D + 2);
                    float m_all = -INFINITY;
                    for (int sidx = 0; sidx < G; ++sidx) m_all = fmaxf(m_all, __ldcg(pb + (size_t)sidx * (MK_D + 2) + MK_D));
                    float l_all = 0.f, o_all = 0.f;
                    for (int sidx = 0; sidx < G; ++sidx) {
                        const float ms = __ldcg(pb + (size_t)sidx * (MK_D + 2) + MK_D);
                        const float wgt = (ms == -INFINITY) ? 0.f : exp2f(ms - m_all);
                        l_all += __ldcg(pb + (size_t)sidx * (MK_D + 2) + MK_D + 1) * wgt;
                        o_all += __ldcg(pb + (size_t)sidx * (MK_D + 2) + w.tid) * wgt;
                    }
                    p.attn[(size_t)b * h + head * MK_D + w.tid] = __float2bfloat16_rn(o_all / l_all);
                    if (w.tid == 0) p.attn_counters[pair] = 0;
                }
            }
        }
        __syncthreads();  // s_part / s_flag are reused by the next pair
    }
}

// phase k of layer l (k: 0 = QKV, 1 = attention, 2 = o_proj, 3 = gate/up, 4 = down); index 5L = lm_head
__device__ __forceinline__ PhaseIO mk_phase_io(const MegaParams& p, const MegaLayer* layers, int ph) {
    PhaseIO c;
    const int l = ph / 5, k = ph % 5;
    c.gamma = nullptr; c.residual = nullptr; c.out_fp32 = 0;
    if (l >= p.L) { c.xin = p.x; c.gamma = p.final_norm; c.out = p.logits; c.ld_out = p.V; c.out_fp32 = 1; }
    else if (k <= 1) { c.xin = p.x; c.gamma = layers[l].ln1; c.out = p.qkv; c.ld_out = 3 * p.h; }
    else if (k == 2) { c.xin = p.attn; c.residual = p.x; c.out = p.x; c.ld_out = p.h; }
    else if (k == 3) { c.xin = p.x; c.gamma = layers[l].ln2; c.out = p.act; c.ld_out = p.I; }
    else { c.xin = p.act; c.residual = p.x; c.out = p.x; c.ld_out = p.h; }
    return c;
}
__device__ __forceinline__ GemvCtx mk_phase_ctx(const MegaParams& p, const MegaLayer* layers, int ph, const WarpId& w) {
    GemvCtx c;
    const int l = ph / 5, k = ph % 5;
    int N;
    c.act = ACT_NONE;
    if (l >= p.L) { c.W = p.lm_head; N = p.V; c.K = p.h; }
    else if (k <= 1) { c.W = layers[l].wqkv; N = 3 * p.h; c.K = p.h; }
    else if (k == 2) { c.W = layers[l].wo; N = p.h; c.K = p.h; }
    else if (k == 3) { c.W = layers[l].wgu; N = 2 * p.I; c.K = p.h; c.act = ACT_SWIGLU; }
    else { c.W = layers[l].wd; N = p.h; c.K = p.I; }
    const long long nblk = N >> 3;
    c.rb_lo = (int)((nblk * blockIdx.x) / gridDim.x);
    c.nb = (int)((nblk * (blockIdx.x + 1)) / gridDim.x) - c.rb_lo;
    const int nk32 = c.K >> 5;
    c.ks_lo = (nk32 * w.warp) / MK_WARPS;
    c.ks_len = (nk32 * (w.warp + 1)) / MK_WARPS - c.ks_lo;
    return c;
}

template <int NB>
__global__ void __launch_bounds__(MK_THREADS, 1) decode_mega_kernel(MegaParams p) {
    extern __shared__ __align__(16) uint8_t mk_smem[];
    __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(mk_smem);  // [NB][Kmax]
    float* s_gpart = reinterpret_cast<float*>(mk_smem + (size_t)NB * (p.h > p.I ? p.h : p.I) * 2);  // [16][MAXNB][NB][8]
    __shared__ float s_red[MK_WARPS][NB];
    __shared__ float s_rstd[NB];
    __shared__ float s_av[MK_WARPS];
    __shared__ int s_ai[MK_WARPS];
    __shared__ float s_part[MK_WARPS][MK_D + 2];
    __shared__ int s_flag;
    __shared__ float s_cos[NB * 64], s_sin[NB * 64];
    __shared__ MegaLayer s_layers[MK_MAXL];  // weight-pointer table: no dependent global load per phase  // RoPE table of this step (HF: cos/sin cast to bf16)

    WarpId w;
    w.tid = threadIdx.x; w.lane = w.tid & 31; w.warp = w.tid >> 5;
    w.gw = blockIdx.x * MK_WARPS + w.warp;
    w.total_warps = gridDim.x * MK_WARPS;
    const int B = p.B;

    uint4 bufA[MK_UB], bufB[MK_UB];
    float acc[4];

    {
        const uint4* src = reinterpret_cast<const uint4*>(p.layers);
        uint4* dst = reinterpret_cast<uint4*>(s_layers);
        for (int i = w.tid; i < p.L * (int)(sizeof(MegaLayer) / 16); i += MK_THREADS) dst[i] = src[i];
        __syncthreads();
    }

    // ---------------- phase "-1": x = embed_tokens[tok]; layer-0 QKV weights already in flight ----------------
    GemvCtx cur = mk_phase_ctx(p, s_layers, 0, w);
    mk_issue(cur, 0, w, bufA);
    mk_issue(cur, 1, w, bufB);
    for (int i = w.tid; i < B * 64; i += MK_THREADS) {
        const int b = i >> 6, f = i & 63;
        const float inv_freq = exp2f(-(2.0f * f / MK_D) * log2f(p.theta));
        float sv, cv;
        sincosf((float)p.cur_len[b] * inv_freq, &sv, &cv);
        s_cos[i] = round_bf16(cv);
        s_sin[i] = round_bf16(sv);
    }
    if (blockIdx.x < B) {
        int t = p.tok[blockIdx.x];
        t = t < 0 ? 0 : (t >= p.V ? p.V - 1 : t);
        const uint4* src = reinterpret_cast<const uint4*>(p.embed + (size_t)t * p.h);
        uint4* dst = reinterpret_cast<uint4*>(p.x + (size_t)blockIdx.x * p.h);
        for (int i = w.tid; i < p.h / 8; i += MK_THREADS) dst[i] = src[i];
    }
    unsigned int bar_target = p.bar_base + gridDim.x;
    grid_sync(p.bar_count, bar_target);

    const int n_phases = 5 * p.L + 1;
    const bool tracing = p.trace != nullptr && blockIdx.x == 0 && w.tid == 0;
#pragma unroll 1
    for (int ph = 0; ph < n_phases; ++ph) {
        if (tracing) p.trace[ph * 4 + 0] = clock64();
        if (ph % 5 == 1 && ph < 5 * p.L) {
            mk_attention(p, s_layers[ph / 5], w, s_part, &s_flag, s_cos, s_sin);
        } else {
            // the first step's weights were issued (into bufA) before the preceding barrier
            mk_prologue<NB>(mk_phase_io(p, s_layers, ph), cur.K, B, p.eps, w, xs, s_red, s_rstd);
            if (tracing) p.trace[ph * 4 + 1] = clock64();
            RowState st;
            st.rb = 0;
            st.kk = 0;
            acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
            const int n_batches = (cur.nb * cur.ks_len + MK_UB - 1) / MK_UB;
            // batches 0 and 1 were issued (bufA, bufB) before the preceding barrier
#pragma unroll 1
            for (int bt = 0; bt < n_batches; bt += 2) {
                mk_compute<NB>(cur, bt, B, w, bufA, xs, st, acc, s_gpart);
                mk_issue(cur, bt + 2, w, bufA);
                mk_compute<NB>(cur, bt + 1, B, w, bufB, xs, st, acc, s_gpart);
                mk_issue(cur, bt + 3, w, bufB);
            }
            __syncthreads();
            if (tracing) p.trace[ph * 4 + 2] = clock64();
            mk_epilogue<NB>(p, s_layers, ph, cur, B, w, s_gpart);
        }
        // prefetch the next GEMV phase's first weights across the barrier (weights don't depend on activations)
        int nxt = ph + 1;
        if (nxt % 5 == 1 && nxt < 5 * p.L) nxt = -1;  // attention follows: no weights to prefetch yet
        if (nxt >= 0 && nxt < n_phases) {
            cur = mk_phase_ctx(p, s_layers, nxt, w);
            mk_issue(cur, 0, w, bufA);
            mk_issue(cur, 1, w, bufB);
        } else {
            cur.nb = 0;
            mk_issue(cur, 0, w, bufA);  // defines the buffers (zeros): nothing is carried across the attention phase
            mk_issue(cur, 1, w, bufB);
        }
        if (tracing) p.trace[ph * 4 + 3] = clock64();
        bar_target += gridDim.x;
        grid_sync(p.bar_count, bar_target);
    }
    if (tracing) p.trace[n_phases * 4] = clock64();

    // ---------------- greedy argmax (first occurrence), token store, counters ----------------
    if (blockIdx.x < B) {
        const int b = blockIdx.x;
        const int tid = w.tid, lane = w.lane, warp = w.warp;
        const float* row = p.logits + (size_t)b * p.V;
        float best = -INFINITY;
        int bi = INT_MAX;
        for (int i = tid; i < p.V; i += MK_THREADS) {
            const float v = __ldcg(row + i);
            if (v == v && (bi == INT_MAX || v > best)) { best = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) { s_av[warp] = best; s_ai[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            best = lane < MK_WARPS ? s_av[lane] : -INFINITY;
            bi = lane < MK_WARPS ? s_ai[lane] : INT_MAX;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (lane == 0) {
                const int t = bi == INT_MAX ? 0 : bi;
                p.tok[b] = t;
                p.out_tokens[(size_t)(*p.step_counter) * B + b] = t;
                p.cur_len[b] += 1;  // every attention phase of this launch is behind the last barrier
            }
        }
    }
    // step counter: bumped by the last CTA to get here (after all token stores read it)
    __syncthreads();
    if (w.tid == 0) {
        __threadfence();
        if (atomicAdd(p.done_count, 1u) == gridDim.x - 1) {
            *p.done_count = 0u;
            *p.step_counter += 1;
        }
    }
}

}  // namespace

// dynamic smem: activation tile [NB][max(h, I)] bf16 + K-slice partial sums [16][MK_MAXNB][NB][8] fp32
static size_t mega_smem_bytes(int NB, int h, int I) {
    return (size_t)NB * (h > I ? h : I) * 2 + (size_t)MK_WARPS * MK_MAXNB * NB * 8 * 4;
}
bool decode_mega_fits(int B, int h, int I) {
    if (B < 1 || B > 8) return false;
    const int NB = B == 1 ? 1 : (B == 2 ? 2 : (B <= 4 ? 4 : 8));
    const size_t stat = (size_t)(MK_WARPS + 1) * NB * 4 + MK_WARPS * (MK_D + 2) * 4 + NB * 512 + MK_MAXL * sizeof(MegaLayer) + 1024;
    return mega_smem_bytes(NB, h, I) + stat <= 227 * 1024;
}

int decode_mega(const MegaParams& p, cudaStream_t stream) {
    B2_CHECK_ARG(p.B >= 1 && p.B <= 8, "decode_mega: batch must be 1..8");
    B2_CHECK_ARG(p.L <= MK_MAXL, "decode_mega: %d layers exceed the shared layer table (%d)", p.L, MK_MAXL);
    B2_CHECK_ARG(p.h % 256 == 0 && p.I % 256 == 0 && p.V % 2 == 0 && p.h / p.H == MK_D,
                 "decode_mega: unsupported dims h=%d I=%d V=%d H=%d", p.h, p.I, p.V, p.H);
    {
        const int grid = num_sms();
        const int nmax = p.V > 3 * p.h ? (p.V > 2 * p.I ? p.V : 2 * p.I) : (3 * p.h > 2 * p.I ? 3 * p.h : 2 * p.I);
        B2_CHECK_ARG((nmax / 8 + grid - 1) / grid + 1 <= MK_MAXNB && p.V % 8 == 0,
                     "decode_mega: %d output rows per CTA exceed the shared-memory block table", nmax / grid);
    }
    const int NB = p.B == 1 ? 1 : (p.B == 2 ? 2 : (p.B <= 4 ? 4 : 8));
    const int kmax = p.h > p.I ? p.h : p.I;
    const size_t smem = mega_smem_bytes(NB, p.h, p.I);
    B2_CHECK_ARG(decode_mega_fits(p.B, p.h, p.I), "decode_mega: activations do not fit shared memory (B=%d K=%d)",
                 p.B, kmax);
    void* fn = nullptr;
    switch (NB) {
        case 1: fn = (void*)decode_mega_kernel<1>; break;
        case 2: fn = (void*)decode_mega_kernel<2>; break;
        case 4: fn = (void*)decode_mega_kernel<4>; break;
        default: fn = (void*)decode_mega_kernel<8>; break;
    }
    static size_t attr_smem[9] = {0};
    if (smem > attr_smem[NB]) {
        const size_t want = smem > 48 * 1024 ? smem : 48 * 1024;
        B2_CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want));
        attr_smem[NB] = want;
    }
    MegaParams pp = p;
    void* args[] = {&pp};
    // cooperative launch: the grid barrier needs every CTA resident (grid = #SMs, 1 CTA/SM)
    B2_CUDA_CHECK(cudaLaunchCooperativeKernel(fn, dim3(num_sms()), dim3(MK_THREADS), args, smem, stream));
    B2_LAUNCH_CHECK();
    return 0;
}

}  // namespace b2
