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

// sense-reversing grid barrier (all CTAs co-resident: cooperative launch, grid = #SMs)
__device__ __forceinline__ void grid_sync(unsigned int* count, unsigned int* gen, unsigned int nblocks) {
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int g;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(g) : "l"(gen) : "memory");
        __threadfence();
        if (atomicAdd(count, 1u) == nblocks - 1) {
            *reinterpret_cast<volatile unsigned int*>(count) = 0u;
            __threadfence();
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(gen) : "memory");
        } else {
            unsigned int cur;
            unsigned int spins = 0;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(cur) : "l"(gen) : "memory");
                if (++spins > (1u << 26)) asm volatile("trap;");  // protocol bug -> CUDA error, not a hang
            } while (cur == g);
        }
    }
    __syncthreads();
}

// hot per-phase state (lives in registers through the weight-streaming loop)
struct GemvCtx {
    const __nv_bfloat16* W;
    int K, act;
    int nchunks, G, total_steps;
};
// cold per-phase I/O, recomputed from the phase index where needed (prologue, per-item epilogue)
struct PhaseIO {
    const __nv_bfloat16* xin;
    const __nv_bfloat16* gamma;
    const __nv_bfloat16* residual;
    void* out;
    int ld_out, out_fp32;
};

struct WarpId { int tid, lane, warp, gw, total_warps; };
__device__ __forceinline__ PhaseIO mk_phase_io(const MegaParams& p, int ph);

__device__ __forceinline__ void mk_item_rows(int act, int item, int& r0, int& r1) {
    if (act == ACT_SWIGLU) { r0 = (item >> 6) * 128 + (item & 63); r1 = r0 + 64; }
    else { r0 = item * 2; r1 = r0 + 1; }
}

__device__ __forceinline__ void mk_issue(const GemvCtx& c, int s, const WarpId& w, uint4 (&buf)[2][MK_U]) {
    if (s < c.total_steps) {
        const int item = w.gw + (s / c.G) * w.total_warps;
        const int g = s % c.G;
        int rr[2];
        mk_item_rows(c.act, item, rr[0], rr[1]);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const __nv_bfloat16* wr = c.W + (size_t)rr[r] * c.K + w.lane * 8;
#pragma unroll
            for (int u = 0; u < MK_U; ++u) {
                const int ch = g * MK_U + u;
                buf[r][u] = (ch < c.nchunks) ? ld_stream_16(wr + ch * 256) : make_uint4(0, 0, 0, 0);
            }
        }
    } else {
        // full (re)definition on every path: keeps the buffers' live ranges short for the register allocator
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int u = 0; u < MK_U; ++u) buf[r][u] = make_uint4(0, 0, 0, 0);
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

template <int NB>
__device__ __forceinline__ void mk_compute(const MegaParams& p, int ph, const GemvCtx& c, int s, int B,
                                           const WarpId& w, const uint4 (&buf)[2][MK_U],
                                           const __nv_bfloat16* xs, float (&acc)[2][NB]) {
    if (s >= c.total_steps) return;
    const int item = w.gw + (s / c.G) * w.total_warps;
    const int g = s % c.G;
    if (g == 0) {
#pragma unroll
        for (int b = 0; b < NB; ++b) acc[0][b] = acc[1][b] = 0.f;
    }
#pragma unroll
    for (int u = 0; u < MK_U; ++u) {
        const int ch = g * MK_U + u;
        if (ch < c.nchunks) {
            float w0[8], w1[8];
            unpack8(buf[0][u], w0);
            unpack8(buf[1][u], w1);
            const int koff = ch * 256 + w.lane * 8;
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                float xf[8];
                unpack8(*reinterpret_cast<const uint4*>(xs + (size_t)b * c.K + koff), xf);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    acc[0][b] = fmaf(w0[e], xf[e], acc[0][b]);
                    acc[1][b] = fmaf(w1[e], xf[e], acc[1][b]);
                }
            }
        }
    }
    if (g == c.G - 1) {
#pragma unroll
        for (int b = 0; b < NB; ++b) { acc[0][b] = warp_sum(acc[0][b]); acc[1][b] = warp_sum(acc[1][b]); }
        int r0, r1;
        mk_item_rows(c.act, item, r0, r1);
        const PhaseIO io = mk_phase_io(p, ph);
        if (c.act == ACT_SWIGLU) {
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                if (w.lane == b && b < B) {
                    const float gt = acc[0][b], up = acc[1][b];
                    reinterpret_cast<__nv_bfloat16*>(io.out)[(size_t)b * io.ld_out + item] =
                        __float2bfloat16_rn(gt / (1.0f + __expf(-gt)) * up);
                }
            }
        } else {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    if (w.lane == r * NB + b && b < B) {
                        const int row = r == 0 ? r0 : r1;
                        float y = acc[r][b];
                        if (io.residual != nullptr) y += __bfloat162float(__ldcg(io.residual + (size_t)b * io.ld_out + row));
                        if (io.out_fp32) reinterpret_cast<float*>(io.out)[(size_t)b * io.ld_out + row] = y;
                        else reinterpret_cast<__nv_bfloat16*>(io.out)[(size_t)b * io.ld_out + row] = __float2bfloat16_rn(y);
                    }
                }
            }
        }
    }
}

// RoPE + cache append + split-KV attention, one (b, head, split) item per warp; last warp of a (b, head) merges
__device__ __forceinline__ void mk_attention(const MegaParams& p, const MegaLayer& Lw, const WarpId& w) {
    const int h = p.h, H = p.H;
    const int n_items = p.B * H * p.nsplit;
    const int lane = w.lane;
    const int hw = lane >> 4, c = lane & 15;
    for (int item = w.gw; item < n_items; item += w.total_warps) {
        const int split = item % p.nsplit;
        const int bh = item / p.nsplit;
        const int head = bh % H, b = bh / H;
        const int pos = p.cur_len[b];
        const int total = pos + 1;
        const int chunk = (total + p.nsplit - 1) / p.nsplit;
        const int k_begin = split * chunk;
        const int k_end = min(k_begin + chunk, total);
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
                const int i = ((c & 7) * 8 + e);  // frequency index 0..63
                const float inv_freq = exp2f(-(2.0f * i / MK_D) * log2f(p.theta));
                float sv, cv;
                sincosf(pos * inv_freq, &sv, &cv);
                const float cb = round_bf16(cv), sb = round_bf16(sv);
                qreg[e] = round_bf16(round_bf16(qa[e] * cb) + round_bf16(sign * qb[e] * sb));
                knew[e] = round_bf16(round_bf16(ka[e] * cb) + round_bf16(sign * kb[e] * sb));
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
        float* part = p.attn_partial + (size_t)item * (MK_D + 2);
        if (hw == 0) {
            // partial rows are (128+2) floats: 8-byte aligned only -> float2 stores
            float2* pp = reinterpret_cast<float2*>(part + c * 8);
            pp[0] = make_float2(av[0], av[1]); pp[1] = make_float2(av[2], av[3]);
            pp[2] = make_float2(av[4], av[5]); pp[3] = make_float2(av[6], av[7]);
            if (c == 0) { part[MK_D] = m_run; part[MK_D + 1] = l_run; }
        }
        __threadfence();
        __syncwarp();
        int last = 0;
        if (lane == 0) last = (atomicAdd(&p.attn_counters[bh], 1) == p.nsplit - 1) ? 1 : 0;
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
            __threadfence();
            const float* pb = p.attn_partial + (size_t)bh * p.nsplit * (MK_D + 2);
            float m_all = -INFINITY;
            for (int s = 0; s < p.nsplit; ++s) m_all = fmaxf(m_all, __ldcg(pb + (size_t)s * (MK_D + 2) + MK_D));
            float l_all = 0.f, o4[4] = {0.f, 0.f, 0.f, 0.f};
            for (int s = 0; s < p.nsplit; ++s) {
                const float ms = __ldcg(pb + (size_t)s * (MK_D + 2) + MK_D);
                const float wgt = (ms == -INFINITY) ? 0.f : exp2f(ms - m_all);
                l_all += __ldcg(pb + (size_t)s * (MK_D + 2) + MK_D + 1) * wgt;
                const float2 oa = __ldcg(reinterpret_cast<const float2*>(pb + (size_t)s * (MK_D + 2) + lane * 4));
                const float2 ob = __ldcg(reinterpret_cast<const float2*>(pb + (size_t)s * (MK_D + 2) + lane * 4 + 2));
                o4[0] += oa.x * wgt; o4[1] += oa.y * wgt; o4[2] += ob.x * wgt; o4[3] += ob.y * wgt;
            }
            const float inv = 1.f / l_all;
            __nv_bfloat16* op = p.attn + (size_t)b * h + head * MK_D + lane * 4;
            *reinterpret_cast<uint2*>(op) = make_uint2(pack_bf16(o4[0] * inv, o4[1] * inv),
                                                       pack_bf16(o4[2] * inv, o4[3] * inv));
            if (lane == 0) p.attn_counters[bh] = 0;
        }
    }
}

// phase k of layer l (k: 0 = QKV, 1 = attention, 2 = o_proj, 3 = gate/up, 4 = down); index 5L = lm_head
__device__ __forceinline__ PhaseIO mk_phase_io(const MegaParams& p, int ph) {
    PhaseIO c;
    const int l = ph / 5, k = ph % 5;
    c.gamma = nullptr; c.residual = nullptr; c.out_fp32 = 0;
    if (l >= p.L) { c.xin = p.x; c.gamma = p.final_norm; c.out = p.logits; c.ld_out = p.V; c.out_fp32 = 1; }
    else if (k <= 1) { c.xin = p.x; c.gamma = p.layers[l].ln1; c.out = p.qkv; c.ld_out = 3 * p.h; }
    else if (k == 2) { c.xin = p.attn; c.residual = p.x; c.out = p.x; c.ld_out = p.h; }
    else if (k == 3) { c.xin = p.x; c.gamma = p.layers[l].ln2; c.out = p.act; c.ld_out = p.I; }
    else { c.xin = p.act; c.residual = p.x; c.out = p.x; c.ld_out = p.h; }
    return c;
}
__device__ __forceinline__ GemvCtx mk_phase_ctx(const MegaParams& p, int ph, const WarpId& w) {
    GemvCtx c;
    const int l = ph / 5, k = ph % 5;
    int N;
    c.act = ACT_NONE;
    if (l >= p.L) { c.W = p.lm_head; N = p.V; c.K = p.h; }
    else if (k <= 1) { c.W = p.layers[l].wqkv; N = 3 * p.h; c.K = p.h; }
    else if (k == 2) { c.W = p.layers[l].wo; N = p.h; c.K = p.h; }
    else if (k == 3) { c.W = p.layers[l].wgu; N = 2 * p.I; c.K = p.h; c.act = ACT_SWIGLU; }
    else { c.W = p.layers[l].wd; N = p.h; c.K = p.I; }
    c.nchunks = c.K >> 8;
    c.G = (c.nchunks + MK_U - 1) / MK_U;
    const int n_items = N >> 1;
    const int n_my = w.gw < n_items ? (n_items - w.gw + w.total_warps - 1) / w.total_warps : 0;
    c.total_steps = n_my * c.G;
    return c;
}

template <int NB>
__global__ void __launch_bounds__(MK_THREADS, 1) decode_mega_kernel(MegaParams p) {
    extern __shared__ __align__(16) uint8_t mk_smem[];
    __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(mk_smem);  // [NB][Kmax]
    __shared__ float s_red[MK_WARPS][NB];
    __shared__ float s_rstd[NB];
    __shared__ float s_av[MK_WARPS];
    __shared__ int s_ai[MK_WARPS];

    WarpId w;
    w.tid = threadIdx.x; w.lane = w.tid & 31; w.warp = w.tid >> 5;
    w.gw = blockIdx.x * MK_WARPS + w.warp;
    w.total_warps = gridDim.x * MK_WARPS;
    const int B = p.B;

    uint4 bufA[2][MK_U], bufB[2][MK_U];
    float acc[2][NB];

    // ---------------- phase "-1": x = embed_tokens[tok]; layer-0 QKV weights already in flight ----------------
    GemvCtx cur = mk_phase_ctx(p, 0, w);
    mk_issue(cur, 0, w, bufA);
    if (blockIdx.x < B) {
        int t = p.tok[blockIdx.x];
        t = t < 0 ? 0 : (t >= p.V ? p.V - 1 : t);
        const uint4* src = reinterpret_cast<const uint4*>(p.embed + (size_t)t * p.h);
        uint4* dst = reinterpret_cast<uint4*>(p.x + (size_t)blockIdx.x * p.h);
        for (int i = w.tid; i < p.h / 8; i += MK_THREADS) dst[i] = src[i];
    }
    grid_sync(p.bar_count, p.bar_gen, gridDim.x);

    const int n_phases = 5 * p.L + 1;
#pragma unroll 1
    for (int ph = 0; ph < n_phases; ++ph) {
        if (ph % 5 == 1 && ph < 5 * p.L) {
            mk_attention(p, p.layers[ph / 5], w);
        } else {
            // the first step's weights were issued (into bufA) before the preceding barrier
            mk_prologue<NB>(mk_phase_io(p, ph), cur.K, B, p.eps, w, xs, s_red, s_rstd);
#pragma unroll 1
            for (int s = 0; s < cur.total_steps; s += 2) {
                mk_issue(cur, s + 1, w, bufB);
                mk_compute<NB>(p, ph, cur, s, B, w, bufA, xs, acc);
                mk_issue(cur, s + 2, w, bufA);
                mk_compute<NB>(p, ph, cur, s + 1, B, w, bufB, xs, acc);
            }
        }
        // prefetch the next GEMV phase's first weights across the barrier (weights don't depend on activations)
        int nxt = ph + 1;
        if (nxt % 5 == 1 && nxt < 5 * p.L) nxt = -1;  // attention follows: no weights to prefetch yet
        if (nxt >= 0 && nxt < n_phases) {
            cur = mk_phase_ctx(p, nxt, w);
            mk_issue(cur, 0, w, bufA);
        } else {
            cur.total_steps = 0;
            mk_issue(cur, 0, w, bufA);  // defines bufA (zeros): nothing is carried across the attention phase
        }
        grid_sync(p.bar_count, p.bar_gen, gridDim.x);
    }

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

int decode_mega(const MegaParams& p, cudaStream_t stream) {
    B2_CHECK_ARG(p.B >= 1 && p.B <= 8, "decode_mega: batch must be 1..8");
    B2_CHECK_ARG(p.h % 256 == 0 && p.I % 256 == 0 && p.V % 2 == 0 && p.h / p.H == MK_D,
                 "decode_mega: unsupported dims h=%d I=%d V=%d H=%d", p.h, p.I, p.V, p.H);
    const int NB = p.B == 1 ? 1 : (p.B == 2 ? 2 : (p.B <= 4 ? 4 : 8));
    const int kmax = p.h > p.I ? p.h : p.I;
    const size_t smem = (size_t)NB * kmax * 2;
    B2_CHECK_ARG(smem <= 224 * 1024, "decode_mega: activations do not fit shared memory (B=%d K=%d)", p.B, kmax);
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
