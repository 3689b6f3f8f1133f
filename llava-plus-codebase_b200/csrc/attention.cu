// Attention kernels.
//   flash_attn_bf16  : prefill / ViT attention, flash-style online softmax, never materialises the
//                      [B,H,S,S] score matrix the reference's eager path builds
//                      (transformers modeling_llama.py:199-221 eager_attention_forward, modeling_clip.py:261-279).
//                      bf16 operands, fp32 softmax statistics and accumulators. Round-1 implementation uses
//                      mma.sync m16n8k16 (HMMA) tiles; it is <2% of prefill FLOPs.
//   rope_kv_write    : RoPE (modeling_llama.py:124-168, half-split rotate_half) on q,k of a prefill chunk and
//                      the KV-cache write (modeling_llama.py:269-270 cache update).
//   decode_attn_bf16 : one-token decode: RoPE + cache append + split-KV attention with coalesced 16-byte
//                      cache reads and an in-kernel last-CTA merge (HBM-bound: reads each K/V row once).
#include <math.h>

#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace b2 {
namespace {

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc, bool valid) {
    const uint32_t d = smem_u32(smem_dst);
    const int sz = valid ? 16 : 0;  // src-size 0 => zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_ptr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_u32(smem_ptr)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_ptr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_u32(smem_ptr)));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
        "{%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ------------------------------------------------------------------------------------------------
// flash attention forward (prefill / ViT)
// ------------------------------------------------------------------------------------------------
constexpr int FA_BM = 64;  // query rows per CTA (4 warps x 16)
constexpr int FA_BN = 64;  // keys per tile

struct FlashParams {
    const __nv_bfloat16* q; int64_t q_bs, q_ts, q_hs;
    const __nv_bfloat16* k; int64_t k_bs, k_ts, k_hs;
    const __nv_bfloat16* v; int64_t v_bs, v_ts, v_hs;
    __nv_bfloat16* o;       int64_t o_bs, o_ts, o_hs;
    const int32_t* seq_lens;
    int S;
    float scale_log2;
};

template <int D, bool CAUSAL>
__global__ void __launch_bounds__(128) flash_fwd_kernel(FlashParams p) {
    constexpr int LD = D + 8;  // padded smem row (elements): 16B-aligned rows, conflict-free ldmatrix
    extern __shared__ __align__(16) uint8_t fa_smem[];
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(fa_smem);  // [64][LD]
    __nv_bfloat16* sK = sQ + FA_BM * LD;                            // [2][64][LD]
    __nv_bfloat16* sV = sK + 2 * FA_BN * LD;                        // [2][64][LD]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * FA_BM;
    const int head = blockIdx.y, b = blockIdx.z;
    const int len = p.seq_lens != nullptr ? p.seq_lens[b] : p.S;  // valid keys of this sample
    const int kv_end = CAUSAL ? min(len, q0 + FA_BM) : len;
    const int n_tiles = (kv_end + FA_BN - 1) / FA_BN;

    const __nv_bfloat16* qg = p.q + b * p.q_bs + head * p.q_hs;
    const __nv_bfloat16* kg = p.k + b * p.k_bs + head * p.k_hs;
    const __nv_bfloat16* vg = p.v + b * p.v_bs + head * p.v_hs;

    constexpr int CHUNKS = D / 8;  // 16B chunks per row
    auto load_q = [&]() {
        for (int i = tid; i < FA_BM * CHUNKS; i += 128) {
            const int r = i / CHUNKS, c = i % CHUNKS;
            const int t = q0 + r;
            cp_async_16(sQ + r * LD + c * 8, qg + (int64_t)min(t, p.S - 1) * p.q_ts + c * 8, t < p.S);
        }
    };
    auto load_kv = [&](int tile, int buf) {
        __nv_bfloat16* dk = sK + buf * FA_BN * LD;
        __nv_bfloat16* dv = sV + buf * FA_BN * LD;
        for (int i = tid; i < FA_BN * CHUNKS; i += 128) {
            const int r = i / CHUNKS, c = i % CHUNKS;
            const int t = tile * FA_BN + r;
            const bool ok = t < len;
            const int tt = ok ? t : 0;
            cp_async_16(dk + r * LD + c * 8, kg + (int64_t)tt * p.k_ts + c * 8, ok);
            cp_async_16(dv + r * LD + c * 8, vg + (int64_t)tt * p.v_ts + c * 8, ok);
        }
    };

    load_q();
    if (n_tiles > 0) load_kv(0, 0);
    cp_async_commit();

    uint32_t qf[D / 16][4];
    float oacc[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};

    const int g = lane >> 2, tq = lane & 3;
    const int qrow0 = q0 + warp * 16 + g;  // this thread's rows: qrow0 and qrow0 + 8

    for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) load_kv(j + 1, (j + 1) & 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        if (j == 0) {
#pragma unroll
            for (int kk = 0; kk < D / 16; ++kk) {
                const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int c = kk * 16 + (lane >> 4) * 8;
                ldmatrix_x4(qf[kk], sQ + r * LD + c);
            }
        }
        const __nv_bfloat16* tK = sK + (j & 1) * FA_BN * LD;
        const __nv_bfloat16* tV = sV + (j & 1) * FA_BN * LD;

        // ---- S = Q K^T (16 x 64 per warp) ----
        float s[FA_BN / 8][4];
#pragma unroll
        for (int nt = 0; nt < FA_BN / 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
#pragma unroll
            for (int np = 0; np < FA_BN / 16; ++np) {
                uint32_t bfr[4];
                const int r = np * 16 + (lane & 7) + (lane >> 4) * 8;  // key row
                const int c = kk * 16 + ((lane >> 3) & 1) * 8;         // d column
                ldmatrix_x4(bfr, tK + r * LD + c);
                mma_bf16_16816(s[2 * np], qf[kk], bfr[0], bfr[1]);
                mma_bf16_16816(s[2 * np + 1], qf[kk], bfr[2], bfr[3]);
            }
        }

        // ---- scale, mask, online softmax ----
        const int key0 = j * FA_BN;
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < FA_BN / 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int key = key0 + nt * 8 + tq * 2 + (e & 1);
                const int qr = qrow0 + (e >> 1) * 8;
                bool ok = key < len;
                if (CAUSAL) ok = ok && (key <= qr);
                const float val = ok ? s[nt][e] * p.scale_log2 : -INFINITY;
                s[nt][e] = val;
                mx[e >> 1] = fmaxf(mx[e >> 1], val);
            }
        }
        float corr[2], m_use[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
            mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
            const float m_new = fmaxf(m_run[h], mx[h]);
            m_use[h] = (m_new == -INFINITY) ? 0.f : m_new;
            corr[h] = exp2f(m_run[h] - m_use[h]);  // m_run = -inf -> 0
            m_run[h] = m_new;
            l_run[h] *= corr[h];
        }
        float rs[2] = {0.f, 0.f};
#pragma unroll
        for (int nt = 0; nt < FA_BN / 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float pv = exp2f(s[nt][e] - m_use[e >> 1]);
                s[nt][e] = pv;
                rs[e >> 1] += pv;
            }
        }
        l_run[0] += rs[0];
        l_run[1] += rs[1];
#pragma unroll
        for (int i = 0; i < D / 8; ++i) {
            oacc[i][0] *= corr[0]; oacc[i][1] *= corr[0];
            oacc[i][2] *= corr[1]; oacc[i][3] *= corr[1];
        }

        // ---- O += P V ----
#pragma unroll
        for (int kk = 0; kk < FA_BN / 16; ++kk) {
            uint32_t pa[4];
            pa[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
            pa[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
            pa[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
            pa[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
            for (int dp = 0; dp < D / 16; ++dp) {
                uint32_t bfr[4];
                const int r = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;  // key row
                const int c = dp * 16 + (lane >> 4) * 8;                      // d column
                ldmatrix_x4_trans(bfr, tV + r * LD + c);
                mma_bf16_16816(oacc[2 * dp], pa, bfr[0], bfr[1]);
                mma_bf16_16816(oacc[2 * dp + 1], pa, bfr[2], bfr[3]);
            }
        }
        __syncthreads();  // everyone done with buffer (j&1) before it is refilled at iteration j+1
    }
    cp_async_wait<0>();

    // ---- finalise ----
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 1);
        l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 2);
    }
    __nv_bfloat16* og = p.o + b * p.o_bs + head * p.o_hs;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int t = qrow0 + h * 8;
        if (t < p.S) {
            const float inv = l_run[h] > 0.f ? 1.f / l_run[h] : 0.f;
#pragma unroll
            for (int i = 0; i < D / 8; ++i) {
                const uint32_t pk = pack_bf16(oacc[i][2 * h] * inv, oacc[i][2 * h + 1] * inv);
                *reinterpret_cast<uint32_t*>(og + (int64_t)t * p.o_ts + i * 8 + tq * 2) = pk;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// RoPE helpers (HF semantics: cos/sin computed in fp32, cast to bf16, products rounded to bf16)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void rope_cos_sin(int pos, int i /*0..D/2*/, int D, float theta, float& c, float& s) {
    // inv_freq = theta^(-2i/D)
    const float inv_freq = exp2f(-(2.0f * i / D) * log2f(theta));
    const float ang = pos * inv_freq;
    float sv, cv;
    sincosf(ang, &sv, &cv);
    c = round_bf16(cv);
    s = round_bf16(sv);
}
__device__ __forceinline__ float rope_apply(float x, float partner_signed, float c, float s) {
    // q*cos + rotate_half(q)*sin with bf16 rounding of each product and of the sum
    return round_bf16(round_bf16(x * c) + round_bf16(partner_signed * s));
}

// the same (cos, sin) values as a table [Smax][D/2] of bf16 pairs (cos | sin << 16) for the QKV GEMM's fused RoPE epilogue
// (gemm_2cta.cu, ACT_ROPE_QKV)
__global__ void rope_table_kernel(uint32_t* __restrict__ table, int Smax, int D, float theta) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int half = D / 2;
    if (idx >= Smax * half) return;
    float c, s;
    rope_cos_sin(idx / half, idx % half, D, theta, c, s);
    table[idx] = pack_bf16(c, s);  // both are bf16-rounded already: the packing is exact
}

// grid: (B*S) tokens, 256 threads. The D/2 (cos, sin) pairs of the token's position are tabulated once in shared
// memory (they are the same for every head); every thread then moves 16-byte vectors: one (head, 8-lane chunk) item
// = q/k/v elements [c*8, c*8+8) and their rotation partners [D/2 + c*8, ...), six loads and six stores of 16 B.
__global__ void __launch_bounds__(256)
rope_kv_write_kernel(__nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ kcache,
                     __nv_bfloat16* __restrict__ vcache, int S, int H, int D, int Smax, float theta) {
    __shared__ float s_c[128], s_s[128];  // D/2 <= 128
    pdl_trigger();
    pdl_wait();  // inputs are outputs of the upstream kernel (programmatic dependent launch)
    const int row = blockIdx.x;  // b*S + t
    const int b = row / S, t = row % S;
    const int hd = H * D;
    const int half = D / 2;
    for (int i = threadIdx.x; i < half; i += blockDim.x) rope_cos_sin(t, i, D, theta, s_c[i], s_s[i]);
    __syncthreads();
    __nv_bfloat16* q = qkv + (size_t)row * 3 * hd;
    __nv_bfloat16* k = q + hd;
    const __nv_bfloat16* v = q + 2 * hd;
    const int cpd = half / 8;  // 8-element chunks per half head
    for (int idx = threadIdx.x; idx < H * cpd; idx += blockDim.x) {
        const int h = idx / cpd, c = idx % cpd;
        const int o1 = h * D + c * 8, o2 = o1 + half;
        const uint4 q1 = *reinterpret_cast<const uint4*>(q + o1), q2 = *reinterpret_cast<const uint4*>(q + o2);
        const uint4 k1 = *reinterpret_cast<const uint4*>(k + o1), k2 = *reinterpret_cast<const uint4*>(k + o2);
        const uint4 v1 = *reinterpret_cast<const uint4*>(v + o1), v2 = *reinterpret_cast<const uint4*>(v + o2);
        const uint32_t qa[4] = {q1.x, q1.y, q1.z, q1.w}, qb[4] = {q2.x, q2.y, q2.z, q2.w};
        const uint32_t ka[4] = {k1.x, k1.y, k1.z, k1.w}, kb[4] = {k2.x, k2.y, k2.z, k2.w};
        uint32_t oq1[4], oq2[4], ok1[4], ok2[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float c0 = s_c[c * 8 + 2 * e], s0 = s_s[c * 8 + 2 * e];
            const float c1 = s_c[c * 8 + 2 * e + 1], s1 = s_s[c * 8 + 2 * e + 1];
            const float qa0 = bf16_lo(qa[e]), qa1 = bf16_hi(qa[e]), qb0 = bf16_lo(qb[e]), qb1 = bf16_hi(qb[e]);
            const float ka0 = bf16_lo(ka[e]), ka1 = bf16_hi(ka[e]), kb0 = bf16_lo(kb[e]), kb1 = bf16_hi(kb[e]);
            oq1[e] = pack_bf16(rope_apply(qa0, -qb0, c0, s0), rope_apply(qa1, -qb1, c1, s1));
            oq2[e] = pack_bf16(rope_apply(qb0, qa0, c0, s0), rope_apply(qb1, qa1, c1, s1));
            ok1[e] = pack_bf16(rope_apply(ka0, -kb0, c0, s0), rope_apply(ka1, -kb1, c1, s1));
            ok2[e] = pack_bf16(rope_apply(kb0, ka0, c0, s0), rope_apply(kb1, ka1, c1, s1));
        }
        *reinterpret_cast<uint4*>(q + o1) = make_uint4(oq1[0], oq1[1], oq1[2], oq1[3]);
        *reinterpret_cast<uint4*>(q + o2) = make_uint4(oq2[0], oq2[1], oq2[2], oq2[3]);
        const size_t co = (((size_t)b * H + h) * Smax + t) * D + c * 8;
        *reinterpret_cast<uint4*>(kcache + co) = make_uint4(ok1[0], ok1[1], ok1[2], ok1[3]);
        *reinterpret_cast<uint4*>(kcache + co + half) = make_uint4(ok2[0], ok2[1], ok2[2], ok2[3]);
        *reinterpret_cast<uint4*>(vcache + co) = v1;
        *reinterpret_cast<uint4*>(vcache + co + half) = v2;
    }
}

// ------------------------------------------------------------------------------------------------
// decode attention: grid (nsplit, H, B), 128 threads. D = 128 fixed: a half-warp (16 lanes x 8 elems)
// covers one K/V row with one 16-byte load per lane.
// ------------------------------------------------------------------------------------------------
constexpr int DA_D = 128;
constexpr int DA_THREADS = 128;
constexpr int DA_UNROLL = 4;

struct DecodeAttnParams {
    const __nv_bfloat16* qkv;
    __nv_bfloat16* kcache;
    __nv_bfloat16* vcache;
    const int32_t* cur_len;
    __nv_bfloat16* out;
    float* partial;
    int32_t* counters;
    int H, Smax, nsplit;
    float theta, scale_log2;
};

__global__ void __launch_bounds__(DA_THREADS) decode_attn_kernel(DecodeAttnParams p) {
    const int split = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int hw = (warp << 1) | (lane >> 4);  // half-warp id 0..7
    const int c = lane & 15;                   // 8-element chunk of the head dim
    pdl_trigger();
    pdl_wait();                                // qkv (previous GEMM) and cur_len (previous step) are upstream outputs
    const int pos = p.cur_len[b];              // position of the new token == number of cached keys
    const int total = pos + 1;
    const int hd = p.H * DA_D;

    __shared__ float s_q[DA_D];
    __shared__ float s_knew[DA_D];
    __shared__ float s_m[8], s_l[8];
    __shared__ float s_o[8][DA_D];
    __shared__ int s_last;

    // ---- RoPE on q and the new k (every CTA: 128 threads, one element each) ----
    {
        const __nv_bfloat16* qrow = p.qkv + (size_t)b * 3 * hd + head * DA_D;
        const __nv_bfloat16* krow = qrow + hd;
        const int i = tid & 63;
        float cs, sn;
        rope_cos_sin(pos, i, DA_D, p.theta, cs, sn);
        const float q1 = __bfloat162float(qrow[i]), q2 = __bfloat162float(qrow[i + 64]);
        const float k1 = __bfloat162float(krow[i]), k2 = __bfloat162float(krow[i + 64]);
        if (tid < 64) {
            s_q[i] = rope_apply(q1, -q2, cs, sn);
            s_knew[i] = rope_apply(k1, -k2, cs, sn);
        } else {
            s_q[i + 64] = rope_apply(q2, q1, cs, sn);
            s_knew[i + 64] = rope_apply(k2, k1, cs, sn);
        }
    }
    __syncthreads();

    // key range of this split (device-side: the launch grid is fixed so the step can live in a CUDA graph)
    const int chunk = (total + p.nsplit - 1) / p.nsplit;
    const int k_begin = split * chunk;
    const int k_end = min(k_begin + chunk, total);
    const bool owns_new = (pos >= k_begin) && (pos < k_end);

    const size_t cbase = ((size_t)b * p.H + head) * p.Smax * DA_D;
    if (owns_new) {
        // append the new token's k (roped) and v to the cache; exactly one CTA per (b, head) does this
        const __nv_bfloat16* vrow = p.qkv + (size_t)b * 3 * hd + 2 * hd + head * DA_D;
        p.kcache[cbase + (size_t)pos * DA_D + tid] = __float2bfloat16_rn(s_knew[tid]);
        p.vcache[cbase + (size_t)pos * DA_D + tid] = vrow[tid];
    }

    float qreg[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) qreg[e] = s_q[c * 8 + e];

    float m_run = -INFINITY, l_run = 0.f;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;

    const __nv_bfloat16* kb = p.kcache + cbase + c * 8;
    const __nv_bfloat16* vb = p.vcache + cbase + c * 8;
    const __nv_bfloat16* vnew = p.qkv + (size_t)b * 3 * hd + 2 * hd + head * DA_D + c * 8;

    // each half-warp walks keys k_begin + hw, +8, ...; DA_UNROLL keys (2*DA_UNROLL 16B loads) in flight
    // (trip count is CTA-uniform: the shuffles below use the full warp mask)
    for (int kbase = k_begin; kbase < k_end; kbase += 8 * DA_UNROLL) {
        const int k0 = kbase + hw;
        uint4 kraw[DA_UNROLL], vraw[DA_UNROLL];
#pragma unroll
        for (int u = 0; u < DA_UNROLL; ++u) {
            const int key = k0 + u * 8;
            if (key < k_end && key != pos) {
                kraw[u] = ld_stream_16(kb + (size_t)key * DA_D);
                vraw[u] = ld_stream_16(vb + (size_t)key * DA_D);
            } else {
                kraw[u] = make_uint4(0, 0, 0, 0);
                vraw[u] = make_uint4(0, 0, 0, 0);
            }
        }
#pragma unroll
        for (int u = 0; u < DA_UNROLL; ++u) {
            const int key = k0 + u * 8;
            const bool valid = key < k_end;  // uniform across the half-warp
            float kf[8], vf[8];
            if (key == pos) {
                // the new token: k from smem (already roped, bf16-rounded), v straight from qkv
#pragma unroll
                for (int e = 0; e < 8; ++e) kf[e] = round_bf16(s_knew[c * 8 + e]);
                const uint4 vv = *reinterpret_cast<const uint4*>(vnew);
                vf[0] = bf16_lo(vv.x); vf[1] = bf16_hi(vv.x); vf[2] = bf16_lo(vv.y); vf[3] = bf16_hi(vv.y);
                vf[4] = bf16_lo(vv.z); vf[5] = bf16_hi(vv.z); vf[6] = bf16_lo(vv.w); vf[7] = bf16_hi(vv.w);
            } else {
                kf[0] = bf16_lo(kraw[u].x); kf[1] = bf16_hi(kraw[u].x); kf[2] = bf16_lo(kraw[u].y);
                kf[3] = bf16_hi(kraw[u].y); kf[4] = bf16_lo(kraw[u].z); kf[5] = bf16_hi(kraw[u].z);
                kf[6] = bf16_lo(kraw[u].w); kf[7] = bf16_hi(kraw[u].w);
                vf[0] = bf16_lo(vraw[u].x); vf[1] = bf16_hi(vraw[u].x); vf[2] = bf16_lo(vraw[u].y);
                vf[3] = bf16_hi(vraw[u].y); vf[4] = bf16_lo(vraw[u].z); vf[5] = bf16_hi(vraw[u].z);
                vf[6] = bf16_lo(vraw[u].w); vf[7] = bf16_hi(vraw[u].w);
            }
            float dot = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) dot += qreg[e] * kf[e];
            // reduce over the 16 lanes of the half-warp (xor 8,4,2,1 stays inside the half)
            dot += __shfl_xor_sync(0xffffffffu, dot, 8);
            dot += __shfl_xor_sync(0xffffffffu, dot, 4);
            dot += __shfl_xor_sync(0xffffffffu, dot, 2);
            dot += __shfl_xor_sync(0xffffffffu, dot, 1);
            if (valid) {
                const float sc = dot * p.scale_log2;
                const float m_new = fmaxf(m_run, sc);
                const float corr = exp2f(m_run - m_new);
                const float pr = exp2f(sc - m_new);
                l_run = l_run * corr + pr;
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] = acc[e] * corr + pr * vf[e];
                m_run = m_new;
            }
        }
    }

    // ---- merge the 8 half-warps of this CTA ----
    if (c == 0) { s_m[hw] = m_run; s_l[hw] = l_run; }
#pragma unroll
    for (int e = 0; e < 8; ++e) s_o[hw][c * 8 + e] = acc[e];
    __syncthreads();
    float m_cta = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i) m_cta = fmaxf(m_cta, s_m[i]);
    float l_cta = 0.f, o_cta = 0.f;  // thread tid owns output element tid
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float w = (s_m[i] == -INFINITY) ? 0.f : exp2f(s_m[i] - m_cta);
        l_cta += s_l[i] * w;
        o_cta += s_o[i][tid] * w;
    }
    const int bh = b * p.H + head;
    float* part = p.partial + ((size_t)bh * p.nsplit + split) * (DA_D + 2);
    part[tid] = o_cta;
    if (tid == 0) { part[DA_D] = m_cta; part[DA_D + 1] = l_cta; }

    // ---- last CTA of this (b, head) merges the splits ----
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const int prev = atomicAdd(&p.counters[bh], 1);
        s_last = (prev == p.nsplit - 1) ? 1 : 0;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        const float* pb = p.partial + (size_t)bh * p.nsplit * (DA_D + 2);
        float m_all = -INFINITY;
#pragma unroll 8
        for (int s = 0; s < p.nsplit; ++s) m_all = fmaxf(m_all, __ldcg(pb + (size_t)s * (DA_D + 2) + DA_D));
        float l_all = 0.f, o_all = 0.f;
#pragma unroll 8
        for (int s = 0; s < p.nsplit; ++s) {  // independent loads: unrolled so they overlap instead of chaining
            const float ms = __ldcg(pb + (size_t)s * (DA_D + 2) + DA_D);
            const float w = (ms == -INFINITY) ? 0.f : exp2f(ms - m_all);
            l_all += __ldcg(pb + (size_t)s * (DA_D + 2) + DA_D + 1) * w;
            o_all += __ldcg(pb + (size_t)s * (DA_D + 2) + tid) * w;
        }
        p.out[(size_t)b * hd + head * DA_D + tid] = __float2bfloat16_rn(o_all / l_all);
        if (tid == 0) p.counters[bh] = 0;  // self-reset for the next launch
    }
}

}  // namespace

int flash_attn_bf16(const FlashArgs& a, cudaStream_t stream) {
    B2_CHECK_ARG(a.D == 64 || a.D == 128, "flash_attn: head_dim must be 64 or 128 (got %d)", a.D);
    B2_CHECK_ARG(a.B > 0 && a.H > 0 && a.S > 0, "flash_attn: empty problem");
    // tcgen05 kernel unless B2_FLASH_TC=0 selects the mma.sync one (A/B knob, re-read per call: scripts/attn_bench.py)
    const char* e = getenv("B2_FLASH_TC");
    const bool use_tc = e == nullptr || e[0] != '0';
    return use_tc ? flash_attn_tc_bf16(a, stream) : flash_attn_mma_bf16(a, stream);
}

int flash_attn_mma_bf16(const FlashArgs& a, cudaStream_t stream) {
    FlashParams p;
    p.q = reinterpret_cast<const __nv_bfloat16*>(a.q); p.q_bs = a.q_bs; p.q_ts = a.q_ts; p.q_hs = a.q_hs;
    p.k = reinterpret_cast<const __nv_bfloat16*>(a.k); p.k_bs = a.k_bs; p.k_ts = a.k_ts; p.k_hs = a.k_hs;
    p.v = reinterpret_cast<const __nv_bfloat16*>(a.v); p.v_bs = a.v_bs; p.v_ts = a.v_ts; p.v_hs = a.v_hs;
    p.o = reinterpret_cast<__nv_bfloat16*>(a.o);       p.o_bs = a.o_bs; p.o_ts = a.o_ts; p.o_hs = a.o_hs;
    p.seq_lens = a.seq_lens;
    p.S = a.S;
    p.scale_log2 = a.scale * 1.4426950408889634f;
    dim3 grid((a.S + FA_BM - 1) / FA_BM, a.H, a.B);
    const int smem = (FA_BM + 4 * FA_BN) * (a.D + 8) * 2;
#define B2_FA_LAUNCH(DD, CC)                                                                              \
    do {                                                                                                  \
        static bool attr_set = false;                                                                     \
        if (!attr_set) {                                                                                  \
            B2_CUDA_CHECK(cudaFuncSetAttribute(flash_fwd_kernel<DD, CC>,                                  \
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, smem));       \
            attr_set = true;                                                                              \
        }                                                                                                 \
        flash_fwd_kernel<DD, CC><<<grid, 128, smem, stream>>>(p);                                         \
    } while (0)
    if (a.D == 64) {
        if (a.causal) B2_FA_LAUNCH(64, true); else B2_FA_LAUNCH(64, false);
    } else {
        if (a.causal) B2_FA_LAUNCH(128, true); else B2_FA_LAUNCH(128, false);
    }
#undef B2_FA_LAUNCH
    B2_LAUNCH_CHECK();
    return 0;
}

int rope_table_build(void* table, int Smax, int D, float theta, cudaStream_t stream) {
    B2_CHECK_ARG(table != nullptr && Smax > 0 && D > 0 && D % 2 == 0, "rope_table_build: bad argument");
    const int n = Smax * (D / 2);
    rope_table_kernel<<<(n + 255) / 256, 256, 0, stream>>>(reinterpret_cast<uint32_t*>(table), Smax, D, theta);
    B2_LAUNCH_CHECK();
    return 0;
}

int rope_kv_write(void* qkv, void* kcache, void* vcache, int B, int S, int H, int D, int Smax, float theta,
                  cudaStream_t stream) {
    B2_CHECK_ARG(S <= Smax, "rope_kv_write: S=%d exceeds cache capacity %d", S, Smax);
    B2_CHECK_ARG(D % 16 == 0 && D <= 256, "rope_kv_write: head_dim must be a multiple of 16, <= 256 (got %d)", D);
    B2_CHECK_ARG(((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(kcache) |
                   reinterpret_cast<uintptr_t>(vcache)) & 15) == 0, "rope_kv_write: buffers must be 16-byte aligned");
    B2_CUDA_CHECK(launch_pdl(rope_kv_write_kernel, dim3(B * S), dim3(256), 0, stream, reinterpret_cast<__nv_bfloat16*>(qkv),
                             reinterpret_cast<__nv_bfloat16*>(kcache), reinterpret_cast<__nv_bfloat16*>(vcache), S, H, D, Smax, theta));
    B2_LAUNCH_CHECK();
    return 0;
}

// resident CTAs of decode_attn_kernel per SM (register-limited: 79 registers x 128 threads -> 6), for the split heuristic
int decode_attn_ctas_per_sm() {
    static int occ = 0;
    if (occ == 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, decode_attn_kernel, DA_THREADS, 0) != cudaSuccess || n < 1) n = 6;
        occ = n;
    }
    return occ;
}

int decode_attn_bf16(const DecodeAttnArgs& a, cudaStream_t stream) {
    B2_CHECK_ARG(a.D == DA_D, "decode_attn: head_dim must be 128 (got %d)", a.D);
    B2_CHECK_ARG(a.nsplit >= 1 && a.B > 0 && a.H > 0, "decode_attn: bad launch shape");
    DecodeAttnParams p;
    p.qkv = reinterpret_cast<const __nv_bfloat16*>(a.qkv);
    p.kcache = reinterpret_cast<__nv_bfloat16*>(a.kcache);
    p.vcache = reinterpret_cast<__nv_bfloat16*>(a.vcache);
    p.cur_len = a.cur_len;
    p.out = reinterpret_cast<__nv_bfloat16*>(a.out);
    p.partial = a.partial;
    p.counters = a.counters;
    p.H = a.H; p.Smax = a.Smax; p.nsplit = a.nsplit;
    p.theta = a.theta;
    p.scale_log2 = a.scale * 1.4426950408889634f;
    dim3 grid(a.nsplit, a.H, a.B);
    B2_CUDA_CHECK(launch_pdl(decode_attn_kernel, grid, dim3(DA_THREADS), 0, stream, p));
    B2_LAUNCH_CHECK();
    return 0;
}

}  // namespace b2
