// tcgen05 flash attention for prefill / ViT (sm_100a): S = Q·K^T and O += P·V on the 5th-gen tensor cores with
// TMEM accumulators, TMA-fed operands, fp32 online softmax by one thread per query row.
//
// Replaces the eager attention of the reference's HF path, which materialises the [B,H,S,S] score matrix
// (transformers modeling_llama.py:199-221 eager_attention_forward; modeling_clip.py:261-279).
//
// CTA = one (batch, head, 128-query tile); 192 threads:
//   warp 0 / lane 0 : TMA producer — Q once, then K_j / V_j tiles (128 keys) into separate rings (K: 2 stages; V: 2
//                     stages at d=128, 1 at d=64 so that TWO CTAs fit one SM: one CTA's exp2-bound softmax overlaps
//                     the other's MMAs; at d=64 a 128x128 tile costs 1024 clk of MUFU but only 512 clk of tensor pipe)
//                     (4-D tensor maps over the strided [b, t, h, d] views: the same kernel reads the fused qkv
//                     activation buffer of the ViT and the [B, H, Smax, 128] KV cache of the decoder)
//   warp 1 / lane 0 : MMA issuer — S = Q K_j^T  (UMMA 128x128xD, both operands K-major, SWIZZLE_128B)
//                                  O += P_j V_j (UMMA 128xDx128; P from smem K-major, V straight from its
//                                  [keys, d] tile as an MN-major operand — no transpose pass)
//   warps 2..5      : softmax — thread r owns query row r: tcgen05.ld the S row, scale/mask, running max / sum, write
//                     P = exp2(s - m) (bf16) into the swizzled smem operand tile; O in TMEM is rescaled only when a
//                     row's max grew by more than 2^8 (lazy rescale: P and l keep using the stale max, which cancels
//                     in O / l), finally O / l -> global.
// mbarriers: q_full, k_full[2], k_empty[2], v_full[2], v_empty[2], s_full, p_full, o_done.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace b2 {
namespace {

constexpr int TC_BM = 128;  // queries per CTA
constexpr int TC_BN = 128;  // keys per tile
constexpr int TC_THREADS = 192;

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]),
        "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]),
        "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
// MUFU.EX2 without exp2f's denormal-range fix-up (inputs here are <= 8, results feed a bf16 / an fp32 sum)
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// smem descriptor for an MN-major bf16 operand tile stored as rows of 128 bytes (64 elements along MN) indexed by k,
// SWIZZLE_128B: atoms of 8 k-rows x 128 B; SBO = 1024 B between k-atoms, LBO = bytes between 64-wide MN slabs.
__device__ __forceinline__ uint64_t make_sw128_mnmajor_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
// kind::f16 instruction descriptor: A K-major, B K-major (b_mn = 0) or MN-major (b_mn = 1); bf16 x bf16 -> fp32
__host__ __device__ constexpr uint32_t make_idesc(uint32_t m, uint32_t n, uint32_t b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn << 16) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// 2^x without the special-function unit: x = n + f (n = round(x), |f| <= 0.5), 2^f by a degree-3 minimax polynomial (max
// relative error 7.5e-5, far below the bf16 rounding P gets anyway), 2^n by adding n to the exponent field. At d = 64 a
// 128 x 128 tile costs 1024 clk of MUFU.EX2 (16/clk/SM) against 512 clk of tensor pipe: the softmax warps are MUFU-bound,
// while the FMA and integer pipes idle. Sending every second element through this routine halves the MUFU time.
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -126.0f);                       // masked scores (-inf) -> 2^-126 ~ 1e-38: vanishes in bf16 P and in the row sum
    const float magic = 12582912.0f;             // 1.5 * 2^23: the sum's low mantissa bits hold round(x)
    const float xr = x + magic;
    const float f = x - (xr - magic);
    float q = fmaf(f, 0.05517144873738289f, 0.2426108419895172f);
    q = fmaf(f, q, 0.6932609677314758f);
    q = fmaf(f, q, 0.9999281167984009f);
    return __int_as_float(__float_as_int(q) + (__float_as_int(xr) << 23));
}

struct FlashTcParams {
    const int32_t* seq_lens;
    __nv_bfloat16* o;
    int64_t o_bs, o_ts, o_hs;
    int S;
    float scale_log2;
    int exp_poly;  // every second exp2 of the softmax on the FMA/ALU pipes instead of MUFU (B2_FLASH_EXP_POLY=1; measured slower)
};

template <int D>
struct TcCfg {
    static constexpr int VS = D == 64 ? 1 : 2;
    static constexpr int MIN_CTAS = D == 64 ? 2 : 1;
    static constexpr int SMEM = (3 + VS) * TC_BN * D * 2 + TC_BM * TC_BN * 2 + 1024 /*align*/ + 128 /*barriers*/;
};

template <int D, bool CAUSAL, int MIN_CTAS>
__global__ void __launch_bounds__(TC_THREADS, MIN_CTAS)
flash_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, FlashTcParams p) {
    constexpr int SLABS = D / 64;                 // 64-wide (128 B) slabs of the head dim
    constexpr int TILE_BYTES = TC_BN * D * 2;     // one Q / K / V tile
    constexpr int SLAB_BYTES = TC_BN * 128;       // [128 rows x 128 B]
    constexpr int P_BYTES = TC_BM * TC_BN * 2;    // [2 slabs of 64 keys][128 rows][128 B]
    constexpr int TMEM_COLS = 256;                // S: cols [0,128), O: cols [128, 128 + D)

    extern __shared__ uint8_t tc_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    constexpr int VS = TcCfg<D>::VS;          // V ring stages
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + TILE_BYTES;            // [2][TILE_BYTES]
    uint8_t* sV = sK + 2 * TILE_BYTES;        // [VS][TILE_BYTES]
    uint8_t* sP = sV + VS * TILE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + P_BYTES);
    uint64_t* q_full = bars;
    uint64_t* k_full = bars + 1;    // [2]
    uint64_t* k_empty = bars + 3;   // [2]
    uint64_t* v_full = bars + 5;    // [2]
    uint64_t* v_empty = bars + 7;   // [2]
    uint64_t* s_full = bars + 9;
    uint64_t* p_full = bars + 10;
    uint64_t* o_done = bars + 11;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * TC_BM;
    const int head = blockIdx.y, b = blockIdx.z;
    pdl_trigger();
    pdl_wait();  // q/k/v (and seq_lens) are outputs of upstream kernels (programmatic dependent launch)
    const int len = p.seq_lens != nullptr ? p.seq_lens[b] : p.S;
    const int kv_end = CAUSAL ? min(len, q0 + TC_BM) : len;
    const int n_tiles = (kv_end + TC_BN - 1) / TC_BN;  // >= 1 (len >= 1)

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_k);
        tma_prefetch_desc(&tm_v);
        mbar_init(q_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
            mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(p_full, 4);  // one arrive per softmax warp
        mbar_init(o_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128;

    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer ----------------
            mbar_arrive_expect_tx(q_full, TILE_BYTES);
            for (int sl = 0; sl < SLABS; ++sl) tma_load_4d(sQ + sl * SLAB_BYTES, &tm_q, q_full, sl * 64, q0, head, b);
            for (int j = 0; j < n_tiles; ++j) {
                const int ks = j & 1, vs = j % VS;
                mbar_wait(&k_empty[ks], ((j >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(&k_full[ks], TILE_BYTES);
                for (int sl = 0; sl < SLABS; ++sl)
                    tma_load_4d(sK + ks * TILE_BYTES + sl * SLAB_BYTES, &tm_k, &k_full[ks], sl * 64, j * TC_BN, head, b);
                mbar_wait(&v_empty[vs], ((j / VS) & 1) ^ 1);
                mbar_arrive_expect_tx(&v_full[vs], TILE_BYTES);
                for (int sl = 0; sl < SLABS; ++sl)
                    tma_load_4d(sV + vs * TILE_BYTES + sl * SLAB_BYTES, &tm_v, &v_full[vs], sl * 64, j * TC_BN, head, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---------------- MMA issuer ----------------
            constexpr uint32_t idesc_s = make_idesc(TC_BM, TC_BN, 0);
            constexpr uint32_t idesc_o = make_idesc(TC_BM, D, 1);
            mbar_wait(q_full, 0);
            for (int j = 0; j < n_tiles; ++j) {
                const int s = j & 1, vs = j % VS;
                mbar_wait(&k_full[s], (j >> 1) & 1);
                tc_fence_after();
                // S = Q K_j^T : K-loop over the head dim in steps of 16 (32 B inside a 128 B swizzle row; slabs of 64)
#pragma unroll
                for (int k = 0; k < D / 16; ++k) {
                    const uint32_t off = (k >> 2) * SLAB_BYTES + (k & 3) * 32;
                    umma_bf16(tmem_S, make_sw128_kmajor_desc(smem_u32(sQ) + off),
                              make_sw128_kmajor_desc(smem_u32(sK + s * TILE_BYTES) + off), idesc_s, k != 0 ? 1u : 0u);
                }
                umma_commit(&k_empty[s]);  // K_j consumed once S_j is complete
                umma_commit(s_full);
                // O += P_j V_j once the softmax warps have written P_j (and rescaled O)
                mbar_wait(&v_full[vs], (j / VS) & 1);
                mbar_wait(p_full, j & 1);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < TC_BN / 16; ++k) {
                    const uint32_t poff = (k >> 2) * (TC_BM * 128) + (k & 3) * 32;   // P: K-major, 2 slabs of 64 keys
                    const uint32_t voff = k * 16 * 128;                               // V: MN-major, 16 key rows per step
                    umma_bf16(tmem_O, make_sw128_kmajor_desc(smem_u32(sP) + poff),
                              make_sw128_mnmajor_desc(smem_u32(sV + vs * TILE_BYTES) + voff, SLAB_BYTES), idesc_o,
                              (j | k) != 0 ? 1u : 0u);
                }
                umma_commit(&v_empty[vs]);  // V_j (and P_j) consumed
                umma_commit(o_done);
            }
        }
    } else {
        // ---------------- softmax / correction / epilogue: thread r <-> query row q0 + r ----------------
        const int quad = warp & 3;
        const int r = quad * 32 + lane;
        const int qrow = q0 + r;
        const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
        float m_run = -INFINITY, l_run = 0.f;
        for (int j = 0; j < n_tiles; ++j) {
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            const int key0 = j * TC_BN;
            // one pass: the scaled, masked score row lives in registers (the d=64 variant runs 2 CTAs/SM under a 168-
            // register cap and spills part of it to L1-resident local memory; a second tcgen05.ld pass over S instead
            // was measured 30% slower: profiles/r1e_attn_tc_v2_twopass_2cta_SLOWER.txt)
            float sc[TC_BN];  // RAW scores (masked -> -inf); the softmax scale is folded into the exponent's FMA below
            float mx = -INFINITY;
            // interior tiles need no per-key mask (CTA-uniform test: every key valid for every query row of this CTA)
            const bool full_tile = (key0 + TC_BN <= len) && (!CAUSAL || key0 + TC_BN - 1 <= q0);
#pragma unroll
            for (int c = 0; c < TC_BN / 32; ++c) {
                uint32_t v[32];
                __syncwarp();
                tmem_ld_32x32(tmem_S + lane_addr + c * 32, v);
                tmem_ld_wait();
                if (full_tile) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        sc[c * 32 + i] = __uint_as_float(v[i]);
                        mx = fmaxf(mx, sc[c * 32 + i]);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int key = key0 + c * 32 + i;
                        bool ok = key < len;
                        if (CAUSAL) ok = ok && (key <= qrow);
                        const float x = ok ? __uint_as_float(v[i]) : -INFINITY;
                        sc[c * 32 + i] = x;
                        mx = fmaxf(mx, x);
                    }
                }
            }
            mx *= p.scale_log2;  // scale > 0: max commutes with the scaling (-inf stays -inf)
            // lazy rescale: keep the stale max while the new one is < 2^8 above it (P <= 256: exact enough in bf16 / fp32;
            // the stale max cancels in O / l)
            const bool grow = mx > m_run + 8.0f;  // also true on the first valid tile (m_run = -inf)
            const float m_new = grow ? mx : m_run;
            const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
            const float corr = grow ? exp2f(m_run - m_use) : 1.0f;  // m_run = -inf -> 0
            // rescale the running output (in TMEM) once the previous P·V has landed
            if (j > 0) {
                mbar_wait(o_done, (j - 1) & 1);
                tc_fence_after();
                if (__any_sync(0xffffffffu, corr != 1.0f)) {
#pragma unroll
                    for (int c = 0; c < D / 32; ++c) {
                        uint32_t v[32];
                        tmem_ld_32x32(tmem_O + lane_addr + c * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * corr);
                        tmem_st_32x32(tmem_O + lane_addr + c * 32, v);
                    }
                    tmem_st_wait();
                }
            }
            l_run *= corr;
            m_run = m_new;
            // P = exp2(s - m) as bf16 into the swizzled K-major operand tile: row r, 16-byte chunk cc of slab sl
            float rs = 0.f;
#pragma unroll
            for (int ch = 0; ch < TC_BN / 8; ++ch) {
                float pv[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float arg = fmaf(sc[ch * 8 + e], p.scale_log2, -m_use);
                    pv[e] = ((e & 1) && p.exp_poly) ? ex2_poly(arg) : ex2_approx(arg);  // -inf -> +0 (MUFU) / 2^-126 (poly)
                    rs += pv[e];
                }
                const int sl = ch >> 3, cc = ch & 7;
                *reinterpret_cast<uint4*>(sP + sl * (TC_BM * 128) + r * 128 + ((cc ^ (r & 7)) << 4)) =
                    make_uint4(pack_bf16(pv[0], pv[1]), pack_bf16(pv[2], pv[3]), pack_bf16(pv[4], pv[5]),
                               pack_bf16(pv[6], pv[7]));
            }
            l_run += rs;
            fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        // ---- epilogue: O / l ----
        mbar_wait(o_done, (n_tiles - 1) & 1);
        tc_fence_after();
        const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
        __nv_bfloat16* orow = p.o + b * p.o_bs + (int64_t)qrow * p.o_ts + head * p.o_hs;
#pragma unroll
        for (int c = 0; c < D / 32; ++c) {
            uint32_t v[32];
            __syncwarp();
            tmem_ld_32x32(tmem_O + lane_addr + c * 32, v);
            tmem_ld_wait();
            if (qrow < p.S) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    *reinterpret_cast<uint4*>(orow + c * 32 + i * 8) = make_uint4(
                        pack_bf16(__uint_as_float(v[i * 8 + 0]) * inv, __uint_as_float(v[i * 8 + 1]) * inv),
                        pack_bf16(__uint_as_float(v[i * 8 + 2]) * inv, __uint_as_float(v[i * 8 + 3]) * inv),
                        pack_bf16(__uint_as_float(v[i * 8 + 4]) * inv, __uint_as_float(v[i * 8 + 5]) * inv),
                        pack_bf16(__uint_as_float(v[i * 8 + 6]) * inv, __uint_as_float(v[i * 8 + 7]) * inv));
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

// 4-D view (d, t, h, b) of a bf16 tensor with element strides (1, ts, hs, bs); box = [64, 128, 1, 1], SWIZZLE_128B
int make_tmap_4d(CUtensorMap* map, const void* ptr, int D, int S, int H, int B, int64_t ts, int64_t hs, int64_t bs) {
    static PFN_encodeTiled fn = nullptr;
    if (fn == nullptr) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess || f == nullptr) {
            set_error("cuTensorMapEncodeTiled entry point unavailable");
            return -2;
        }
        fn = reinterpret_cast<PFN_encodeTiled>(f);
    }
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ts * 2) % 16 != 0 || (hs * 2) % 16 != 0 || (bs * 2) % 16 != 0) {
        set_error("flash_attn(tc): operands must be 16B aligned with 16B-multiple strides");
        return -1;
    }
    cuuint64_t dims[4] = {(cuuint64_t)D, (cuuint64_t)S, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)(ts * 2), (cuuint64_t)(hs * 2), (cuuint64_t)(bs * 2)};
    cuuint32_t box[4] = {64, (cuuint32_t)TC_BN, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(4d) failed (%d): D=%d S=%d H=%d B=%d ts=%lld hs=%lld bs=%lld", (int)r, D, S, H,
                  B, (long long)ts, (long long)hs, (long long)bs);
        return -2;
    }
    return 0;
}

template <int D, bool CAUSAL, int MIN_CTAS>
int launch_tc(const FlashArgs& a, cudaStream_t stream) {
    CUtensorMap tq, tk, tv;
    B2_TRY(make_tmap_4d(&tq, a.q, D, a.S, a.H, a.B, a.q_ts, a.q_hs, a.q_bs));
    B2_TRY(make_tmap_4d(&tk, a.k, D, a.S, a.H, a.B, a.k_ts, a.k_hs, a.k_bs));
    B2_TRY(make_tmap_4d(&tv, a.v, D, a.S, a.H, a.B, a.v_ts, a.v_hs, a.v_bs));
    FlashTcParams p;
    p.seq_lens = a.seq_lens;
    p.o = reinterpret_cast<__nv_bfloat16*>(a.o);
    p.o_bs = a.o_bs; p.o_ts = a.o_ts; p.o_hs = a.o_hs;
    p.S = a.S;
    p.scale_log2 = a.scale * 1.4426950408889634f;
    {
        // measured on B200 (profiles/r2g_attn_bench.txt): 3 % SLOWER at every shape (ViT B=32: 220 vs 213 us) — the softmax
        // warps are not MUFU-bound after all at two CTAs per SM; kept as a knob, off by default
        const char* e = getenv("B2_FLASH_EXP_POLY");
        p.exp_poly = (e != nullptr && e[0] == '1');
    }
    constexpr int smem = TcCfg<D>::SMEM;
    static bool attr_set = false;
    auto kern = flash_tc_kernel<D, CAUSAL, MIN_CTAS>;
    if (!attr_set) {
        B2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    dim3 grid((a.S + TC_BM - 1) / TC_BM, a.H, a.B);
    B2_CUDA_CHECK(launch_pdl(kern, grid, dim3(TC_THREADS), (size_t)smem, stream, tq, tk, tv, p));
    B2_LAUNCH_CHECK();
    return 0;
}

}  // namespace

int flash_attn_tc_bf16(const FlashArgs& a, cudaStream_t stream) {
    B2_CHECK_ARG(a.D == 64 || a.D == 128, "flash_attn(tc): head_dim must be 64 or 128 (got %d)", a.D);
    B2_CHECK_ARG((a.o_ts % 8) == 0 && (a.o_hs % 8) == 0 && (a.o_bs % 8) == 0 && (reinterpret_cast<uintptr_t>(a.o) & 15) == 0,
                 "flash_attn(tc): output must be 16B aligned with 16B-multiple strides");
    if (a.D == 64) {
        // d=64: the smem/TMEM/register budget admits two CTAs per SM (168-register cap, part of the score row spills):
        // one CTA's exp2-bound softmax overlaps the other's MMAs, +26% at B=32 (224.6 vs 178.6 TFLOP/s), but a grid that
        // does not even fill the SMs once (B=1: 80 CTAs) only pays for the spills (21.2 vs 16.8 us) -> one-CTA build.
        // B2_FLASH_TC_CTAS=1|2 forces a build (scripts/attn_bench.py; profiles/r1e_attn_tc_v3_lazy_rescale_1v2cta.txt).
        const char* e = getenv("B2_FLASH_TC_CTAS");
        const long long ctas = (long long)((a.S + TC_BM - 1) / TC_BM) * a.H * a.B;
        const bool one = e != nullptr ? e[0] == '1' : ctas <= num_sms();
        if (one)
            return a.causal ? launch_tc<64, true, 1>(a, stream) : launch_tc<64, false, 1>(a, stream);
        return a.causal ? launch_tc<64, true, 2>(a, stream) : launch_tc<64, false, 2>(a, stream);
    }
    return a.causal ? launch_tc<128, true, 1>(a, stream) : launch_tc<128, false, 1>(a, stream);
}

}  // namespace b2
