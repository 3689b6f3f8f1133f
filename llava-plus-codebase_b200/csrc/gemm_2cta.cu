// CTA-pair (cta_group::2) bf16 GEMM for the big-M launches (ViT at B >= 16, LLaMA prefill at B >= 8):
//     C[M,N] = epilogue(A[M,K] · W[N,K]^T)          one 256 x 256 output tile per CTA PAIR
//
// Selected by gemm_bf16's cost model (gemm_tcgen05.cu) wherever its pair tiles fill their waves; first run on a B200 in
// round 2 (tests/test_ops_gpu.py::test_gemm_2cta_*), measured in profiles/r2b_gemm_sweep.json.
//
// Why: the 1-CTA kernel (gemm_tcgen05.cu) reads 96 B/clk of operands from shared memory at its best tile (128x256) — every
// CTA stages the whole 256-row B tile — and tops out at 75-90 % of cuBLAS on these shapes (profiles/r1e_gemm_sweep_bn192.json).
// In a CTA pair each CTA stages HALF of the B tile (128 of the 256 W rows) plus its own 128 A rows, and one
// tcgen05.mma.cta_group::2 (M = 256) issued by the leader CTA drives the tensor cores of both SMs against the union of the two
// shared memories: 64 B/clk per SM, and half the L2->SM traffic for B.
//
// Protocol (ranks 0 = leader, 1 = peer; the two CTAs of a cluster land on the two SMs of a TPC):
//   TMA producer (1 thread per CTA)   waits its own empty[s]; the leader arms ITS full[s] with the bytes of BOTH CTAs; both CTAs
//                                     issue cp.async.bulk.tensor...cta_group::2 whose mbarrier operand is the LEADER's full[s]
//   MMA issuer   (1 thread, leader)   waits full[s] and tmem_empty[a]; 4 x tcgen05.mma.cta_group::2.kind::f16 (256x256x16) per
//                                     k-block; tcgen05.commit ... multicast::cluster (mask 0b11) onto empty[s] / tmem_full[a] of
//                                     both CTAs
//   epilogue     (8 warps per CTA)    each CTA drains its own 128 TMEM lanes (its 128 rows of the tile) with the same fused
//                                     epilogues as the 1-CTA kernel, then arrives on the LEADER's tmem_empty[a] (remote arrive
//                                     from the peer)
#include <cuda.h>
#include <math.h>

#include "common.cuh"
#include "kernels.h"

namespace b2 {

int make_tmap_bf16(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows);

namespace {

constexpr int P_BM = 128;            // rows of A / of W staged per CTA
constexpr int P_TM = 2 * P_BM;       // pair tile M
constexpr int P_TN = 2 * P_BM;       // pair tile N
constexpr int P_BK = 64;
constexpr int P_THREADS = 320;       // warp 0: TMA, warp 1: MMA (leader only), warps 2..9: epilogue
constexpr int P_EPI_WARPS = 8;
constexpr int P_TILE_BYTES = P_BM * P_BK * 2;      // 16 KB
constexpr int P_STAGE_BYTES = 2 * P_TILE_BYTES;    // A half + B half per CTA
constexpr int P_STAGES = 6;
constexpr int P_TMEM_COLS = 512;                   // two 256-column accumulator stages
constexpr int P_SMEM = P_STAGES * P_STAGE_BYTES + 1024 + 256;

struct PairEpi {
    const __nv_bfloat16* bias;
    const __nv_bfloat16* residual;
    void* out;
    int ld_out, ld_res, out_fp32;
    // ACT_ROPE_QKV
    const uint32_t* rope_tab;
    __nv_bfloat16* kcache;
    __nv_bfloat16* vcache;
    int rope_S, rope_H, rope_Smax;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` (a shared::cta address of THIS CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// 2-SM TMA load: destination in this CTA's shared memory, completion bytes on an mbarrier that may live in the peer CTA
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const void* tmap, uint32_t bar_cluster_addr, int32_t c0, int32_t c1,
                                                uint64_t cache_hint) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "l"(cache_hint)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive (once all MMAs issued so far have retired) on the barrier at this shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
    const uint16_t mask = 0b11;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}

__device__ __forceinline__ float p_quick_gelu(float x) { return __fdividef(x, 1.0f + __expf(-1.702f * x)); }
__device__ __forceinline__ float p_gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float p_silu(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
// q*cos + rotate_half(q)*sin with bf16 rounding of each product and of the sum (attention.cu rope_apply: same expression)
__device__ __forceinline__ float p_rope(float x, float partner_signed, float c, float s) {
    return round_bf16(round_bf16(x * c) + round_bf16(partner_signed * s));
}

template <int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P_THREADS, 1)
gemm_bf16_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, int M, int N, int K,
                      PairEpi ep) {
    extern __shared__ uint8_t p_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P_STAGES * P_STAGE_BYTES);
    uint64_t* full_bar = bars;                        // [P_STAGES]   used in the leader
    uint64_t* empty_bar = bars + P_STAGES;            // [P_STAGES]   one per CTA
    uint64_t* tmem_full = bars + 2 * P_STAGES;        // [2]          one per CTA
    uint64_t* tmem_empty = bars + 2 * P_STAGES + 2;   // [2]          used in the leader
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * P_STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

    const int num_m = (M + P_TM - 1) / P_TM;
    const int num_n = (N + P_TN - 1) / P_TN;
    const int num_tiles = num_m * num_n;
    const int num_kb = (K + P_BK - 1) / P_BK;
    constexpr int GM = 4;  // pair-tile m-blocks per raster group
    auto tile_coords = [&](int t, int& m_blk, int& n_blk) {
        const int per_group = GM * num_n;
        const int group = t / per_group;
        const int first_m = group * GM;
        const int gsize = min(GM, num_m - first_m);
        const int within = t - group * per_group;
        m_blk = first_m + within % gsize;
        n_blk = within / gsize;
    };

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int s = 0; s < P_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);    // the leader producer's arrive.expect_tx (+ the TMA bytes of both CTAs)
            mbar_init(&empty_bar[s], 1);   // one multicast tcgen05.commit per use
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);                   // one multicast tcgen05.commit
            mbar_init(&tmem_empty[a], 2 * P_EPI_WARPS);    // every epilogue warp of BOTH CTAs
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm(tmem_slot, P_TMEM_COLS);  // collective over the pair: one warp in each CTA
    tc_fence_before();
    cluster_sync_all();  // barriers initialised and TMEM allocated in both CTAs before anyone signals across
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    pdl_wait();  // set-up above overlapped the upstream kernel's tail (programmatic dependent launch); A / residual are its outputs

    if (warp == 0) {
        if (lane == 0) {
            // ===================== TMA producer (both CTAs) =====================
            int stage = 0;
            uint32_t phase = 0;
            for (int t = pair; t < num_tiles; t += num_pairs) {
                int m_blk, n_blk;
                tile_coords(t, m_blk, n_blk);
                const int a_row = m_blk * P_TM + (int)rank * P_BM;  // this CTA's 128 rows of A
                const int b_row = n_blk * P_TN + (int)rank * P_BM;  // this CTA's 128 rows of W (half of the N tile)
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * P_STAGE_BYTES;
                    const uint32_t leader_full = mapa_rank(smem_u32(&full_bar[stage]), 0);
                    if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * P_STAGE_BYTES);
                    tma_load_2d_2sm(sa, &tmap_a, leader_full, kb * P_BK, a_row, kEvictNormal);
                    tma_load_2d_2sm(sa + P_TILE_BYTES, &tmap_b, leader_full, kb * P_BK, b_row, kEvictNormal);
                    if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (leader && lane == 0) {
            // ===================== MMA issuer (leader CTA only) =====================
            constexpr uint32_t idesc = make_idesc_bf16_f32(P_TM, P_TN);
            int stage = 0;
            uint32_t phase = 0;
            int local = 0;
            for (int t = pair; t < num_tiles; t += num_pairs, ++local) {
                const int as = local & 1;
                mbar_wait(&tmem_empty[as], ((local >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * P_TN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * P_STAGE_BYTES);
                    const uint64_t da = make_sw128_kmajor_desc(sa);
                    const uint64_t db = make_sw128_kmajor_desc(sa + P_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < P_BK / 16; ++k)
                        umma_bf16_2sm(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit_2sm(&empty_bar[stage]);
                    if (kb == num_kb - 1) umma_commit_2sm(&tmem_full[as]);
                    if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ===================== epilogue warps 2..9 (both CTAs: own 128 rows of the pair tile) =====================
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        int local = 0;
        for (int t = pair; t < num_tiles; t += num_pairs, ++local) {
            int m_blk, n_blk;
            tile_coords(t, m_blk, n_blk);
            const int as = local & 1;
            mbar_wait(&tmem_full[as], (local >> 1) & 1);
            tc_fence_after();
            const int row = m_blk * P_TM + (int)rank * P_BM + q * 32 + lane;
            const bool row_ok = row < M;
            const uint32_t taddr_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * P_TN;
            if constexpr (ACT == ACT_SWIGLU) {
#pragma unroll 1
                for (int it = half * (P_TN / 128); it < (half + 1) * (P_TN / 128); ++it) {
                    const int g = it >> 1, j = it & 1;
                    uint32_t vg[32], vu[32];
                    __syncwarp();
                    tmem_ld_32x32(taddr_row + g * 128 + j * 32, vg);
                    tmem_ld_32x32(taddr_row + g * 128 + 64 + j * 32, vu);
                    tmem_ld_wait();
                    const int ocol0 = (n_blk * P_TN) / 2 + g * 64 + j * 32;
                    if (row_ok && ocol0 < N / 2) {
                        __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(ep.out) + (size_t)row * ep.ld_out + ocol0;
#pragma unroll
                        for (int v8 = 0; v8 < 4; ++v8) {
                            uint32_t pk[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float g0 = __uint_as_float(vg[v8 * 8 + 2 * e]), g1 = __uint_as_float(vg[v8 * 8 + 2 * e + 1]);
                                const float u0 = __uint_as_float(vu[v8 * 8 + 2 * e]), u1 = __uint_as_float(vu[v8 * 8 + 2 * e + 1]);
                                pk[e] = pack_bf16(p_silu(g0) * u0, p_silu(g1) * u1);
                            }
                            *reinterpret_cast<uint4*>(op + v8 * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        }
                    }
                }
            } else if constexpr (ACT == ACT_ROPE_QKV) {
                // This warp's half of the pair tile is the 128 columns of ONE head of q, k or v (hidden % 256 == 0, head_dim 128).
                // Element i of the head rotates with element i + 64: the two 32-column chunks j and j + 2 are read together.
                const int col_base = n_blk * P_TN + half * 128;
                const int hdim = ep.rope_H * 128;
                const int part = col_base / hdim;  // 0 = q, 1 = k, 2 = v
                const int head = (col_base - part * hdim) >> 7;
                const int b = row / ep.rope_S, tpos = row - b * ep.rope_S;
#pragma unroll 1
                for (int j = 0; j < 2; ++j) {
                    uint32_t lo[32], hi[32];
                    __syncwarp();
                    tmem_ld_32x32(taddr_row + half * 128 + j * 32, lo);
                    tmem_ld_32x32(taddr_row + half * 128 + 64 + j * 32, hi);
                    tmem_ld_wait();
                    if (!(row_ok && col_base < N)) continue;
                    __nv_bfloat16* dst;
                    if (part == 0) dst = reinterpret_cast<__nv_bfloat16*>(ep.out) + (size_t)row * ep.ld_out + col_base + j * 32;
                    else dst = (part == 1 ? ep.kcache : ep.vcache) + (((size_t)b * ep.rope_H + head) * ep.rope_Smax + tpos) * 128 + j * 32;
                    const uint4* tab = reinterpret_cast<const uint4*>(ep.rope_tab + (size_t)tpos * 64 + j * 32);  // 8 x 16 B: 32 (cos, sin) pairs
#pragma unroll
                    for (int v8 = 0; v8 < 4; ++v8) {
                        float ol[8], oh[8];
                        uint32_t csw[8];
                        if (part < 2) {
                            const uint4 t0 = __ldg(tab + v8 * 2), t1 = __ldg(tab + v8 * 2 + 1);
                            csw[0] = t0.x; csw[1] = t0.y; csw[2] = t0.z; csw[3] = t0.w;
                            csw[4] = t1.x; csw[5] = t1.y; csw[6] = t1.z; csw[7] = t1.w;
                        }
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float xl = round_bf16(__uint_as_float(lo[v8 * 8 + e]));  // the projection as bf16, like the unfused path
                            const float xh = round_bf16(__uint_as_float(hi[v8 * 8 + e]));
                            if (part < 2) {
                                const float cs_c = bf16_lo(csw[e]), cs_s = bf16_hi(csw[e]);
                                ol[e] = p_rope(xl, -xh, cs_c, cs_s);
                                oh[e] = p_rope(xh, xl, cs_c, cs_s);
                            } else {
                                ol[e] = xl; oh[e] = xh;
                            }
                        }
                        *reinterpret_cast<uint4*>(dst + v8 * 8) =
                            make_uint4(pack_bf16(ol[0], ol[1]), pack_bf16(ol[2], ol[3]), pack_bf16(ol[4], ol[5]), pack_bf16(ol[6], ol[7]));
                        *reinterpret_cast<uint4*>(dst + 64 + v8 * 8) =
                            make_uint4(pack_bf16(oh[0], oh[1]), pack_bf16(oh[2], oh[3]), pack_bf16(oh[4], oh[5]), pack_bf16(oh[6], oh[7]));
                    }
                }
            } else {
#pragma unroll 1
                for (int c = half * (P_TN / 64); c < (half + 1) * (P_TN / 64); ++c) {
                    uint32_t v[32];
                    __syncwarp();
                    tmem_ld_32x32(taddr_row + c * 32, v);
                    tmem_ld_wait();
                    const int col0 = n_blk * P_TN + c * 32;
#pragma unroll
                    for (int v8 = 0; v8 < 4; ++v8) {
                        const int col = col0 + v8 * 8;
                        if (!(row_ok && col < N)) continue;
                        float x[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) x[e] = __uint_as_float(v[v8 * 8 + e]);
                        if (ep.bias != nullptr) {
                            const uint4 b = *reinterpret_cast<const uint4*>(ep.bias + col);
                            x[0] += bf16_lo(b.x); x[1] += bf16_hi(b.x); x[2] += bf16_lo(b.y); x[3] += bf16_hi(b.y);
                            x[4] += bf16_lo(b.z); x[5] += bf16_hi(b.z); x[6] += bf16_lo(b.w); x[7] += bf16_hi(b.w);
                        }
                        if constexpr (ACT == ACT_QUICK_GELU) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) x[e] = p_quick_gelu(x[e]);
                        } else if constexpr (ACT == ACT_GELU_ERF) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) x[e] = p_gelu_erf(x[e]);
                        }
                        if (ep.residual != nullptr) {
                            const uint4 r = *reinterpret_cast<const uint4*>(ep.residual + (size_t)row * ep.ld_res + col);
                            x[0] += bf16_lo(r.x); x[1] += bf16_hi(r.x); x[2] += bf16_lo(r.y); x[3] += bf16_hi(r.y);
                            x[4] += bf16_lo(r.z); x[5] += bf16_hi(r.z); x[6] += bf16_lo(r.w); x[7] += bf16_hi(r.w);
                        }
                        if (ep.out_fp32) {
                            float* op = reinterpret_cast<float*>(ep.out) + (size_t)row * ep.ld_out + col;
                            *reinterpret_cast<float4*>(op) = make_float4(x[0], x[1], x[2], x[3]);
                            *reinterpret_cast<float4*>(op + 4) = make_float4(x[4], x[5], x[6], x[7]);
                        } else {
                            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(ep.out) + (size_t)row * ep.ld_out + col;
                            *reinterpret_cast<uint4*>(op) = make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]),
                                                                       pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
                        }
                    }
                }
            }
            // this warp has read its part of accumulator stage `as`: tell the leader's MMA issuer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_rank(smem_u32(&tmem_empty[as]), 0));
        }
    }

    __syncwarp();  // lanes 1..31 of the producer / MMA warps waited here: the cluster barrier below is .aligned
    tc_fence_before();
    cluster_sync_all();  // no CTA may free TMEM / exit while its partner can still signal into its shared memory
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, P_TMEM_COLS);
    }
}

template <int ACT>
int launch_pair(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const PairEpi& ep, cudaStream_t stream) {
    static bool attr_set = false;
    auto kern = gemm_bf16_2cta_kernel<ACT>;
    if (!attr_set) {
        B2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM));
        attr_set = true;
    }
    const int tiles = ((M + P_TM - 1) / P_TM) * ((N + P_TN - 1) / P_TN);
    const int max_pairs = num_sms() / 2;
    const int pairs = tiles < max_pairs ? tiles : max_pairs;
    B2_CUDA_CHECK(launch_pdl(kern, dim3(2 * pairs), dim3(P_THREADS), (size_t)P_SMEM, stream, ta, tb, M, N, K, ep));  // cluster dims (2,1,1) are a kernel attribute
    B2_LAUNCH_CHECK();
    return 0;
}

}  // namespace

// Same contract as gemm_bf16 (kernels.h). SwiGLU needs N % 128 == 0.
int gemm_bf16_2cta(const GemmArgs& g, cudaStream_t stream) {
    B2_CHECK_ARG(g.M > 0 && g.N > 0 && g.K > 0, "gemm_2cta: empty problem M=%d N=%d K=%d", g.M, g.N, g.K);
    B2_CHECK_ARG(g.K % 8 == 0 && g.N % 8 == 0, "gemm_2cta: K and N must be multiples of 8 (K=%d N=%d)", g.K, g.N);
    B2_CHECK_ARG(g.act != ACT_SWIGLU || (g.N % 128 == 0 && !g.out_fp32 && g.bias == nullptr && g.residual == nullptr),
                 "gemm_2cta: swiglu needs N %% 128 == 0, no bias/residual, bf16 output");
    B2_CHECK_ARG((reinterpret_cast<uintptr_t>(g.out) & 15) == 0 && (g.ld_out % 8) == 0, "gemm_2cta: out alignment");
    B2_CHECK_ARG(g.residual == nullptr || ((reinterpret_cast<uintptr_t>(g.residual) & 15) == 0 && (g.ld_res % 8) == 0),
                 "gemm_2cta: residual alignment");
    CUtensorMap ta, tb;
    B2_TRY(make_tmap_bf16(&ta, g.A, g.M, g.K, g.lda, P_BM));
    B2_TRY(make_tmap_bf16(&tb, g.W, g.N, g.K, g.ldw, P_BM));
    PairEpi ep;
    ep.bias = reinterpret_cast<const __nv_bfloat16*>(g.bias);
    ep.residual = reinterpret_cast<const __nv_bfloat16*>(g.residual);
    ep.out = g.out; ep.ld_out = g.ld_out; ep.ld_res = g.ld_res; ep.out_fp32 = g.out_fp32;
    ep.rope_tab = reinterpret_cast<const uint32_t*>(g.rope.table);
    ep.kcache = reinterpret_cast<__nv_bfloat16*>(g.rope.kcache); ep.vcache = reinterpret_cast<__nv_bfloat16*>(g.rope.vcache);
    ep.rope_S = g.rope.S; ep.rope_H = g.rope.H; ep.rope_Smax = g.rope.Smax;
    if (g.act == ACT_ROPE_QKV) {
        B2_CHECK_ARG(g.rope.table && g.rope.kcache && g.rope.vcache && g.rope.S > 0 && g.rope.H > 0 && g.rope.Smax >= g.rope.S,
                     "gemm_2cta(rope_qkv): rope arguments missing");
        B2_CHECK_ARG(g.N == 3 * g.rope.H * 128 && (g.rope.H * 128) % 256 == 0 && g.M % g.rope.S == 0 && !g.out_fp32 &&
                         g.bias == nullptr && g.residual == nullptr,
                     "gemm_2cta(rope_qkv): needs N = 3*H*128 with H*128 %% 256 == 0, M = B*S, bf16 output, no bias/residual");
    }
    switch (g.act) {
        case ACT_NONE: return launch_pair<ACT_NONE>(ta, tb, g.M, g.N, g.K, ep, stream);
        case ACT_QUICK_GELU: return launch_pair<ACT_QUICK_GELU>(ta, tb, g.M, g.N, g.K, ep, stream);
        case ACT_GELU_ERF: return launch_pair<ACT_GELU_ERF>(ta, tb, g.M, g.N, g.K, ep, stream);
        case ACT_SWIGLU: return launch_pair<ACT_SWIGLU>(ta, tb, g.M, g.N, g.K, ep, stream);
        case ACT_ROPE_QKV: return launch_pair<ACT_ROPE_QKV>(ta, tb, g.M, g.N, g.K, ep, stream);
        default: break;
    }
    set_error("gemm_2cta: unsupported activation %d", g.act);
    return -1;
}

}  // namespace b2
