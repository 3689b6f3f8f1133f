// Persistent warp-specialised bf16 GEMM for sm_100a:  C[M,N] = epilogue(A[M,K] · W[N,K]^T)
//
//   * A (activations, row-major, K contiguous) and W (nn.Linear weight [out,in], K contiguous) are both
//     "K-major" operands: TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) stages 128x64 / BNx64 bf16 tiles
//     into shared memory, one elected thread issues tcgen05.mma (UMMA 128xBNx16, fp32 accumulate in TMEM),
//     four epilogue warps drain TMEM with tcgen05.ld and apply the fused epilogue.
//   * Three pipelines: smem full/empty ring (TMA <-> MMA), TMEM full/empty double buffer (MMA <-> epilogue),
//     static persistent tile scheduler (grid = #SMs) with grouped rasterisation for L2 reuse.
//   * Fused epilogues replace the separate bias / activation / residual / silu*mul kernels of the
//     reference's HF path (transformers modeling_clip.py:347-351 CLIPMLP, modeling_llama.py:182-184
//     LlamaMLP, :303-332 residual adds; llava/model/multimodal_projector/builder.py:42-46).
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace b2 {

static int g_num_sms = 0;
int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) return nullptr;
        fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// 2D bf16 row-major [rows, cols] with leading dimension ld (elements); box = [box_rows, 64]. Shared with gemm_skinny.cu.
int make_tmap_bf16(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    PFN_encodeTiled fn = get_encode_fn();
    if (fn == nullptr) {
        set_error("cuTensorMapEncodeTiled entry point unavailable");
        return -2;
    }
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld * 2) % 16 != 0) {
        set_error("TMA operand must be 16B aligned with a 16B-multiple row pitch (ptr=%p ld=%lld)", ptr,
                  (long long)ld);
        return -1;
    }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)(ld * 2)};
    cuuint32_t box[2] = {64u /* bf16 = one 128 B swizzle row */, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)r, (long long)rows,
                  (long long)cols, (long long)ld);
        return -2;
    }
    return 0;
}

// 2D byte tensor (fp8 operands) row-major [rows, cols] with row pitch ld bytes; box = [box_rows, 128 B]. gemm_skinny.cu.
int make_tmap_u8(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    PFN_encodeTiled fn = get_encode_fn();
    if (fn == nullptr) {
        set_error("cuTensorMapEncodeTiled entry point unavailable");
        return -2;
    }
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || ld % 16 != 0) {
        set_error("TMA operand must be 16B aligned with a 16B-multiple row pitch (ptr=%p ld=%lld)", ptr, (long long)ld);
        return -1;
    }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld};
    cuuint32_t box[2] = {128u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(u8) failed (%d) rows=%lld cols=%lld ld=%lld", (int)r, (long long)rows,
                  (long long)cols, (long long)ld);
        return -2;
    }
    return 0;
}

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int kNumThreads = 320;   // warp 0: TMA, warp 1: MMA, warps 2..9: epilogue (two per TMEM lane quadrant)
constexpr int kNumEpiWarps = 8;
constexpr int A_TILE_BYTES = BM * BK * 2;

template <int BN>
struct GemmCfg {
    static constexpr int B_TILE_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
    static constexpr int STAGES = (BN == 256) ? 4 : (BN == 192 ? 5 : (BN == 128 ? 7 : 9));
    static constexpr int ACC_STRIDE = (BN == 192) ? 256 : BN;      // TMEM columns between the two accumulator stages
    static constexpr int TMEM_COLS = 2 * ACC_STRIDE;               // power of two (128 / 256 / 512)
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct EpiParams {
    const __nv_bfloat16* bias;      // [N] or nullptr
    const __nv_bfloat16* residual;  // [M, ld_res] or nullptr (may alias out)
    void* out;                      // bf16 [M, ld_out] or fp32 [M, ld_out]
    int ld_out;
    int ld_res;
    int out_fp32;
};

__device__ __forceinline__ float act_quick_gelu(float x) { return __fdividef(x, 1.0f + __expf(-1.702f * x)); }
__device__ __forceinline__ float act_gelu_erf(float x) {
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float act_silu(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

template <int BN, int ACT>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a,
                         const __grid_constant__ CUtensorMap tmap_b, int M, int N, int K,
                         EpiParams ep) {
    using Cfg = GemmCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;

    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* full_bar = bars;                    // [STAGES]
    uint64_t* empty_bar = bars + STAGES;          // [STAGES]
    uint64_t* tmem_full = bars + 2 * STAGES;      // [2]
    uint64_t* tmem_empty = bars + 2 * STAGES + 2; // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int num_m = (M + BM - 1) / BM;
    const int num_n = (N + BN - 1) / BN;
    const int num_tiles = num_m * num_n;
    const int num_kb = (K + BK - 1) / BK;
    constexpr int GM = 8;  // m-blocks per raster group

    auto tile_coords = [&](int t, int& m_blk, int& n_blk) {
        const int tiles_per_group = GM * num_n;
        const int group = t / tiles_per_group;
        const int first_m = group * GM;
        const int gsize = min(GM, num_m - first_m);
        const int within = t - group * tiles_per_group;
        m_blk = first_m + within % gsize;
        n_blk = within / gsize;
    };

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], kNumEpiWarps);  // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    pdl_trigger();
    if (warp == 0) {
        // ===================== TMA producer (one thread) =====================
        if (lane == 0) {
            // Programmatic dependent launch: the WEIGHT halves of the first ring-full of stages are requested before the
            // upstream kernel (which produces A) is known to have finished; the A halves follow after pdl_wait().
            int stage = 0;
            uint32_t phase = 0;
            int pre_n = 0;
            if ((int)blockIdx.x < num_tiles) {
                int m_blk, n_blk;
                tile_coords(blockIdx.x, m_blk, n_blk);
                for (int kb = 0; kb < num_kb && pre_n < STAGES; ++kb, ++pre_n) {
                    uint8_t* sb = smem + pre_n * Cfg::STAGE_BYTES + A_TILE_BYTES;
                    mbar_arrive_expect_tx(&full_bar[pre_n], Cfg::STAGE_BYTES);
                    tma_load_2d(sb, &tmap_b, &full_bar[pre_n], kb * BK, n_blk * BN, kEvictNormal);
                }
            }
            pdl_wait();
            int seen = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                int m_blk, n_blk;
                tile_coords(t, m_blk, n_blk);
                for (int kb = 0; kb < num_kb; ++kb, ++seen) {
                    uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                    uint8_t* sb = sa + A_TILE_BYTES;
                    if (seen >= pre_n) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                        tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BK, n_blk * BN, kEvictNormal);
                    }
                    tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, m_blk * BM, kEvictNormal);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16_f32(BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int local = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++local) {
                const int as = local & 1;
                const uint32_t aphase = (local >> 1) & 1;
                mbar_wait(&tmem_empty[as], aphase ^ 1);  // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * Cfg::ACC_STRIDE;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);  // TMA bytes have landed
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint64_t da = make_sw128_kmajor_desc(sa);
                    const uint64_t db = make_sw128_kmajor_desc(sa + A_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        // advance 16 bf16 = 32 B along K inside the 128B swizzle atom: +2 in the addr field
                        umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
                    if (kb == num_kb - 1) umma_commit(&tmem_full[as]);  // accumulator complete
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ===================== epilogue warps 2..9 =====================
        pdl_wait();                        // residual / out may be read or written by the upstream kernel
        const int q = warp & 3;            // TMEM lane quadrant this warp may access (warp_id % 4)
        const int half = (warp - 2) >> 2;  // the two warps of a quadrant split the tile's columns
        int local = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++local) {
            int m_blk, n_blk;
            tile_coords(t, m_blk, n_blk);
            const int as = local & 1;
            const uint32_t aphase = (local >> 1) & 1;
            mbar_wait(&tmem_full[as], aphase);
            tc_fence_after();
            const int row = m_blk * BM + q * 32 + lane;
            const bool row_ok = row < M;
            const uint32_t taddr_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * Cfg::ACC_STRIDE;

            if constexpr (ACT == ACT_SWIGLU) {
                // W rows are block-interleaved: within each 128-col group, cols [0,64) = gate, [64,128) = up
                // for the same 64 output channels. out is [M, N/2].
                static_assert(BN % 128 == 0, "swiglu needs BN multiple of 128");
#pragma unroll 1
                for (int it = half * (BN / 128); it < (half + 1) * (BN / 128); ++it) {
                    {
                        const int g = it >> 1, j = it & 1;
                        uint32_t vg[32], vu[32];
                        __syncwarp();
                        tmem_ld_32x32(taddr_row + g * 128 + j * 32, vg);
                        tmem_ld_32x32(taddr_row + g * 128 + 64 + j * 32, vu);
                        tmem_ld_wait();
                        const int ocol0 = (n_blk * BN) / 2 + g * 64 + j * 32;
                        if (row_ok && ocol0 < N / 2) {
                            __nv_bfloat16* op =
                                reinterpret_cast<__nv_bfloat16*>(ep.out) + (size_t)row * ep.ld_out + ocol0;
#pragma unroll
                            for (int v8 = 0; v8 < 4; ++v8) {
                                uint32_t pk[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    float g0 = __uint_as_float(vg[v8 * 8 + 2 * e]);
                                    float g1 = __uint_as_float(vg[v8 * 8 + 2 * e + 1]);
                                    float u0 = __uint_as_float(vu[v8 * 8 + 2 * e]);
                                    float u1 = __uint_as_float(vu[v8 * 8 + 2 * e + 1]);
                                    pk[e] = pack_bf16(act_silu(g0) * u0, act_silu(g1) * u1);
                                }
                                *reinterpret_cast<uint4*>(op + v8 * 8) =
                                    make_uint4(pk[0], pk[1], pk[2], pk[3]);
                            }
                        }
                    }
                }
            } else {
#pragma unroll 1
                for (int c = half * (BN / 64); c < (half + 1) * (BN / 64); ++c) {
                    uint32_t v[32];
                    __syncwarp();
                    tmem_ld_32x32(taddr_row + c * 32, v);
                    tmem_ld_wait();
                    const int col0 = n_blk * BN + c * 32;
#pragma unroll
                    for (int v8 = 0; v8 < 4; ++v8) {
                        const int col = col0 + v8 * 8;
                        // N % 8 == 0 is enforced on the host; no early-outs here: the tcgen05.ld above is
                        // .sync.aligned, so the warp must stay converged across loop iterations
                        if (!(row_ok && col < N)) continue;
                        float x[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) x[e] = __uint_as_float(v[v8 * 8 + e]);
                        if (ep.bias != nullptr) {
                            const uint4 b = *reinterpret_cast<const uint4*>(ep.bias + col);
                            x[0] += bf16_lo(b.x); x[1] += bf16_hi(b.x);
                            x[2] += bf16_lo(b.y); x[3] += bf16_hi(b.y);
                            x[4] += bf16_lo(b.z); x[5] += bf16_hi(b.z);
                            x[6] += bf16_lo(b.w); x[7] += bf16_hi(b.w);
                        }
                        if constexpr (ACT == ACT_QUICK_GELU) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) x[e] = act_quick_gelu(x[e]);
                        } else if constexpr (ACT == ACT_GELU_ERF) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) x[e] = act_gelu_erf(x[e]);
                        }
                        if (ep.residual != nullptr) {
                            const uint4 r = *reinterpret_cast<const uint4*>(
                                ep.residual + (size_t)row * ep.ld_res + col);
                            x[0] += bf16_lo(r.x); x[1] += bf16_hi(r.x);
                            x[2] += bf16_lo(r.y); x[3] += bf16_hi(r.y);
                            x[4] += bf16_lo(r.z); x[5] += bf16_hi(r.z);
                            x[6] += bf16_lo(r.w); x[7] += bf16_hi(r.w);
                        }
                        if (ep.out_fp32) {
                            float* op = reinterpret_cast<float*>(ep.out) + (size_t)row * ep.ld_out + col;
                            *reinterpret_cast<float4*>(op) = make_float4(x[0], x[1], x[2], x[3]);
                            *reinterpret_cast<float4*>(op + 4) = make_float4(x[4], x[5], x[6], x[7]);
                        } else {
                            __nv_bfloat16* op =
                                reinterpret_cast<__nv_bfloat16*>(ep.out) + (size_t)row * ep.ld_out + col;
                            *reinterpret_cast<uint4*>(op) =
                                make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]),
                                           pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
                        }
                    }
                }
            }
            // all TMEM reads of this accumulator stage are complete (wait::ld above): hand it back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[as]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------
// mm_projector as ONE kernel: out = (gelu_erf(X·W1^T + b1))·W2^T + b2   (reference: nn.Sequential(Linear, GELU, Linear),
// llava/model/multimodal_projector/builder.py:39-46)
// ---------------------------------------------------------------------------------------------
// Same warp-specialised pipeline as gemm_bf16_tcgen05_kernel, but the persistent tile loop walks the tiles of BOTH GEMMs in
// one dependency-ordered sequence:  P1(g0) P1(g1) P2(g0) P1(g2) P2(g1) ... P2(gLast)   (g = group of FP_GM row blocks).
// A phase-2 tile of row block m needs the whole 128 x N1 slab of the intermediate H, i.e. all N1/BN phase-1 tiles of m: their
// epilogue warps publish completion with st.global -> __threadfence -> atomicAdd(row_done[m]); the phase-2 TMA producer
// acquires the counter, crosses from the generic to the async proxy (fence.proxy.async) and only then issues its loads.
// Every dependency points backwards in the sequence and every CTA walks it in increasing order, so the spin cannot
// deadlock (grid <= #SMs, one CTA per SM: all CTAs are resident). H never needs a second launch to become visible and, for
// batches of images, is consumed one group (1024 rows = 8 MB) behind its production, out of L2.
constexpr int FP_GM = 8;

struct FusedProjParams {
    int M, N1, K1, N2;       // K2 == N1
    EpiParams ep1, ep2;      // ep1.out = H (bf16 [M, N1]); ep2.out = result
    int* row_done;           // [ceil(M/128)] completion counters, zeroed (stream-ordered) before the launch
    int target;              // value row_done[m] reaches when every epilogue warp of every phase-1 tile of row block m has arrived
};

struct FpTile { int phase, m_blk, n_blk; };

__device__ __forceinline__ FpTile fp_tile(int t, int num_m, int nn1, int nn2) {
    // sequence: P1(0) | P1(1) P2(0) | P1(2) P2(1) | ... | P2(G-1)
    const int G = (num_m + FP_GM - 1) / FP_GM;
    FpTile r;
    for (int slot = 0; slot < 2 * G; ++slot) {
        int phase, g;
        if (slot == 0) { phase = 1; g = 0; }
        else if (slot == 2 * G - 1) { phase = 2; g = G - 1; }
        else { phase = (slot & 1) ? 1 : 2; g = (slot & 1) ? (slot + 1) / 2 : slot / 2 - 1; }
        const int gs = min(FP_GM, num_m - g * FP_GM);
        const int n = gs * (phase == 1 ? nn1 : nn2);
        if (t < n) {
            r.phase = phase;
            r.m_blk = g * FP_GM + t % gs;
            r.n_blk = t / gs;
            return r;
        }
        t -= n;
    }
    r.phase = 0; r.m_blk = 0; r.n_blk = 0;
    return r;
}

template <int BN>
__global__ void __launch_bounds__(kNumThreads, 1)
projector_fused_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w1,
                       const __grid_constant__ CUtensorMap tmap_h, const __grid_constant__ CUtensorMap tmap_w2,
                       FusedProjParams P) {
    using Cfg = GemmCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + STAGES;
    uint64_t* tmem_full = bars + 2 * STAGES;
    uint64_t* tmem_empty = bars + 2 * STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_m = (P.M + BM - 1) / BM;
    const int nn1 = (P.N1 + BN - 1) / BN, nn2 = (P.N2 + BN - 1) / BN;
    const int total_tiles = num_m * (nn1 + nn2);
    const int nkb1 = (P.K1 + BK - 1) / BK, nkb2 = (P.N1 + BK - 1) / BK;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_x); tma_prefetch_desc(&tmap_w1); tma_prefetch_desc(&tmap_h); tma_prefetch_desc(&tmap_w2);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], kNumEpiWarps); }
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer =====
            pdl_wait();   // X comes from the vision tower's last kernel
            int stage = 0; uint32_t phase = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const FpTile tl = fp_tile(t, num_m, nn1, nn2);
                if (tl.phase == 2) {
                    // all phase-1 tiles of this row block have been stored and fenced by their CTAs
                    unsigned int spins = 0;
                    int seen;
                    do {
                        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(P.row_done + tl.m_blk) : "memory");
                        if (++spins > (1u << 26)) asm volatile("trap;");
                    } while (seen - P.target < 0);
                    asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy stores of H -> async-proxy (TMA) reads
                }
                const CUtensorMap* ta = tl.phase == 1 ? &tmap_x : &tmap_h;
                const CUtensorMap* tb = tl.phase == 1 ? &tmap_w1 : &tmap_w2;
                const int nkb = tl.phase == 1 ? nkb1 : nkb2;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                    tma_load_2d(sa, ta, &full_bar[stage], kb * BK, tl.m_blk * BM, kEvictNormal);
                    tma_load_2d(sa + A_TILE_BYTES, tb, &full_bar[stage], kb * BK, tl.n_blk * BN, kEvictNormal);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ===== MMA issuer =====
            constexpr uint32_t idesc = make_idesc_bf16_f32(BM, BN);
            int stage = 0; uint32_t phase = 0; int local = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++local) {
                const FpTile tl = fp_tile(t, num_m, nn1, nn2);
                const int nkb = tl.phase == 1 ? nkb1 : nkb2;
                const int as = local & 1;
                mbar_wait(&tmem_empty[as], ((local >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * Cfg::ACC_STRIDE;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint64_t da = make_sw128_kmajor_desc(sa);
                    const uint64_t db = make_sw128_kmajor_desc(sa + A_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit(&empty_bar[stage]);
                    if (kb == nkb - 1) umma_commit(&tmem_full[as]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ===== epilogue warps 2..9: bias (+ erf-GELU in phase 1), bf16 stores; phase 1 publishes the row block =====
        pdl_wait();
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        int local = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++local) {
            const FpTile tl = fp_tile(t, num_m, nn1, nn2);
            const EpiParams& ep = tl.phase == 1 ? P.ep1 : P.ep2;
            const int N = tl.phase == 1 ? P.N1 : P.N2;
            const int as = local & 1;
            mbar_wait(&tmem_full[as], (local >> 1) & 1);
            tc_fence_after();
            const int row = tl.m_blk * BM + q * 32 + lane;
            const bool row_ok = row < P.M;
            const uint32_t taddr_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * Cfg::ACC_STRIDE;
#pragma unroll 1
            for (int c = half * (BN / 64); c < (half + 1) * (BN / 64); ++c) {
                uint32_t v[32];
                __syncwarp();
                tmem_ld_32x32(taddr_row + c * 32, v);
                tmem_ld_wait();
                const int col0 = tl.n_blk * BN + c * 32;
#pragma unroll
                for (int v8 = 0; v8 < 4; ++v8) {
                    const int col = col0 + v8 * 8;
                    if (!(row_ok && col < N)) continue;
                    float x[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) x[e] = __uint_as_float(v[v8 * 8 + e]);
                    const uint4 b = *reinterpret_cast<const uint4*>(ep.bias + col);
                    x[0] += bf16_lo(b.x); x[1] += bf16_hi(b.x); x[2] += bf16_lo(b.y); x[3] += bf16_hi(b.y);
                    x[4] += bf16_lo(b.z); x[5] += bf16_hi(b.z); x[6] += bf16_lo(b.w); x[7] += bf16_hi(b.w);
                    if (tl.phase == 1) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) x[e] = act_gelu_erf(x[e]);
                    }
                    __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(ep.out) + (size_t)row * ep.ld_out + col;
                    *reinterpret_cast<uint4*>(op) = make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]),
                                                               pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
                }
            }
            tc_fence_before();
            if (tl.phase == 1) {
                __threadfence();  // this lane's H stores are visible device-wide before the warp's arrival below
                asm volatile("fence.proxy.async;" ::: "memory");
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&tmem_empty[as]);
                if (tl.phase == 1) asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(P.row_done + tl.m_blk) : "memory");
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, Cfg::TMEM_COLS); }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <int BN, int ACT>
int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const EpiParams& ep,
                cudaStream_t stream) {
    using Cfg = GemmCfg<BN>;
    static bool attr_set = false;
    auto kern = gemm_bf16_tcgen05_kernel<BN, ACT>;
    if (!attr_set) {
        B2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_set = true;
    }
    const int num_tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    const int grid = num_tiles < num_sms() ? num_tiles : num_sms();
    B2_CUDA_CHECK(launch_pdl(kern, dim3(grid), dim3(kNumThreads), (size_t)Cfg::SMEM_BYTES, stream, ta, tb, M, N, K, ep));
    B2_LAUNCH_CHECK();
    return 0;
}

template <int BN>
int dispatch_act(int act, const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K,
                 const EpiParams& ep, cudaStream_t stream) {
    switch (act) {
        case ACT_NONE: return launch_gemm<BN, ACT_NONE>(ta, tb, M, N, K, ep, stream);
        case ACT_QUICK_GELU: return launch_gemm<BN, ACT_QUICK_GELU>(ta, tb, M, N, K, ep, stream);
        case ACT_GELU_ERF: return launch_gemm<BN, ACT_GELU_ERF>(ta, tb, M, N, K, ep, stream);
        default: break;
    }
    set_error("gemm: unsupported activation %d", act);
    return -1;
}

template <int BN>
int launch_fused_projector(const CUtensorMap& tx, const CUtensorMap& tw1, const CUtensorMap& th, const CUtensorMap& tw2,
                           const FusedProjParams& P, cudaStream_t stream) {
    using Cfg = GemmCfg<BN>;
    static bool attr_set = false;
    auto kern = projector_fused_kernel<BN>;
    if (!attr_set) {
        B2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_set = true;
    }
    const int num_m = (P.M + BM - 1) / BM;
    const int tiles = num_m * ((P.N1 + BN - 1) / BN + (P.N2 + BN - 1) / BN);
    const int grid = tiles < num_sms() ? tiles : num_sms();  // <= #SMs, 1 CTA/SM: every CTA is resident (the spin needs it)
    B2_CUDA_CHECK(launch_pdl(kern, dim3(grid), dim3(kNumThreads), (size_t)Cfg::SMEM_BYTES, stream, tx, tw1, th, tw2, P));
    B2_LAUNCH_CHECK();
    return 0;
}

}  // namespace

// out[M,N2] = (gelu_erf(X[M,K1]·W1[N1,K1]^T + b1))·W2[N2,N1]^T + b2 in one launch; H [M,N1] bf16 is the caller's scratch,
// row_done the caller's int[ceil(M/128)] scratch (zeroed here, on the stream).
int projector_fused_bf16(const void* X, int ldx, const void* W1, const void* b1, const void* W2, const void* b2, void* H,
                         void* out, int ld_out, int M, int K1, int N1, int N2, int* row_done, cudaStream_t stream) {
    B2_CHECK_ARG(M > 0 && K1 % 8 == 0 && N1 % 8 == 0 && N2 % 8 == 0 && b1 && b2 && row_done,
                 "projector_fused: bad problem M=%d K1=%d N1=%d N2=%d", M, K1, N1, N2);
    const int num_m = (M + BM - 1) / BM;
    // tile width: one wave of 192-wide tiles for a single image (5 row blocks x 22 = 110 tiles per GEMM), 256 otherwise
    const int bn = (num_m * ((N1 + 191) / 192) <= num_sms()) ? 192 : 256;
    CUtensorMap tx, tw1, th, tw2;
    B2_TRY(make_tmap_bf16(&tx, X, M, K1, ldx, BM));
    B2_TRY(make_tmap_bf16(&tw1, W1, N1, K1, K1, bn));
    B2_TRY(make_tmap_bf16(&th, H, M, N1, N1, BM));
    B2_TRY(make_tmap_bf16(&tw2, W2, N2, N1, N1, bn));
    FusedProjParams P;
    P.M = M; P.N1 = N1; P.K1 = K1; P.N2 = N2;
    P.ep1.bias = reinterpret_cast<const __nv_bfloat16*>(b1); P.ep1.residual = nullptr; P.ep1.out = H; P.ep1.ld_out = N1;
    P.ep1.ld_res = 0; P.ep1.out_fp32 = 0;
    P.ep2.bias = reinterpret_cast<const __nv_bfloat16*>(b2); P.ep2.residual = nullptr; P.ep2.out = out; P.ep2.ld_out = ld_out;
    P.ep2.ld_res = 0; P.ep2.out_fp32 = 0;
    P.row_done = row_done;
    P.target = ((N1 + bn - 1) / bn) * kNumEpiWarps;
    B2_CUDA_CHECK(cudaMemsetAsync(row_done, 0, (size_t)num_m * sizeof(int), stream));
    if (bn == 192) return launch_fused_projector<192>(tx, tw1, th, tw2, P, stream);
    return launch_fused_projector<256>(tx, tw1, th, tw2, P, stream);
}

// C[M, N(/2 for swiglu)] = epi(A[M,K](lda) · W[N,K](ldw)^T). See kernels.h for the contract.
int gemm_bf16(const GemmArgs& g, cudaStream_t stream) {
    B2_CHECK_ARG(g.M > 0 && g.N > 0 && g.K > 0, "gemm: empty problem M=%d N=%d K=%d", g.M, g.N, g.K);
    B2_CHECK_ARG(g.K % 8 == 0 && g.N % 8 == 0, "gemm: K and N must be multiples of 8 (K=%d N=%d)", g.K, g.N);
    B2_CHECK_ARG(g.act != ACT_SWIGLU || g.N % 128 == 0, "gemm: swiglu needs N %% 128 == 0 (N=%d)", g.N);
    B2_CHECK_ARG(g.act != ACT_SWIGLU || (!g.out_fp32 && g.bias == nullptr && g.residual == nullptr),
                 "gemm: swiglu epilogue takes no bias/residual and writes bf16");
    B2_CHECK_ARG((reinterpret_cast<uintptr_t>(g.out) & 15) == 0 && (g.ld_out % 8) == 0,
                 "gemm: out must be 16B aligned with ld_out %% 8 == 0");
    B2_CHECK_ARG(g.residual == nullptr ||
                     ((reinterpret_cast<uintptr_t>(g.residual) & 15) == 0 && (g.ld_res % 8) == 0),
                 "gemm: residual must be 16B aligned with ld_res %% 8 == 0");

    if (g.bn_override == 2) return gemm_bf16_2cta(g, stream);  // CTA-pair kernel (cta_group::2, 256x256 pair tiles)

    // Tile-N choice. Cost model fitted to profiles/r1b_gemm_sweep.json: a tile costs ~BN / rel(BN) (rel = tensor-pipe
    // feed efficiency of the shape: BN=256 needs 96 B/clk of smem operand traffic, BN=128 sits on the 128 B/clk limit,
    // BN=64 is far over it) and the launch takes ceil(tiles / #SMs) waves. BN=192 exists for the N=4096 projections of a
    // B=1 prefill (M=704): 96 tiles of 256 leave a third of the SMs idle, 132 tiles of 192 fill one wave.
    // The CTA-pair kernel competes as a fifth shape: a 256x256 pair tile is per-SM the work of a 128x256 tile with half the
    // B-operand smem traffic; measured 4-10 % above BN=256 wherever it fills its waves (profiles/r2b_gemm_sweep.json: 7B
    // prefill qkv 1146 -> 1237, gate/up 1157 -> 1255 TFLOP/s at M=704; 1322 -> 1442 / 1410 -> 1556 at M=5632; ViT fc2 at B=64
    // 1389 -> 1488), and behind BN=192 where 48 pair tiles leave two thirds of the pairs idle (N=4096 projections at M=704).
    const int num_m = (g.M + BM - 1) / BM;
    int bn = 64;
    {
        const int cand[4] = {256, 192, 128, 64};
        const float rel[4] = {1.00f, 0.93f, 0.81f, 0.44f};
        float best = 0.f;
        for (int i = 0; i < 4; ++i) {
            if (g.act == ACT_SWIGLU && cand[i] % 128 != 0) continue;
            const long long tiles = (long long)num_m * ((g.N + cand[i] - 1) / cand[i]);
            const long long waves = (tiles + num_sms() - 1) / num_sms();
            const float cost = (float)waves * ((float)cand[i] / rel[i] + 24.f /*per-tile fixed cost*/);
            if (best == 0.f || cost < best) { best = cost; bn = cand[i]; }
        }
        if (g.bn_override == 0 && g.M >= 512 && g.N >= 256) {
            const long long pairs = (long long)((g.M + 255) / 256) * ((g.N + 255) / 256);
            const long long pair_slots = num_sms() / 2;
            const long long waves = (pairs + pair_slots - 1) / pair_slots;
            const float cost = (float)waves * (256.f / 1.08f + 24.f);
            if (cost < best) return gemm_bf16_2cta(g, stream);
        }
    }
    if (g.bn_override == 64 || g.bn_override == 128 || g.bn_override == 192 || g.bn_override == 256) bn = g.bn_override;
    if (g.act == ACT_SWIGLU && bn % 128 != 0) bn = 128;

    CUtensorMap ta, tb;
    B2_TRY(make_tmap_bf16(&ta, g.A, g.M, g.K, g.lda, BM));
    B2_TRY(make_tmap_bf16(&tb, g.W, g.N, g.K, g.ldw, bn));
    EpiParams ep;
    ep.bias = reinterpret_cast<const __nv_bfloat16*>(g.bias);
    ep.residual = reinterpret_cast<const __nv_bfloat16*>(g.residual);
    ep.out = g.out;
    ep.ld_out = g.ld_out;
    ep.ld_res = g.ld_res;
    ep.out_fp32 = g.out_fp32;

    if (g.act == ACT_SWIGLU) {
        if (bn == 256) return launch_gemm<256, ACT_SWIGLU>(ta, tb, g.M, g.N, g.K, ep, stream);
        return launch_gemm<128, ACT_SWIGLU>(ta, tb, g.M, g.N, g.K, ep, stream);
    }
    if (bn == 64) return dispatch_act<64>(g.act, ta, tb, g.M, g.N, g.K, ep, stream);
    if (bn == 192) return dispatch_act<192>(g.act, ta, tb, g.M, g.N, g.K, ep, stream);
    if (bn == 256) return dispatch_act<256>(g.act, ta, tb, g.M, g.N, g.K, ep, stream);
    return dispatch_act<128>(g.act, ta, tb, g.M, g.N, g.K, ep, stream);
}

}  // namespace b2
