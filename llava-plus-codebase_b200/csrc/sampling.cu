// Token selection + publication for the decode loop (one CTA per sample):
//
//   * greedy: argmax over the fp32 last-position logits (first occurrence wins, like torch.argmax) — the HF
//     GenerationMixin greedy_search step the reference's callers use with do_sample=False
//     (llava/eval/model_vqa_loader.py:98-106, llava/serve/model_worker.py:161 when temperature <= 0.001);
//   * sampling: HF's warper chain for do_sample=True as invoked by llava/serve/model_worker.py:155-185
//     (temperature, top_p; top_k from the GenerationConfig default): logits / T -> keep the k largest (ties kept) ->
//     softmax over the survivors -> nucleus: keep token i iff the probability mass STRICTLY above it is < top_p
//     (== TopPLogitsWarper's "remove where ascending cumsum <= 1 - top_p", at least one token kept) -> draw from the
//     renormalised survivors with a counter-based Philox4x32-10 stream keyed by (seed; token index, row).
//     Everything after exp() is integer arithmetic (probabilities as 2^-40 fixed point, radix select over float bit
//     patterns, integer prefix sums), so a draw is reproducible bit for bit from (logits, seed, index) whatever the
//     thread schedule (the test suite carries a numpy restatement).
//   * publication: the chosen token goes to kv->tok (next step's input, stays on the device), to the step's slot
//     in out_tokens, and — tagged with the generation's epoch — into a ring in MAPPED PINNED HOST memory, so the host
//     loop of generate() (streamer / stopping criteria, llava/serve/model_worker.py:166-188) reads token t while the
//     device is already running step t+k: no D2H copy, no stream sync per token.
#include <limits.h>
#include <math.h>

#include "common.cuh"
#include "kernels.h"

namespace b2 {
namespace {

constexpr int SP_THREADS = 1024;
constexpr float SP_FIXED_ONE = 1099511627776.0f;  // 2^40: probability mass in fixed point

__device__ __forceinline__ uint32_t order_key(float x) {  // monotone float -> uint (x < y  <=>  key(x) < key(y))
    const uint32_t u = __float_as_uint(x);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ unsigned long long mass_of(float e) {
    return e > 0.f ? __float2ull_rz(e * SP_FIXED_ONE) : 0ull;  // NaN / -inf survivors carry no mass
}

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ unsigned long long philox_u64(unsigned long long seed, uint32_t index, uint32_t row) {
    uint32_t c[4] = {index, row, 0u, 0u};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return ((unsigned long long)c[0] << 32) | c[1];
}

template <typename T>
__device__ __forceinline__ T block_sum(T v, T* s_w, int tid) {  // fixed order: warp shuffle tree, then warp sums in order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((tid & 31) == 0) s_w[tid >> 5] = v;
    __syncthreads();
    T t = 0;
    for (int i = 0; i < SP_THREADS / 32; ++i) t += s_w[i];
    return t;
}

// argmax with torch.argmax tie-breaking (value desc, index asc), NaN skipped; result valid in every thread
__device__ __forceinline__ int block_argmax(const float* row, int V, int tid, float* s_v, int* s_i, float* max_out) {
    float best = -INFINITY;
    int bi = INT_MAX;
    for (int i = tid; i < V; i += SP_THREADS) {
        const float v = row[i];
        if (v == v && (bi == INT_MAX || v > best)) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    __syncthreads();
    if ((tid & 31) == 0) { s_v[tid >> 5] = best; s_i[tid >> 5] = bi; }
    __syncthreads();
    best = s_v[0]; bi = s_i[0];
    for (int w = 1; w < SP_THREADS / 32; ++w) {
        const float ov = s_v[w];
        const int oi = s_i[w];
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (max_out) *max_out = best;
    return bi == INT_MAX ? 0 : bi;
}

// Radix descent, 8 bits per level from the top. Level state lives in shared memory (s_prefix / s_above).
//   COUNT mode (top-k): weight of an element = 1; selects the key of the k-th largest element.
//   MASS  mode (top-p): weight = mass_of(e); selects the smallest key whose strictly-above mass is < limit.
template <bool MASS>
__device__ __forceinline__ uint32_t radix_select(const float* s_x, int V, unsigned long long limit, int tid,
                                                 unsigned long long* s_hist, uint32_t* s_prefix,
                                                 unsigned long long* s_above) {
    if (tid == 0) { *s_prefix = 0u; *s_above = 0ull; }
    for (int level = 0; level < 4; ++level) {
        const int shift = 24 - 8 * level;
        if (tid < 256) s_hist[tid] = 0ull;
        __syncthreads();
        const uint32_t prefix = *s_prefix;
        for (int i = tid; i < V; i += SP_THREADS) {
            const float x = s_x[i];
            const uint32_t key = MASS ? __float_as_uint(x > 0.f ? x : 0.f) : order_key(x);
            if (level == 0 || (key >> (shift + 8)) == prefix) {
                const unsigned long long w = MASS ? mass_of(x) : 1ull;
                if (w) atomicAdd(&s_hist[(key >> shift) & 255u], w);  // integer adds: order-independent
            }
        }
        __syncthreads();
        if (tid == 0) {
            unsigned long long acc = *s_above;
            int pick = 255;
            if (MASS) {
                // smallest digit d whose strictly-above mass acc_d is still < limit (acc_255 = above < limit by construction)
                for (int d = 255; d >= 0; --d) {
                    if (acc >= limit) break;
                    pick = d;
                    *s_above = acc;
                    acc += s_hist[d];
                }
            } else {
                // digit holding the limit-th largest element: walk down until the running count reaches it
                pick = 0;
                for (int d = 255; d >= 0; --d) {
                    if (acc + s_hist[d] >= limit) { pick = d; *s_above = acc; break; }
                    acc += s_hist[d];
                }
            }
            *s_prefix = (prefix << 8) | (uint32_t)pick;
        }
        __syncthreads();
    }
    return *s_prefix;
}

__global__ void __launch_bounds__(SP_THREADS, 1)
sample_publish_kernel(const float* __restrict__ logits, int V, int B, SampleState* st, RowState* rows, int32_t* tok, int32_t* out_tokens,
                      int32_t* step_counter, int32_t* cur_len, volatile int32_t* ring, int ring_cap, int flags,
                      int step_offset) {
    extern __shared__ __align__(16) uint8_t sp_smem[];
    float* s_x = reinterpret_cast<float*>(sp_smem);
    __shared__ unsigned long long s_hist[256];
    __shared__ unsigned long long s_wsum[SP_THREADS / 32];
    __shared__ unsigned long long s_above, s_target;
    __shared__ uint32_t s_prefix;
    __shared__ float s_v[SP_THREADS / 32];
    __shared__ int s_i[SP_THREADS / 32];
    __shared__ int s_choice;

    const int tid = threadIdx.x, b = blockIdx.x;
    pdl_trigger();
    pdl_wait();  // logits / tok / the selection state are outputs of earlier kernels in the stream
    const bool per_row = st->per_row != 0;
    const bool active = !per_row || rows[b].active != 0;
    const int do_sample = per_row ? rows[b].do_sample : st->do_sample;
    const float temperature = per_row ? rows[b].temperature : st->temperature;
    const float top_p = per_row ? rows[b].top_p : st->top_p;
    const int top_k = per_row ? rows[b].top_k : st->top_k;
    const unsigned long long seed = per_row ? rows[b].seed : st->seed;
    const int pub = st->pub_counter;
    const int draw = per_row ? rows[b].index : pub;
    int choice = 0;

    if (!active) {
        choice = tok[b];  // an idle slot keeps its token; nothing is selected or counted for it
    } else if (!(flags & SP_SELECT)) {
        choice = tok[b];  // already chosen by the producer of `tok` (decode megakernel's fused argmax)
    } else if (!do_sample) {
        choice = block_argmax(logits + (size_t)b * V, V, tid, s_v, s_i, nullptr);
    } else {
        const float* row = logits + (size_t)b * V;
        const float inv_t = 1.0f / temperature;
        for (int i = tid; i < V; i += SP_THREADS) s_x[i] = row[i] * inv_t;
        __syncthreads();
        // ---- top-k: keep everything >= the k-th largest scaled logit (ties kept, like TopKLogitsWarper) ----
        const int k = top_k;
        if (k > 0 && k < V) {
            const uint32_t kth = radix_select<false>(s_x, V, (unsigned long long)k, tid, s_hist, &s_prefix, &s_above);
            for (int i = tid; i < V; i += SP_THREADS)
                if (order_key(s_x[i]) < kth) s_x[i] = -INFINITY;
            __syncthreads();
        }
        // ---- softmax numerators over the survivors: e_i = exp(x_i - max) (max itself always survives) ----
        float mx;
        block_argmax(s_x, V, tid, s_v, s_i, &mx);
        for (int i = tid; i < V; i += SP_THREADS) {
            const float x = s_x[i];
            s_x[i] = (x == x && x > -INFINITY) ? expf(x - mx) : 0.f;
        }
        __syncthreads();
        // ---- top-p over the fixed-point masses ----
        if (top_p < 1.0f) {
            unsigned long long part = 0ull;
            for (int i = tid; i < V; i += SP_THREADS) part += mass_of(s_x[i]);
            const unsigned long long total = block_sum<unsigned long long>(part, s_wsum, tid);
            unsigned long long limit = (unsigned long long)((double)total * (double)(top_p > 0.f ? top_p : 0.f));
            if (limit < 1ull) limit = 1ull;  // min_tokens_to_keep = 1: the largest probability always survives
            const uint32_t thr = radix_select<true>(s_x, V, limit, tid, s_hist, &s_prefix, &s_above);
            for (int i = tid; i < V; i += SP_THREADS) {
                const float x = s_x[i];
                if (__float_as_uint(x > 0.f ? x : 0.f) < thr) s_x[i] = 0.f;
            }
            __syncthreads();
        }
        // ---- inverse CDF in index order: thread t owns the contiguous chunk [t*C, (t+1)*C) ----
        const int C = ((V + SP_THREADS - 1) / SP_THREADS) | 1;  // odd stride: conflict-free shared-memory walks
        const int lo = tid * C, hi = min(lo + C, V);
        unsigned long long mine = 0ull;
        for (int i = lo; i < hi; ++i) mine += mass_of(s_x[i]);
        // exclusive prefix over threads: warp scan + scan of the 32 warp totals
        unsigned long long incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
            if ((tid & 31) >= o) incl += up;
        }
        if ((tid & 31) == 31) s_wsum[tid >> 5] = incl;
        __syncthreads();
        unsigned long long warp_base = 0ull, total = 0ull;
        for (int w = 0; w < SP_THREADS / 32; ++w) {
            if (w < (tid >> 5)) warp_base += s_wsum[w];
            total += s_wsum[w];
        }
        const unsigned long long excl = warp_base + incl - mine;
        if (tid == 0) {
            s_target = total ? __umul64hi(total, philox_u64(seed, (uint32_t)draw, (uint32_t)b)) : 0ull;
            s_choice = 0;
        }
        __syncthreads();
        const unsigned long long target = s_target;
        if (mine != 0ull && target >= excl && target < excl + mine) {  // exactly one thread
            unsigned long long run = excl;
            int pick = lo;
            for (int i = lo; i < hi; ++i) {
                const unsigned long long m = mass_of(s_x[i]);
                if (m != 0ull && target < run + m) { pick = i; break; }
                run += m;
            }
            s_choice = pick;
        }
        __syncthreads();
        choice = s_choice;
    }

    if (tid == 0) {
        if ((flags & SP_SELECT) && active) tok[b] = choice;
        if (per_row && active) rows[b].index = draw + 1;
        if (flags & SP_WRITE_OUT) out_tokens[(size_t)(*step_counter + step_offset) * B + b] = choice;
        if (ring != nullptr && st->tag != 0) {
            // the tag advances every time the ring wraps, so an entry left from ring_cap steps ago is never taken for a new one
            const int tag = 1 + (st->tag - 1 + pub / ring_cap) % 2047;
            ring[(size_t)(pub % ring_cap) * B + b] = (tag << 20) | (choice & 0xFFFFF);
            __threadfence_system();
        }
        __threadfence();
        if (atomicAdd(&st->done, 1u) == (unsigned)B - 1u) {  // last row of this step: advance the counters
            st->done = 0u;
            st->pub_counter = pub + 1;
            if (flags & SP_BUMP) {
                *step_counter += 1;
                for (int i = 0; i < B; ++i)
                    if (!per_row || rows[i].active) cur_len[i] += 1;
            }
        }
    }
}

__global__ void sample_state_set_kernel(SampleState* st, SampleState v) {
    if (threadIdx.x == 0) *st = v;
}
__global__ void row_state_set_kernel(RowState* row, RowState v, int32_t* tok, int token) {
    if (threadIdx.x == 0) {
        *row = v;
        if (tok != nullptr) *tok = token;
    }
}

}  // namespace

size_t sample_smem_bytes(int V) { return (size_t)V * sizeof(float); }

int sample_state_set(SampleState* st_dev, const SampleState& v, cudaStream_t stream) {
    sample_state_set_kernel<<<1, 32, 0, stream>>>(st_dev, v);
    B2_LAUNCH_CHECK();
    return 0;
}

int row_state_set(RowState* row_dev, const RowState& v, int32_t* tok_dev, int token, cudaStream_t stream) {
    row_state_set_kernel<<<1, 32, 0, stream>>>(row_dev, v, tok_dev, token);
    B2_LAUNCH_CHECK();
    return 0;
}

int sample_publish(const float* logits, int V, int B, SampleState* st_dev, RowState* rows_dev, int32_t* tok, int32_t* out_tokens,
                   int32_t* step_counter, int32_t* cur_len, int32_t* ring_dev, int ring_cap, int flags, int step_offset,
                   cudaStream_t stream) {
    B2_CHECK_ARG(B >= 1 && V >= 1 && st_dev != nullptr && tok != nullptr, "sample_publish: bad argument");
    B2_CHECK_ARG(V < (1 << 20), "sample_publish: vocab %d does not fit the 20-bit token field of the host ring", V);
    const size_t smem = sample_smem_bytes(V);
    B2_CHECK_ARG(smem <= 200 * 1024, "sample_publish: vocab %d exceeds the shared-memory staging of the sampling kernel", V);
    static size_t attr = 0;
    if (smem > attr) {
        B2_CUDA_CHECK(cudaFuncSetAttribute(sample_publish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    B2_CUDA_CHECK(launch_pdl(sample_publish_kernel, dim3(B), dim3(SP_THREADS), smem, stream, logits, V, B, st_dev, rows_dev, tok,
                             out_tokens, step_counter, cur_len, ring_dev, ring_cap, flags, step_offset));
    B2_LAUNCH_CHECK();
    return 0;
}

}  // namespace b2
