// Engine + C ABI of libb2llava.so: weight ingestion/repack, workspaces, and the orchestration of the
// hot path (CLIP ViT -> mm_projector -> splice -> LLaMA prefill -> KV-cache decode) over the kernels in this
// directory. Entry points are declared in include/b2llava.h, which cites the reference function each replaces.
#include <cuda_fp16.h>
#include <limits.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/b2llava.h"
#include "common.cuh"
#include "kernels.h"

namespace b2 {

static thread_local char g_err[1024] = {0};
unsigned long long g_launch_count = 0;

bool pdl_enabled() {  // read per launch (one getenv) so that one process can A/B it; graphs keep what they were captured with
    const char* e = getenv("B2_PDL");
    return !(e != nullptr && e[0] == '0');
}

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

namespace {
typedef __nv_bfloat16 bf16;

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int alloc(size_t n) {
        free();
        if (n == 0) return 0;
        cudaError_t e = cudaMalloc(&p, n);
        if (e != cudaSuccess) {
            p = nullptr;
            set_error("cudaMalloc(%zu bytes) failed: %s", n, cudaGetErrorString(e));
            return -2;
        }
        bytes = n;
        return 0;
    }
    void free() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    template <typename T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

struct VitLayer {
    DevBuf ln1_g, ln1_b, wqkv, bqkv, wo, bo, ln2_g, ln2_b, w1, b1, w2, b2;
    unsigned qkv_w = 0, qkv_b = 0;  // bit j set: part j (q/k/v) of the fused buffer has arrived
};
struct LlamaLayer {
    DevBuf ln1, wqkv, wo, ln2, wgu, wd;
    unsigned qkv_parts = 0;
    DevBuf wqkv8, wo8, wgu8, wd8, s_qkv, s_o, s_gu, s_d;  // e4m3 copies + per-row fp32 scales (b2_model_enable_fp8_decode)
    DevBuf tmp_gate, tmp_up;  // staging until both halves arrived
    bool has_gate = false, has_up = false;
};
}  // namespace
}  // namespace b2

using namespace b2;

struct b2_model {
    b2_model_desc d;
    int device = 0;
    std::mutex mu;
    bool finalized = false;
    // derived
    int P = 0, T = 0, kpad = 0, vit_live = 0, vit_hd = 64, hd = 128;
    // weights
    DevBuf patch_w, cls, pos, pre_g, pre_b;
    std::vector<VitLayer> vit;
    DevBuf p0_w, p0_b, p2_w, p2_b;
    DevBuf embed, final_norm, lm_head;
    std::vector<LlamaLayer> ll;
    // errors detected by kernels (bad token ids / image rows): int[8] in mapped pinned host memory, slot = log2(B2_ERR_*)
    int* err_host = nullptr;
    int* err_dev = nullptr;
    // the workspaces below are shared by every call on this model: a call on stream B must not start before the previous
    // call's kernels on stream A are done with them (calls on one stream are ordered anyway)
    cudaEvent_t ws_event = nullptr;
    cudaStream_t ws_stream = nullptr;
    bool ws_valid = false;
    // ViT workspace (per chunk of max_images)
    DevBuf v_col, v_patch, v_hidden, v_xn, v_qkv, v_attn, v_mlp, v_feats, p_mid, p_done;
    // LLaMA workspace
    DevBuf x, xn, qkv, attn, act, last_idx, xlast, logits, splice_idx;
    // encode_images replays a CUDA graph per chunk size (~190 launches per image, 5-20 us each at B = 1: the host cannot
    // keep the GPU fed launch by launch). Inputs/outputs are staged through fixed buffers so the captured pointers stay valid.
    DevBuf enc_pixels, enc_out;
    std::vector<cudaGraphExec_t> enc_graph;  // index = images in the chunk (0 = unused)
    std::vector<char> enc_warm;              // an eager run has set the function attributes for this chunk size
    cudaStream_t enc_stream = nullptr;       // capture is illegal on the legacy default stream: such callers run here
    cudaEvent_t enc_fork = nullptr, enc_join = nullptr;
    int enc_launches = 0;                    // kernels in one captured chunk (b2_launch_count bookkeeping)
    // fp8 decode (BASELINE configs[4]): e4m3 lm_head + scales, quantised activation row buffer + per-token scales
    bool fp8_decode = false;
    DevBuf lm_head8, s_head, xq8, xscale;
};

struct b2_kv {
    b2_model* m = nullptr;
    int max_batch = 0, max_seq = 0;
    DevBuf k, v;  // [L][B][H][Smax][D]
    DevBuf len_dev, tok, step_counter, out_tokens, attn_partial, attn_counters;
    DevBuf sk_partial, sk_counters;  // stream-K workspace of the skinny decode GEMM (batch 9..128)
    std::vector<int32_t> len_host;
    int out_capacity = 0;  // steps
    // cached decode-step graph
    cudaGraphExec_t graph = nullptr;
    int graph_B = 0;
    // stream capture is illegal on the legacy default stream (torch's default current stream): decode steps
    // run on this library-owned stream, ordered against the caller's stream with events
    unsigned int mega_bar_base = 0;  // value of the grid-barrier counter before the next megakernel launch
    DevBuf mega_layers, mega_sync;  // MegaLayer[L] table and {bar_count, bar_gen, done_count}
    DevBuf rope_tab;                // uint32 [max_seq][hd/2]: bf16 (cos, sin) per position for the QKV GEMM's fused RoPE epilogue
    cudaStream_t own_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int warm_B = 0;  // an eager step has run for this B (function attributes set, driver entry points resolved)
    // token selection state (sampling.cu) + the host-visible token ring of the streaming decode API
    DevBuf sstate;                 // SampleState on the device
    SampleState samp_host = {};    // what was last written there (pub_counter / done excluded: the device advances them)
    bool samp_valid = false;
    int32_t* ring_host = nullptr;  // cudaHostAlloc(mapped) [ring_cap][max_batch]; entry = tag << 20 | token
    int32_t* ring_dev = nullptr;
    int ring_cap = 0;
    int epoch = 0;                 // generations started on this cache (tag = 1 + epoch % 2047)
    int stream_B = 0, stream_tag = 0, stream_scheduled = 0;  // streaming generation in progress: tokens scheduled so far
    DevBuf rows_dev;               // RowState[max_batch] (continuous batching)
    std::vector<RowState> rows_host;
    size_t layer_stride() const { return (size_t)max_batch * m->d.heads * max_seq * m->hd; }
};

namespace {

bool starts_with(const std::string& s, const char* p) { return s.rfind(p, 0) == 0; }

// copy `n` elements of `dtype` from host-or-device `src` into bf16 device memory `dst` (contiguous)
int ingest(const void* src, int dtype, void* dst, int64_t n, cudaStream_t st) {
    if (dtype == DT_BF16) {
        B2_CUDA_CHECK(cudaMemcpyAsync(dst, src, (size_t)n * 2, cudaMemcpyDefault, st));
        B2_CUDA_CHECK(cudaStreamSynchronize(st));
        return 0;
    }
    const size_t esz = dtype == DT_F32 ? 4 : 2;
    cudaPointerAttributes attr;
    bool on_device = false;
    if (cudaPointerGetAttributes(&attr, src) == cudaSuccess)
        on_device = attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
    else
        cudaGetLastError();
    DevBuf tmp;
    const void* dsrc = src;
    if (!on_device) {
        B2_TRY(tmp.alloc((size_t)n * esz));
        B2_CUDA_CHECK(cudaMemcpy(tmp.p, src, (size_t)n * esz, cudaMemcpyDefault));
        dsrc = tmp.p;
    }
    int r = convert_to_bf16(dsrc, dtype, dst, n, st);
    cudaError_t e = cudaStreamSynchronize(st);
    tmp.free();
    if (r != 0) return r;
    B2_CUDA_CHECK(e);
    return 0;
}

int64_t numel(const int64_t* shape, int ndim) {
    int64_t n = 1;
    for (int i = 0; i < ndim; ++i) n *= shape[i];
    return n;
}

int expect_shape(const char* key, const int64_t* shape, int ndim, int64_t a, int64_t b = -1) {
    const int64_t n = numel(shape, ndim);
    const int64_t want = b < 0 ? a : a * b;
    bool ok = n == want;
    if (ok && b >= 0 && ndim >= 2) ok = shape[0] == a;
    if (!ok) {
        set_error("set_weight(%s): unexpected shape (numel %lld, expected %lld x %lld)", key, (long long)n,
                  (long long)a, (long long)(b < 0 ? 1 : b));
        return -1;
    }
    return 0;
}

// alloc-if-needed + ingest into dst buffer at element offset `off`
int put(DevBuf& buf, size_t total_elems, size_t off, const void* src, int dtype, int64_t n) {
    if (buf.p == nullptr) B2_TRY(buf.alloc(total_elems * 2));
    return ingest(src, dtype, buf.as<bf16>() + off, n, 0);
}

int gemm(const void* A, int lda, const void* W, int ldw, const void* bias, const void* res, int ld_res, void* out,
         int ld_out, int out_fp32, int M, int N, int K, int act, cudaStream_t st) {
    GemmArgs g;
    g.A = A; g.lda = lda; g.W = W; g.ldw = ldw; g.bias = bias; g.residual = res; g.ld_res = ld_res;
    g.out = out; g.ld_out = ld_out; g.out_fp32 = out_fp32; g.M = M; g.N = N; g.K = K; g.act = act;
    return gemm_bf16(g, st);
}

int gemv(const void* x, int64_t ldx, const void* W, int ldw, const void* gamma, float eps, const void* res,
         int ld_res, void* out, int ld_out, int out_fp32, int B, int N, int K, int act, cudaStream_t st) {
    GemvArgs g;
    g.x = x; g.ldx = ldx; g.W = W; g.ldw = ldw; g.norm_gamma = gamma; g.eps = eps; g.residual = res;
    g.ld_res = ld_res; g.out = out; g.ld_out = ld_out; g.out_fp32 = out_fp32; g.B = B; g.N = N; g.K = K;
    g.act = act;
    return gemv_bf16(g, st);
}

// decode Linear at batch 9..128: swap-AB stream-K tcgen05 GEMM over the kv-owned workspace
int skinny(b2_kv* kv, const void* x, int ldx, const void* W, int ldw, const void* res, int ld_res, void* out, int ld_out,
           int out_fp32, int B, int N, int K, int act, cudaStream_t st) {
    SkinnyArgs g;
    g.x = x; g.ldx = ldx; g.W = W; g.ldw = ldw; g.residual = res; g.ld_res = ld_res;
    g.out = out; g.ld_out = ld_out; g.out_fp32 = out_fp32; g.B = B; g.N = N; g.K = K; g.act = act;
    g.partial = kv->sk_partial.as<float>(); g.partial_bytes = kv->sk_partial.bytes;
    g.counters = kv->sk_counters.as<int>();
    return gemm_skinny_bf16(g, st);
}
// Batch 7..128. Measured ladder (7B, ms/step, profiles/r1e_config_sweep_*.jsonl): the stream-K GEMM path costs 4.13-4.19 for
// every B <= 8 (7 kernels/layer, ~8 us of fixed cost per GEMM launch: profiles/r1e_decode_launch_shares_b4.txt), the GEMV
// path (5 kernels/layer, RMSNorm fused) 3.63 at B=4 and 4.51 at B=8 -> crossover between 6 and 7. B2_DECODE_SKINNY=0
// restores the round-1 paths for A/B runs (GEMV kernels for B <= 8, tile GEMM with the batch padded to M=128 above),
// B2_DECODE_SKINNY=3 lowers the threshold to 3.
// fp8 variant: xq8 / xscale hold the quantised activations of this GEMM
int skinny8(b2_model* m, b2_kv* kv, const void* W8, const float* w_scale, const void* res, int ld_res, void* out, int ld_out,
            int out_fp32, int B, int N, int K, int act, cudaStream_t st) {
    SkinnyArgs g;
    g.x = m->xq8.p; g.ldx = K; g.W = W8; g.ldw = K; g.residual = res; g.ld_res = ld_res;
    g.out = out; g.ld_out = ld_out; g.out_fp32 = out_fp32; g.B = B; g.N = N; g.K = K; g.act = act;
    g.partial = kv->sk_partial.as<float>(); g.partial_bytes = kv->sk_partial.bytes;
    g.counters = kv->sk_counters.as<int>();
    g.w_scale = w_scale; g.x_scale = m->xscale.as<float>();
    return gemm_skinny_fp8(g, st);
}
bool use_skinny(const b2_kv* kv, int B) {
    const char* e0 = getenv("B2_DECODE_SKINNY");
    const int lo = (e0 != nullptr && e0[0] == '3') ? 3 : 7;
    if (B < lo || B > 128 || kv->sk_partial.p == nullptr) return false;
    const char* e = getenv("B2_DECODE_SKINNY");
    return !(e != nullptr && e[0] == '0');
}

int decode_nsplit(int B, int H, int max_seq) {
    // Split-KV factor of the multi-kernel decode step. The kernel is register-limited to `occ` resident CTAs per SM (6), so
    // one wave is occ*SMs = 888 CTAs. Measured over the ladder (profiles/r2j_*, r2k_decode_ab_nsplit_stages.jsonl, 7B, ctx ~800):
    // B=8 (256 pairs): n=3 4.19 ms vs n=5 4.32 / n=10 4.38; B=16: n=3 4.80 vs n=1 5.05 / n=5 4.90; B=32: n=3 6.50 vs n=2 6.55 /
    // n=6 6.68; B=64: n=1..3 equal. Few pairs: fill one wave; otherwise 3 splits (short enough ranges for the tail wave to
    // overlap, few enough partials for the merge to stay cheap).
    const char* e = getenv("B2_DECODE_NSPLIT");
    if (e != nullptr && atoi(e) >= 1) return atoi(e) > 32 ? 32 : atoi(e);
    const int cap = decode_attn_ctas_per_sm() * num_sms();
    int n = cap / (B * H);
    if (n < 3) n = 3;
    if (n > 16) n = 16;
    const int n_hi = max_seq / 64;  // keep >= 64 keys per split at the cache's capacity
    if (n > n_hi) n = n_hi < 1 ? 1 : n_hi;
    return n;
}

// ------------------------------------------------------------------------------------------------------
// weight routing
// ------------------------------------------------------------------------------------------------------
int set_vision_weight(b2_model* m, const std::string& k, const void* ptr, const int64_t* shape, int ndim, int dt) {
    const int D = m->d.vit_hidden, I = m->d.vit_inter;
    const char* key = k.c_str();
    if (k == "embeddings.class_embedding") {
        B2_TRY(expect_shape(key, shape, ndim, D));
        return put(m->cls, D, 0, ptr, dt, D);
    }
    if (k == "embeddings.patch_embedding.weight") {
        const int kk = 3 * m->d.patch_size * m->d.patch_size;
        B2_TRY(expect_shape(key, shape, ndim, D, kk));
        // [D, 3, ps, ps] -> rows of kpad (zero padded) so the TMA row pitch is a multiple of 16 B
        DevBuf tmp;
        B2_TRY(tmp.alloc((size_t)D * kk * 2));
        B2_TRY(ingest(ptr, dt, tmp.p, (int64_t)D * kk, 0));
        if (m->patch_w.p == nullptr) B2_TRY(m->patch_w.alloc((size_t)D * m->kpad * 2));
        B2_CUDA_CHECK(cudaMemset(m->patch_w.p, 0, (size_t)D * m->kpad * 2));
        B2_CUDA_CHECK(cudaMemcpy2D(m->patch_w.p, (size_t)m->kpad * 2, tmp.p, (size_t)kk * 2, (size_t)kk * 2, D,
                                   cudaMemcpyDeviceToDevice));
        return 0;
    }
    if (k == "embeddings.position_embedding.weight") {
        B2_TRY(expect_shape(key, shape, ndim, m->T, D));
        return put(m->pos, (size_t)m->T * D, 0, ptr, dt, (int64_t)m->T * D);
    }
    if (k == "embeddings.position_ids") return 0;  // buffer in old checkpoints
    if (k == "pre_layrnorm.weight") { B2_TRY(expect_shape(key, shape, ndim, D)); return put(m->pre_g, D, 0, ptr, dt, D); }
    if (k == "pre_layrnorm.bias") { B2_TRY(expect_shape(key, shape, ndim, D)); return put(m->pre_b, D, 0, ptr, dt, D); }
    if (starts_with(k, "post_layernorm.")) return 0;  // dead for select_layer=-2
    if (starts_with(k, "encoder.layers.")) {
        int li = -1, consumed = 0;
        if (sscanf(key, "encoder.layers.%d.%n", &li, &consumed) != 1 || li < 0 || li >= m->d.vit_layers) {
            set_error("set_weight: bad vision layer index in '%s'", key);
            return -1;
        }
        if (li >= m->vit_live) return 0;  // layers above the selected hidden state are never computed
        VitLayer& L = m->vit[li];
        const std::string s = k.substr(consumed);
        if (s == "layer_norm1.weight") { B2_TRY(expect_shape(key, shape, ndim, D)); return put(L.ln1_g, D, 0, ptr, dt, D); }
        if (s == "layer_norm1.bias") { B2_TRY(expect_shape(key, shape, ndim, D)); return put(L.ln1_b, D, 0, ptr, dt, D); }
        if (s == "layer_norm2.weight") { B2_TRY(expect_shape(key, shape, ndim, D)); return put(L.ln2_g, D, 0, ptr, dt, D); }
        if (s == "layer_norm2.bias") { B2_TRY(expect_shape(key, shape, ndim, D)); return put(L.ln2_b, D, 0, ptr, dt, D); }
        const char* names[3] = {"self_attn.q_proj.", "self_attn.k_proj.", "self_attn.v_proj."};
        for (int j = 0; j < 3; ++j) {
            if (starts_with(s, names[j])) {
                if (s == std::string(names[j]) + "weight") {
                    B2_TRY(expect_shape(key, shape, ndim, D, D));
                    B2_TRY(put(L.wqkv, (size_t)3 * D * D, (size_t)j * D * D, ptr, dt, (int64_t)D * D));
                    L.qkv_w |= 1u << j;
                    return 0;
                }
                if (s != std::string(names[j]) + "bias") break;
                B2_TRY(expect_shape(key, shape, ndim, D));
                B2_TRY(put(L.bqkv, (size_t)3 * D, (size_t)j * D, ptr, dt, D));
                L.qkv_b |= 1u << j;
                return 0;
            }
        }
        if (s == "self_attn.out_proj.weight") { B2_TRY(expect_shape(key, shape, ndim, D, D)); return put(L.wo, (size_t)D * D, 0, ptr, dt, (int64_t)D * D); }
        if (s == "self_attn.out_proj.bias") { B2_TRY(expect_shape(key, shape, ndim, D)); return put(L.bo, D, 0, ptr, dt, D); }
        if (s == "mlp.fc1.weight") { B2_TRY(expect_shape(key, shape, ndim, I, D)); return put(L.w1, (size_t)I * D, 0, ptr, dt, (int64_t)I * D); }
        if (s == "mlp.fc1.bias") { B2_TRY(expect_shape(key, shape, ndim, I)); return put(L.b1, I, 0, ptr, dt, I); }
        if (s == "mlp.fc2.weight") { B2_TRY(expect_shape(key, shape, ndim, D, I)); return put(L.w2, (size_t)D * I, 0, ptr, dt, (int64_t)D * I); }
        if (s == "mlp.fc2.bias") { B2_TRY(expect_shape(key, shape, ndim, D)); return put(L.b2, D, 0, ptr, dt, D); }
    }
    set_error("set_weight: unknown vision key '%s'", key);
    return -1;
}

int set_llama_layer_weight(b2_model* m, int li, const std::string& s, const char* key, const void* ptr,
                           const int64_t* shape, int ndim, int dt) {
    const int h = m->d.hidden, I = m->d.inter;
    LlamaLayer& L = m->ll[li];
    if (s == "input_layernorm.weight") { B2_TRY(expect_shape(key, shape, ndim, h)); return put(L.ln1, h, 0, ptr, dt, h); }
    if (s == "post_attention_layernorm.weight") { B2_TRY(expect_shape(key, shape, ndim, h)); return put(L.ln2, h, 0, ptr, dt, h); }
    const char* names[3] = {"self_attn.q_proj.weight", "self_attn.k_proj.weight", "self_attn.v_proj.weight"};
    for (int j = 0; j < 3; ++j) {
        if (s == names[j]) {
            B2_TRY(expect_shape(key, shape, ndim, h, h));
            B2_TRY(put(L.wqkv, (size_t)3 * h * h, (size_t)j * h * h, ptr, dt, (int64_t)h * h));
            L.qkv_parts |= 1u << j;
            return 0;
        }
    }
    if (s == "self_attn.o_proj.weight") { B2_TRY(expect_shape(key, shape, ndim, h, h)); return put(L.wo, (size_t)h * h, 0, ptr, dt, (int64_t)h * h); }
    if (s == "mlp.down_proj.weight") { B2_TRY(expect_shape(key, shape, ndim, h, I)); return put(L.wd, (size_t)h * I, 0, ptr, dt, (int64_t)h * I); }
    if (s == "mlp.gate_proj.weight" || s == "mlp.up_proj.weight") {
        B2_TRY(expect_shape(key, shape, ndim, I, h));
        const bool is_gate = s == "mlp.gate_proj.weight";
        DevBuf& tmp = is_gate ? L.tmp_gate : L.tmp_up;
        B2_TRY(put(tmp, (size_t)I * h, 0, ptr, dt, (int64_t)I * h));
        (is_gate ? L.has_gate : L.has_up) = true;
        if (L.has_gate && L.has_up) {
            if (L.wgu.p == nullptr) B2_TRY(L.wgu.alloc((size_t)2 * I * h * 2));
            B2_TRY(interleave_gate_up(L.tmp_gate.p, L.tmp_up.p, L.wgu.p, I, h, 0));
            B2_CUDA_CHECK(cudaStreamSynchronize(0));
            L.tmp_gate.free();
            L.tmp_up.free();
            L.has_gate = L.has_up = false;  // allow a later reload
        }
        return 0;
    }
    if (s == "self_attn.rotary_emb.inv_freq") return 0;  // buffer in 4.31-era checkpoints
    set_error("set_weight: unknown decoder key '%s'", key);
    return -1;
}

// ------------------------------------------------------------------------------------------------------
// forward passes (caller holds the model lock)
// ------------------------------------------------------------------------------------------------------
int vit_forward_chunk(b2_model* m, const void* pixels, int B, void* out_feats, cudaStream_t st) {
    const b2_model_desc& d = m->d;
    const int D = d.vit_hidden, I = d.vit_inter, H = d.vit_heads, P = m->P, T = m->T;
    const int rows = B * T;
    B2_TRY(vit_im2col(pixels, m->v_col.p, B, d.image_size, d.patch_size, m->kpad, st));
    B2_TRY(gemm(m->v_col.p, m->kpad, m->patch_w.p, m->kpad, nullptr, nullptr, 0, m->v_patch.p, D, 0, B * P, D,
                m->kpad, ACT_NONE, st));
    B2_TRY(vit_embed_ln(m->v_patch.p, m->cls.p, m->pos.p, m->pre_g.p, m->pre_b.p, m->v_hidden.p, B, P, D,
                        d.vit_ln_eps, st));
    for (int l = 0; l < m->vit_live; ++l) {
        VitLayer& L = m->vit[l];
        B2_TRY(layernorm_bf16(m->v_hidden.p, L.ln1_g.p, L.ln1_b.p, m->v_xn.p, rows, D, d.vit_ln_eps, st));
        B2_TRY(gemm(m->v_xn.p, D, L.wqkv.p, D, L.bqkv.p, nullptr, 0, m->v_qkv.p, 3 * D, 0, rows, 3 * D, D,
                    ACT_NONE, st));
        FlashArgs fa;
        bf16* qkv = m->v_qkv.as<bf16>();
        fa.q = qkv;         fa.q_bs = (int64_t)T * 3 * D; fa.q_ts = 3 * D; fa.q_hs = m->vit_hd;
        fa.k = qkv + D;     fa.k_bs = fa.q_bs; fa.k_ts = 3 * D; fa.k_hs = m->vit_hd;
        fa.v = qkv + 2 * D; fa.v_bs = fa.q_bs; fa.v_ts = 3 * D; fa.v_hs = m->vit_hd;
        fa.o = m->v_attn.p; fa.o_bs = (int64_t)T * D; fa.o_ts = D; fa.o_hs = m->vit_hd;
        fa.B = B; fa.H = H; fa.S = T; fa.D = m->vit_hd; fa.causal = 0;
        fa.scale = 1.0f / sqrtf((float)m->vit_hd);
        B2_TRY(flash_attn_bf16(fa, st));
        B2_TRY(gemm(m->v_attn.p, D, L.wo.p, D, L.bo.p, m->v_hidden.p, D, m->v_hidden.p, D, 0, rows, D, D, ACT_NONE,
                    st));
        B2_TRY(layernorm_bf16(m->v_hidden.p, L.ln2_g.p, L.ln2_b.p, m->v_xn.p, rows, D, d.vit_ln_eps, st));
        B2_TRY(gemm(m->v_xn.p, D, L.w1.p, D, L.b1.p, nullptr, 0, m->v_mlp.p, I, 0, rows, I, D, ACT_QUICK_GELU, st));
        B2_TRY(gemm(m->v_mlp.p, I, L.w2.p, I, L.b2.p, m->v_hidden.p, D, m->v_hidden.p, D, 0, rows, D, I, ACT_NONE,
                    st));
    }
    B2_TRY(vit_drop_cls(m->v_hidden.p, out_feats, B, P, D, st));
    return 0;
}

int project_rows(b2_model* m, const void* feats, int rows, void* out, cudaStream_t st) {
    const int D = m->d.vit_hidden, h = m->d.hidden;
    const int max_rows = m->d.max_images * m->P;
    for (int r0 = 0; r0 < rows; r0 += max_rows) {
        const int n = rows - r0 < max_rows ? rows - r0 : max_rows;
        const bf16* a = reinterpret_cast<const bf16*>(feats) + (size_t)r0 * D;
        bf16* o = reinterpret_cast<bf16*>(out) + (size_t)r0 * h;
        const char* pf = getenv("B2_PROJECTOR_FUSED");  // =0 restores the two-launch form (A/B runs, read per call)
        if (!(pf != nullptr && pf[0] == '0')) {
            // north_star: "mm_projector as one fused GEMM->GELU->GEMM kernel" (phase-2 tiles gated on per-row-block counters)
            B2_TRY(projector_fused_bf16(a, D, m->p0_w.p, m->p0_b.p, m->p2_w.p, m->p2_b.p, m->p_mid.p, o, h, n, D, h, h,
                                        m->p_done.as<int>(), st));
        } else {
            B2_TRY(gemm(a, D, m->p0_w.p, D, m->p0_b.p, nullptr, 0, m->p_mid.p, h, 0, n, h, D, ACT_GELU_ERF, st));
            B2_TRY(gemm(m->p_mid.p, h, m->p2_w.p, h, m->p2_b.p, nullptr, 0, o, h, 0, n, h, h, ACT_NONE, st));
        }
    }
    return 0;
}

// vision tower + projector for one chunk of n <= max_images images: pixels -> out [n*P, hidden]. First call per chunk size runs
// eagerly (function attributes, driver entry points), the second captures, later ones replay.
int encode_chunk(b2_model* m, const void* pixels, int n, void* out, cudaStream_t st) {
    const b2_model_desc& d = m->d;
    const char* eg = getenv("B2_ENCODE_GRAPH");  // =0: launch by launch (A/B runs)
    if (eg != nullptr && eg[0] == '0') {
        B2_TRY(vit_forward_chunk(m, pixels, n, m->v_feats.p, st));
        return project_rows(m, m->v_feats.p, n * m->P, out, st);
    }
    const size_t in_bytes = (size_t)n * 3 * d.image_size * d.image_size * 2, out_bytes = (size_t)n * m->P * d.hidden * 2;
    cudaStream_t run = st;
    if (st == nullptr || st == cudaStreamLegacy) {
        if (m->enc_stream == nullptr) {
            B2_CUDA_CHECK(cudaStreamCreateWithFlags(&m->enc_stream, cudaStreamNonBlocking));
            B2_CUDA_CHECK(cudaEventCreateWithFlags(&m->enc_fork, cudaEventDisableTiming));
            B2_CUDA_CHECK(cudaEventCreateWithFlags(&m->enc_join, cudaEventDisableTiming));
        }
        B2_CUDA_CHECK(cudaEventRecord(m->enc_fork, st));
        B2_CUDA_CHECK(cudaStreamWaitEvent(m->enc_stream, m->enc_fork, 0));
        run = m->enc_stream;
    }
    B2_CUDA_CHECK(cudaMemcpyAsync(m->enc_pixels.p, pixels, in_bytes, cudaMemcpyDeviceToDevice, run));
    if (!m->enc_warm[n]) {
        B2_TRY(vit_forward_chunk(m, m->enc_pixels.p, n, m->v_feats.p, run));
        B2_TRY(project_rows(m, m->v_feats.p, n * m->P, m->enc_out.p, run));
        m->enc_warm[n] = 1;
    } else {
        if (m->enc_graph[n] == nullptr) {
            cudaGraph_t graph = nullptr;
            B2_CUDA_CHECK(cudaStreamBeginCapture(run, cudaStreamCaptureModeThreadLocal));
            const unsigned long long launches_before = g_launch_count;
            int r = vit_forward_chunk(m, m->enc_pixels.p, n, m->v_feats.p, run);
            if (r == 0) r = project_rows(m, m->v_feats.p, n * m->P, m->enc_out.p, run);
            m->enc_launches = (int)(g_launch_count - launches_before);
            g_launch_count = launches_before;  // capture records launches, it does not run them
            cudaError_t e = cudaStreamEndCapture(run, &graph);
            if (r != 0) { if (graph) cudaGraphDestroy(graph); return r; }
            B2_CUDA_CHECK(e);
            e = cudaGraphInstantiate(&m->enc_graph[n], graph, 0);
            cudaGraphDestroy(graph);
            B2_CUDA_CHECK(e);
        }
        B2_CUDA_CHECK(cudaGraphLaunch(m->enc_graph[n], run));
        g_launch_count += (unsigned long long)m->enc_launches;
    }
    B2_CUDA_CHECK(cudaMemcpyAsync(out, m->enc_out.p, out_bytes, cudaMemcpyDeviceToDevice, run));
    if (run != st) {
        B2_CUDA_CHECK(cudaEventRecord(m->enc_join, run));
        B2_CUDA_CHECK(cudaStreamWaitEvent(st, m->enc_join, 0));
    }
    return 0;
}

// one decode step on kv-owned buffers: tok -> logits (m->logits) -> argmax -> tok, out_tokens[step], len += 1
int decode_step_launch(b2_model* m, b2_kv* kv, int B, cudaStream_t st) {
    const b2_model_desc& d = m->d;
    const int h = d.hidden, I = d.inter, H = d.heads, V = d.vocab;
    const int nsplit = decode_nsplit(B, H, kv->max_seq);
    B2_TRY(embed_tokens(kv->tok.as<int32_t>(), m->embed.p, m->x.p, B, h, V, m->err_dev, st));
    // batch <= 8: tensor-core GEMV kernels (falls back to the skinny-M tcgen05 GEMM when the activations do not fit smem)
    // batch 7..128: swap-AB stream-K GEMM (weights streamed once, all SMs busy); otherwise GEMV kernels (B <= 8) or
    // the tile GEMM
    const bool sk = use_skinny(kv, B);
    const bool small = !sk && B <= 8 && gemv_fits(B, h, I, ACT_NONE) && gemv_fits(B, 2 * I, h, ACT_SWIGLU) &&
                       gemv_fits(B, 3 * h, h, ACT_NONE) && gemv_fits(B, V, h, ACT_NONE);
    const bool f8 = sk && m->fp8_decode;  // e4m3 weights x e4m3 activations through the same stream-K GEMM
    for (int l = 0; l < d.layers; ++l) {
        LlamaLayer& L = m->ll[l];
        if (f8) {
            B2_TRY(rmsnorm_quant_e4m3(m->x.p, h, L.ln1.p, m->xq8.p, h, m->xscale.as<float>(), B, h, d.rms_eps, st));
            B2_TRY(skinny8(m, kv, L.wqkv8.p, L.s_qkv.as<float>(), nullptr, 0, m->qkv.p, 3 * h, 0, B, 3 * h, h, ACT_NONE, st));
        } else if (small) {
            B2_TRY(gemv(m->x.p, h, L.wqkv.p, h, L.ln1.p, d.rms_eps, nullptr, 0, m->qkv.p, 3 * h, 0, B, 3 * h, h,
                        ACT_NONE, st));
        } else if (sk) {
            B2_TRY(rmsnorm_bf16(m->x.p, h, L.ln1.p, m->xn.p, B, h, d.rms_eps, st));
            B2_TRY(skinny(kv, m->xn.p, h, L.wqkv.p, h, nullptr, 0, m->qkv.p, 3 * h, 0, B, 3 * h, h, ACT_NONE, st));
        } else {
            B2_TRY(rmsnorm_bf16(m->x.p, h, L.ln1.p, m->xn.p, B, h, d.rms_eps, st));
            B2_TRY(gemm(m->xn.p, h, L.wqkv.p, h, nullptr, nullptr, 0, m->qkv.p, 3 * h, 0, B, 3 * h, h, ACT_NONE, st));
        }
        DecodeAttnArgs da;
        da.qkv = m->qkv.p;
        da.kcache = kv->k.as<bf16>() + (size_t)l * kv->layer_stride();
        da.vcache = kv->v.as<bf16>() + (size_t)l * kv->layer_stride();
        da.cur_len = kv->len_dev.as<int32_t>();
        da.out = m->attn.p;
        da.partial = kv->attn_partial.as<float>();
        da.counters = kv->attn_counters.as<int32_t>();
        da.B = B; da.H = H; da.D = m->hd; da.Smax = kv->max_seq; da.nsplit = nsplit;
        da.theta = d.rope_theta;
        da.scale = 1.0f / sqrtf((float)m->hd);
        B2_TRY(decode_attn_bf16(da, st));
        if (f8) {
            B2_TRY(quantize_rows_e4m3(m->attn.p, h, B, h, m->xq8.p, h, m->xscale.as<float>(), st));
            B2_TRY(skinny8(m, kv, L.wo8.p, L.s_o.as<float>(), m->x.p, h, m->x.p, h, 0, B, h, h, ACT_NONE, st));
            B2_TRY(rmsnorm_quant_e4m3(m->x.p, h, L.ln2.p, m->xq8.p, h, m->xscale.as<float>(), B, h, d.rms_eps, st));
            B2_TRY(skinny8(m, kv, L.wgu8.p, L.s_gu.as<float>(), nullptr, 0, m->act.p, I, 0, B, 2 * I, h, ACT_SWIGLU, st));
            B2_TRY(quantize_rows_e4m3(m->act.p, I, B, I, m->xq8.p, I, m->xscale.as<float>(), st));
            B2_TRY(skinny8(m, kv, L.wd8.p, L.s_d.as<float>(), m->x.p, h, m->x.p, h, 0, B, h, I, ACT_NONE, st));
        } else if (small) {
            B2_TRY(gemv(m->attn.p, h, L.wo.p, h, nullptr, 0.f, m->x.p, h, m->x.p, h, 0, B, h, h, ACT_NONE, st));
            B2_TRY(gemv(m->x.p, h, L.wgu.p, h, L.ln2.p, d.rms_eps, nullptr, 0, m->act.p, I, 0, B, 2 * I, h,
                        ACT_SWIGLU, st));
            B2_TRY(gemv(m->act.p, I, L.wd.p, I, nullptr, 0.f, m->x.p, h, m->x.p, h, 0, B, h, I, ACT_NONE, st));
        } else if (sk) {
            B2_TRY(skinny(kv, m->attn.p, h, L.wo.p, h, m->x.p, h, m->x.p, h, 0, B, h, h, ACT_NONE, st));
            B2_TRY(rmsnorm_bf16(m->x.p, h, L.ln2.p, m->xn.p, B, h, d.rms_eps, st));
            B2_TRY(skinny(kv, m->xn.p, h, L.wgu.p, h, nullptr, 0, m->act.p, I, 0, B, 2 * I, h, ACT_SWIGLU, st));
            B2_TRY(skinny(kv, m->act.p, I, L.wd.p, I, m->x.p, h, m->x.p, h, 0, B, h, I, ACT_NONE, st));
        } else {
            B2_TRY(gemm(m->attn.p, h, L.wo.p, h, nullptr, m->x.p, h, m->x.p, h, 0, B, h, h, ACT_NONE, st));
            B2_TRY(rmsnorm_bf16(m->x.p, h, L.ln2.p, m->xn.p, B, h, d.rms_eps, st));
            B2_TRY(gemm(m->xn.p, h, L.wgu.p, h, nullptr, nullptr, 0, m->act.p, I, 0, B, 2 * I, h, ACT_SWIGLU, st));
            B2_TRY(gemm(m->act.p, I, L.wd.p, I, nullptr, m->x.p, h, m->x.p, h, 0, B, h, I, ACT_NONE, st));
        }
    }
    if (f8) {
        B2_TRY(rmsnorm_quant_e4m3(m->x.p, h, m->final_norm.p, m->xq8.p, h, m->xscale.as<float>(), B, h, d.rms_eps, st));
        B2_TRY(skinny8(m, kv, m->lm_head8.p, m->s_head.as<float>(), nullptr, 0, m->logits.p, V, 1, B, V, h, ACT_NONE, st));
    } else if (small) {
        B2_TRY(gemv(m->x.p, h, m->lm_head.p, h, m->final_norm.p, d.rms_eps, nullptr, 0, m->logits.p, V, 1, B, V, h,
                    ACT_NONE, st));
    } else if (sk) {
        B2_TRY(rmsnorm_bf16(m->x.p, h, m->final_norm.p, m->xn.p, B, h, d.rms_eps, st));
        B2_TRY(skinny(kv, m->xn.p, h, m->lm_head.p, h, nullptr, 0, m->logits.p, V, 1, B, V, h, ACT_NONE, st));
    } else {
        B2_TRY(rmsnorm_bf16(m->x.p, h, m->final_norm.p, m->xn.p, B, h, d.rms_eps, st));
        B2_TRY(gemm(m->xn.p, h, m->lm_head.p, h, nullptr, nullptr, 0, m->logits.p, V, 1, B, V, h, ACT_NONE, st));
    }
    {   // A/B switch for measurements: the round-1 tail (argmax + store_token + 2 x add_i32), greedy only, no host ring
        const char* e = getenv("B2_SAMPLE_LEGACY");
        if (e != nullptr && e[0] == '1' && kv->samp_host.do_sample == 0 && kv->samp_host.tag == 0 && kv->samp_host.per_row == 0) {
            B2_TRY(argmax_f32(m->logits.as<float>(), B, V, kv->tok.as<int32_t>(), st));
            B2_TRY(store_token(kv->tok.as<int32_t>(), kv->out_tokens.as<int32_t>(), kv->step_counter.as<int32_t>(), B, st));
            B2_TRY(add_i32(kv->step_counter.as<int32_t>(), 1, 1, st));
            B2_TRY(add_i32(kv->len_dev.as<int32_t>(), B, 1, st));
            return 0;
        }
    }
    // argmax or temperature/top-k/top-p draw (device-resident SampleState), token feedback, host-ring publication and the
    // step / cache-length counters in ONE launch (was: argmax + store_token + 2 x add_i32)
    B2_TRY(sample_publish(m->logits.as<float>(), V, B, kv->sstate.as<SampleState>(), kv->rows_dev.as<RowState>(), kv->tok.as<int32_t>(),
                          kv->out_tokens.as<int32_t>(), kv->step_counter.as<int32_t>(), kv->len_dev.as<int32_t>(),
                          kv->ring_dev, kv->ring_cap, SP_SELECT | SP_WRITE_OUT | SP_BUMP, 0, st));
    return 0;
}

// caller stream -> stream the decode steps run on (fork), and back (join)
int fork_stream(b2_kv* kv, cudaStream_t st, cudaStream_t* run) {
    if (st != nullptr && st != cudaStreamLegacy) { *run = st; return 0; }
    if (kv->own_stream == nullptr) {
        B2_CUDA_CHECK(cudaStreamCreateWithFlags(&kv->own_stream, cudaStreamNonBlocking));
        B2_CUDA_CHECK(cudaEventCreateWithFlags(&kv->ev_fork, cudaEventDisableTiming));
        B2_CUDA_CHECK(cudaEventCreateWithFlags(&kv->ev_join, cudaEventDisableTiming));
    }
    B2_CUDA_CHECK(cudaEventRecord(kv->ev_fork, st));
    B2_CUDA_CHECK(cudaStreamWaitEvent(kv->own_stream, kv->ev_fork, 0));
    *run = kv->own_stream;
    return 0;
}
int join_stream(b2_kv* kv, cudaStream_t st, cudaStream_t run) {
    if (run == st) return 0;
    B2_CUDA_CHECK(cudaEventRecord(kv->ev_join, run));
    B2_CUDA_CHECK(cudaStreamWaitEvent(st, kv->ev_join, 0));
    return 0;
}

// run one step, through the cached CUDA graph when possible
bool use_mega(const b2_model* m, int B) {
    if (!decode_mega_fits(B, m->d.hidden, m->d.inter) || m->d.layers > 48) return false;
    static int flag = -1;
    if (flag < 0) {
        const char* e = getenv("B2_DECODE_MEGA");
        flag = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return flag == 1;
}

// defaults of the megakernel knobs (measured on B200: profiles/r1e_mega_sweep.txt)
constexpr int kMegaL2AheadDefault = 0;
constexpr int kMegaFastPrologueDefault = 0;

int decode_step_mega(b2_model* m, b2_kv* kv, int B, cudaStream_t st) {
    const b2_model_desc& d = m->d;
    MegaParams p;
    p.layers = kv->mega_layers.as<MegaLayer>();
    p.L = d.layers; p.h = d.hidden; p.I = d.inter; p.H = d.heads; p.V = d.vocab; p.B = B; p.Smax = kv->max_seq;
    int ns = (num_sms() * 16) / (B * d.heads);
    p.nsplit = ns < 1 ? 1 : (ns > 64 ? 64 : ns);
    p.embed = m->embed.as<bf16>(); p.final_norm = m->final_norm.as<bf16>(); p.lm_head = m->lm_head.as<bf16>();
    p.tok = kv->tok.as<int32_t>(); p.cur_len = kv->len_dev.as<int32_t>();
    p.out_tokens = kv->out_tokens.as<int32_t>(); p.step_counter = kv->step_counter.as<int32_t>();
    p.x = m->x.as<bf16>(); p.qkv = m->qkv.as<bf16>(); p.attn = m->attn.as<bf16>(); p.act = m->act.as<bf16>();
    p.logits = m->logits.as<float>();
    p.attn_partial = kv->attn_partial.as<float>(); p.attn_counters = kv->attn_counters.as<int32_t>();
    unsigned int* sync = kv->mega_sync.as<unsigned int>();
    p.bar_count = sync; p.done_count = sync + 2;
    p.bar_base = kv->mega_bar_base;
    kv->mega_bar_base += (unsigned int)(5 * d.layers + 2) * (unsigned int)num_sms();
    p.eps = d.rms_eps; p.theta = d.rope_theta;
    p.scale_log2 = (1.0f / sqrtf((float)m->hd)) * 1.4426950408889634f;
    // greedy streaming: the kernel's fused argmax publishes to the host ring itself; with do_sample the sample_publish
    // launch below overrides the fused argmax (token feedback, out_tokens slot) and publishes instead
    const bool sampling = kv->samp_host.do_sample != 0 || kv->samp_host.per_row != 0;
    if (!sampling && kv->samp_host.tag != 0) { p.sstate = kv->sstate.as<SampleState>(); p.ring = kv->ring_dev; p.ring_cap = kv->ring_cap; }
    if (kv->samp_host.per_row) p.rows = kv->rows_dev.as<RowState>();
    {   // tuning knobs, re-read every launch so a sweep can flip them inside one process (scripts/mega_sweep.py)
        const char* e = getenv("B2_MEGA_L2_AHEAD");
        int ahead = e ? atoi(e) : kMegaL2AheadDefault;
        p.l2_ahead = ahead < 0 ? 0 : (ahead > 64 ? 64 : ahead) / 4 * 4;
        e = getenv("B2_MEGA_L2_MODE");
        p.l2_mode = (e && e[0] == '2') ? 2 : 1;
        e = getenv("B2_MEGA_FAST_PROLOGUE");
        p.fast_prologue = e ? (e[0] != '0') : kMegaFastPrologueDefault;
        e = getenv("B2_MEGA_GAMMA_SMEM");
        p.gamma_smem = e ? (e[0] != '0') : 1;  // measured: 2.854 -> 2.820 ms/token (profiles/r2g_mega_sweep.jsonl)
    }
    static int trace_mode = -1;
    if (trace_mode < 0) { const char* e = getenv("B2_MEGA_TRACE"); trace_mode = (e && e[0] == '1') ? 1 : 0; }
    if (trace_mode == 1) {
        // debug: per-phase SM-clock timestamps of CTA 0 for one step, dumped to gpurun_out/mega_trace.txt
        static int steps_seen = 0;
        if (++steps_seen == 40) {
            const int n = (5 * d.layers + 2) * 4;
            DevBuf tb;
            B2_TRY(tb.alloc((size_t)n * 8));
            cudaMemset(tb.p, 0, (size_t)n * 8);
            p.trace = tb.as<long long>();
            int r = decode_mega(p, st);
            cudaStreamSynchronize(st);
            std::vector<long long> host(n);
            cudaMemcpy(host.data(), tb.p, (size_t)n * 8, cudaMemcpyDeviceToHost);
            FILE* f = fopen("gpurun_out/mega_trace.txt", "w");
            if (f) {
                for (int ph = 0; ph <= 5 * d.layers; ++ph)
                    fprintf(f, "%d %lld %lld %lld %lld %lld\n", ph, host[ph * 4], host[ph * 4 + 1], host[ph * 4 + 2],
                            host[ph * 4 + 3], host[(ph + 1) * 4]);
                fclose(f);
            }
            tb.free();
            return r;
        }
    }
    B2_TRY(decode_mega(p, st));
    if (sampling)
        B2_TRY(sample_publish(m->logits.as<float>(), d.vocab, B, kv->sstate.as<SampleState>(), kv->rows_dev.as<RowState>(), kv->tok.as<int32_t>(),
                              kv->out_tokens.as<int32_t>(), kv->step_counter.as<int32_t>(), kv->len_dev.as<int32_t>(),
                              kv->ring_dev, kv->ring_cap, SP_SELECT | SP_WRITE_OUT, -1, st));
    return 0;
}

int decode_step_run(b2_model* m, b2_kv* kv, int B, cudaStream_t st) {
    // B <= 8: one persistent cooperative launch per token (no graph needed: launches queue asynchronously)
    if (use_mega(m, B)) return decode_step_mega(m, kv, B, st);
    if (kv->warm_B != B) {
        // first step for this batch size runs eagerly: sets function attributes, resolves driver entry points
        if (kv->graph) { cudaGraphExecDestroy(kv->graph); kv->graph = nullptr; kv->graph_B = 0; }
        B2_TRY(decode_step_launch(m, kv, B, st));
        kv->warm_B = B;
        return 0;
    }
    if (kv->graph == nullptr || kv->graph_B != B) {
        if (kv->graph) { cudaGraphExecDestroy(kv->graph); kv->graph = nullptr; }
        cudaGraph_t graph = nullptr;
        B2_CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        const unsigned long long launches_before = g_launch_count;
        int r = decode_step_launch(m, kv, B, st);
        g_launch_count = launches_before;  // capture records launches, it does not run them
        cudaError_t e = cudaStreamEndCapture(st, &graph);
        if (r != 0) { if (graph) cudaGraphDestroy(graph); return r; }
        B2_CUDA_CHECK(e);
        e = cudaGraphInstantiate(&kv->graph, graph, 0);
        cudaGraphDestroy(graph);
        B2_CUDA_CHECK(e);
        kv->graph_B = B;
    }
    B2_CUDA_CHECK(cudaGraphLaunch(kv->graph, st));
    // kernels per step: embed + L*(qkv, attn, o, gate/up, down [+2 norms when B>8]) + head(+norm) + sample_publish
    const bool small = !use_skinny(kv, B) && B <= 8 && gemv_fits(B, m->d.hidden, m->d.inter, ACT_NONE) &&
                       gemv_fits(B, m->d.vocab, m->d.hidden, ACT_NONE);
    const int per_layer = small ? 5 : ((m->fp8_decode && use_skinny(kv, B)) ? 9 : 7);
    g_launch_count += 1 + (unsigned long long)m->d.layers * per_layer + (small ? 1 : 2) + 1;
    return 0;
}

// makes `dev` current for the duration of an ABI call and restores the caller's device afterwards
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev); else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// Cross-stream ordering of the shared model workspaces (see b2_model::ws_event). Caller holds the model lock.
int ws_enter(b2_model* m, cudaStream_t st) {
    if (m->ws_valid && m->ws_stream != st) B2_CUDA_CHECK(cudaStreamWaitEvent(st, m->ws_event, 0));
    return 0;
}
int ws_leave(b2_model* m, cudaStream_t st) {
    B2_CUDA_CHECK(cudaEventRecord(m->ws_event, st));
    m->ws_stream = st;
    m->ws_valid = true;
    return 0;
}

}  // namespace

// ==========================================================================================================
// C ABI
// ==========================================================================================================
extern "C" {

int b2_init(int device) {
    B2_CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    B2_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("b2_init: device %d is sm_%d%d; this library only contains sm_100a code (no fallback path)", device,
                  prop.major, prop.minor);
        return -3;
    }
    return 0;
}

const char* b2_last_error(void) { return g_err; }
int b2_version(void) { return 1; }
unsigned long long b2_launch_count(void) { return g_launch_count; }

int b2_model_create(const b2_model_desc* desc, b2_model** out) {
    B2_CHECK_ARG(desc != nullptr && out != nullptr, "b2_model_create: null argument");
    const b2_model_desc& d = *desc;
    B2_CHECK_ARG(d.vit_hidden > 0 && d.vit_heads > 0 && d.vit_hidden / d.vit_heads == 64 &&
                     d.vit_hidden % 256 == 0 && d.vit_hidden <= 2048,
                 "b2_model_create: vision head_dim must be 64 and hidden a multiple of 256 (hidden=%d heads=%d)",
                 d.vit_hidden, d.vit_heads);
    B2_CHECK_ARG(d.hidden > 0 && d.heads > 0 && d.hidden / d.heads == 128 && d.hidden % 256 == 0,
                 "b2_model_create: decoder head_dim must be 128 and hidden a multiple of 256 (hidden=%d heads=%d)",
                 d.hidden, d.heads);
    B2_CHECK_ARG(d.inter % 256 == 0 && d.vit_inter % 8 == 0 && d.vocab % 8 == 0,
                 "b2_model_create: inter %% 256, vit_inter %% 8, vocab %% 8 must be 0");
    B2_CHECK_ARG(d.image_size % d.patch_size == 0, "b2_model_create: image_size not a multiple of patch_size");
    B2_CHECK_ARG(d.max_batch >= 1 && d.max_seq >= 1 && d.max_images >= 1, "b2_model_create: bad workspace sizing");
    const int live = d.vit_select_layer < 0 ? d.vit_layers + 1 + d.vit_select_layer : d.vit_select_layer;
    B2_CHECK_ARG(live >= 0 && live <= d.vit_layers, "b2_model_create: vit_select_layer %d out of range",
                 d.vit_select_layer);
    b2_model* m = new b2_model();
    m->d = d;
    cudaGetDevice(&m->device);
    const int g = d.image_size / d.patch_size;
    m->P = g * g;
    m->T = m->P + 1;
    const int kk = 3 * d.patch_size * d.patch_size;
    m->kpad = (kk + 7) / 8 * 8;
    m->vit_live = live;  // hidden_states[live] = output of encoder layer `live` (0 = embeddings)
    m->vit.resize(live);
    m->ll.resize(d.layers);
    *out = m;
    return 0;
}

int b2_model_set_weight(b2_model* m, const char* hf_key, const void* ptr, const int64_t* shape, int ndim, int dtype) {
    B2_CHECK_ARG(m && hf_key && ptr && shape, "b2_model_set_weight: null argument");
    B2_CHECK_ARG(dtype == DT_BF16 || dtype == DT_F16 || dtype == DT_F32, "b2_model_set_weight: bad dtype %d", dtype);
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    std::string k(hf_key);
    const int h = m->d.hidden, V = m->d.vocab, D = m->d.vit_hidden;
    int r = -1;
    size_t pos;
    if ((pos = k.find("vision_model.")) != std::string::npos) {
        r = set_vision_weight(m, k.substr(pos + strlen("vision_model.")), ptr, shape, ndim, dtype);
    } else if (k == "model.embed_tokens.weight") {
        B2_TRY(expect_shape(hf_key, shape, ndim, V, h));
        r = put(m->embed, (size_t)V * h, 0, ptr, dtype, (int64_t)V * h);
    } else if (k == "model.norm.weight") {
        B2_TRY(expect_shape(hf_key, shape, ndim, h));
        r = put(m->final_norm, h, 0, ptr, dtype, h);
    } else if (k == "lm_head.weight") {
        B2_TRY(expect_shape(hf_key, shape, ndim, V, h));
        r = put(m->lm_head, (size_t)V * h, 0, ptr, dtype, (int64_t)V * h);
    } else if (k == "model.mm_projector.0.weight") {
        B2_TRY(expect_shape(hf_key, shape, ndim, h, D));
        r = put(m->p0_w, (size_t)h * D, 0, ptr, dtype, (int64_t)h * D);
    } else if (k == "model.mm_projector.0.bias") {
        B2_TRY(expect_shape(hf_key, shape, ndim, h));
        r = put(m->p0_b, h, 0, ptr, dtype, h);
    } else if (k == "model.mm_projector.2.weight") {
        B2_TRY(expect_shape(hf_key, shape, ndim, h, h));
        r = put(m->p2_w, (size_t)h * h, 0, ptr, dtype, (int64_t)h * h);
    } else if (k == "model.mm_projector.2.bias") {
        B2_TRY(expect_shape(hf_key, shape, ndim, h));
        r = put(m->p2_b, h, 0, ptr, dtype, h);
    } else if (starts_with(k, "model.layers.")) {
        int li = -1, consumed = 0;
        if (sscanf(hf_key, "model.layers.%d.%n", &li, &consumed) != 1 || li < 0 || li >= m->d.layers) {
            set_error("set_weight: bad decoder layer index in '%s'", hf_key);
            return -1;
        }
        r = set_llama_layer_weight(m, li, k.substr(consumed), hf_key, ptr, shape, ndim, dtype);
    } else {
        set_error("set_weight: unknown key '%s'", hf_key);
        return -1;
    }
    return r;
}

int b2_model_finalize(b2_model* m) {
    B2_CHECK_ARG(m != nullptr, "b2_model_finalize: null model");
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    const b2_model_desc& d = m->d;
    // completeness
    std::string missing;
    auto need = [&](const DevBuf& b, const char* name) {
        if (b.p == nullptr) { if (missing.size() < 600) { missing += name; missing += ' '; } }
    };
    need(m->patch_w, "patch_embedding"); need(m->cls, "class_embedding"); need(m->pos, "position_embedding");
    need(m->pre_g, "pre_layrnorm.weight"); need(m->pre_b, "pre_layrnorm.bias");
    for (size_t i = 0; i < m->vit.size(); ++i) {
        VitLayer& L = m->vit[i];
        const DevBuf* bs[12] = {&L.ln1_g, &L.ln1_b, &L.wqkv, &L.bqkv, &L.wo, &L.bo, &L.ln2_g, &L.ln2_b, &L.w1, &L.b1, &L.w2, &L.b2};
        for (int j = 0; j < 12; ++j)
            if (bs[j]->p == nullptr) { char t[64]; snprintf(t, sizeof t, "vision.layer%zu.#%d", i, j); need(*bs[j], t); }
    }
    need(m->p0_w, "mm_projector.0.weight"); need(m->p0_b, "mm_projector.0.bias");
    need(m->p2_w, "mm_projector.2.weight"); need(m->p2_b, "mm_projector.2.bias");
    need(m->embed, "embed_tokens"); need(m->final_norm, "model.norm"); need(m->lm_head, "lm_head");
    for (size_t i = 0; i < m->ll.size(); ++i) {
        LlamaLayer& L = m->ll[i];
        const DevBuf* bs[6] = {&L.ln1, &L.wqkv, &L.wo, &L.ln2, &L.wgu, &L.wd};
        for (int j = 0; j < 6; ++j)
            if (bs[j]->p == nullptr) { char t[64]; snprintf(t, sizeof t, "layers.%zu.#%d", i, j); need(*bs[j], t); }
    }
    if (!missing.empty()) {
        set_error("b2_model_finalize: missing weights: %s", missing.c_str());
        return -3;
    }
    // fused buffers are filled piecewise (q/k/v): every part must have arrived, or the buffer holds uninitialised memory
    // (gate/up is only materialised once both halves are there, so its pointer check above is already exact)
    for (size_t i = 0; i < m->vit.size(); ++i)
        if (m->vit[i].qkv_w != 7u || m->vit[i].qkv_b != 7u) {
            set_error("b2_model_finalize: vision layer %zu is missing q/k/v parts (weights mask %u, bias mask %u of 7)", i,
                      m->vit[i].qkv_w, m->vit[i].qkv_b);
            return -3;
        }
    for (size_t i = 0; i < m->ll.size(); ++i)
        if (m->ll[i].qkv_parts != 7u) {
            set_error("b2_model_finalize: decoder layer %zu is missing q/k/v projections (mask %u of 7)", i, m->ll[i].qkv_parts);
            return -3;
        }
    if (m->err_host == nullptr) {
        B2_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&m->err_host), 8 * sizeof(int), cudaHostAllocMapped));
        memset(m->err_host, 0, 8 * sizeof(int));
        B2_CUDA_CHECK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&m->err_dev), m->err_host, 0));
        B2_CUDA_CHECK(cudaEventCreateWithFlags(&m->ws_event, cudaEventDisableTiming));
    }
    // workspaces
    const int D = d.vit_hidden, I = d.vit_inter, h = d.hidden;
    const size_t vrows = (size_t)d.max_images * m->T, prow = (size_t)d.max_images * m->P;
    B2_TRY(m->v_col.alloc(prow * m->kpad * 2));
    B2_TRY(m->v_patch.alloc(prow * D * 2));
    B2_TRY(m->v_hidden.alloc(vrows * D * 2));
    B2_TRY(m->v_xn.alloc(vrows * D * 2));
    B2_TRY(m->v_qkv.alloc(vrows * 3 * D * 2));
    B2_TRY(m->v_attn.alloc(vrows * D * 2));
    B2_TRY(m->v_mlp.alloc(vrows * I * 2));
    B2_TRY(m->v_feats.alloc(prow * D * 2));
    B2_TRY(m->p_mid.alloc(prow * h * 2));
    B2_TRY(m->p_done.alloc(((prow + 127) / 128 + 1) * sizeof(int)));
    B2_TRY(m->enc_pixels.alloc((size_t)d.max_images * 3 * d.image_size * d.image_size * 2));
    B2_TRY(m->enc_out.alloc(prow * h * 2));
    m->enc_graph.assign(d.max_images + 1, nullptr);
    m->enc_warm.assign(d.max_images + 1, 0);
    const size_t rows = (size_t)d.max_batch * d.max_seq;
    B2_TRY(m->x.alloc(rows * h * 2));
    B2_TRY(m->xn.alloc(rows * h * 2));
    B2_TRY(m->qkv.alloc(rows * 3 * h * 2));
    B2_TRY(m->attn.alloc(rows * h * 2));
    B2_TRY(m->act.alloc(rows * d.inter * 2));
    B2_TRY(m->last_idx.alloc((size_t)d.max_batch * 4));
    B2_TRY(m->splice_idx.alloc(rows * 4));
    B2_TRY(m->xlast.alloc((size_t)d.max_batch * h * 2));
    B2_TRY(m->logits.alloc((size_t)d.max_batch * d.vocab * 4));
    m->finalized = true;
    return 0;
}

int b2_model_destroy(b2_model* m) {
    if (m == nullptr) return 0;
    DeviceGuard dg(m->device);
    cudaDeviceSynchronize();
    if (m->err_host) cudaFreeHost(m->err_host);
    if (m->ws_event) cudaEventDestroy(m->ws_event);
    for (cudaGraphExec_t g : m->enc_graph) if (g) cudaGraphExecDestroy(g);
    if (m->enc_stream) cudaStreamDestroy(m->enc_stream);
    if (m->enc_fork) cudaEventDestroy(m->enc_fork);
    if (m->enc_join) cudaEventDestroy(m->enc_join);
    DevBuf* top[] = {&m->patch_w, &m->cls, &m->pos, &m->pre_g, &m->pre_b, &m->p0_w, &m->p0_b, &m->p2_w, &m->p2_b,
                     &m->embed, &m->final_norm, &m->lm_head, &m->v_col, &m->v_patch, &m->v_hidden, &m->v_xn,
                     &m->v_qkv, &m->v_attn, &m->v_mlp, &m->v_feats, &m->p_mid, &m->p_done, &m->enc_pixels, &m->enc_out, &m->x, &m->xn, &m->qkv, &m->attn,
                     &m->act, &m->last_idx, &m->xlast, &m->logits, &m->splice_idx, &m->lm_head8, &m->s_head, &m->xq8, &m->xscale};
    for (DevBuf* b : top) b->free();
    for (VitLayer& L : m->vit) {
        DevBuf* bs[12] = {&L.ln1_g, &L.ln1_b, &L.wqkv, &L.bqkv, &L.wo, &L.bo, &L.ln2_g, &L.ln2_b, &L.w1, &L.b1, &L.w2, &L.b2};
        for (DevBuf* b : bs) b->free();
    }
    for (LlamaLayer& L : m->ll) {
        DevBuf* bs[16] = {&L.ln1, &L.wqkv, &L.wo, &L.ln2, &L.wgu, &L.wd, &L.tmp_gate, &L.tmp_up,
                          &L.wqkv8, &L.wo8, &L.wgu8, &L.wd8, &L.s_qkv, &L.s_o, &L.s_gu, &L.s_d};
        for (DevBuf* b : bs) b->free();
    }
    delete m;
    return 0;
}

// BASELINE configs[4]: quantise every decode Linear to e4m3 with per-output-channel scales (the bf16 copies stay: prefill
// and the B <= 6 decode paths keep using them). Decode steps at batch >= 7 then stream 1-byte weights.
int b2_model_enable_fp8_decode(b2_model* m) {
    B2_CHECK_ARG(m != nullptr, "b2_model_enable_fp8_decode: null model");
    B2_CHECK_ARG(m->finalized, "b2_model_enable_fp8_decode: model not finalized");
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    if (m->fp8_decode) return 0;
    const b2_model_desc& d = m->d;
    const int h = d.hidden, I = d.inter, V = d.vocab;
    B2_CHECK_ARG(h % 16 == 0 && I % 16 == 0, "b2_model_enable_fp8_decode: hidden/inter must be multiples of 16");
    auto quant = [&](DevBuf& w, int N, int K, DevBuf& w8, DevBuf& sc) -> int {
        B2_TRY(w8.alloc((size_t)N * K));
        B2_TRY(sc.alloc((size_t)N * sizeof(float)));
        return quantize_rows_e4m3(w.p, K, N, K, w8.p, K, sc.as<float>(), nullptr);
    };
    for (LlamaLayer& L : m->ll) {
        B2_TRY(quant(L.wqkv, 3 * h, h, L.wqkv8, L.s_qkv));
        B2_TRY(quant(L.wo, h, h, L.wo8, L.s_o));
        B2_TRY(quant(L.wgu, 2 * I, h, L.wgu8, L.s_gu));  // rows stay block-64 interleaved; scales follow the physical rows
        B2_TRY(quant(L.wd, h, I, L.wd8, L.s_d));
    }
    B2_TRY(quant(m->lm_head, V, h, m->lm_head8, m->s_head));
    const int mb = d.max_batch > 128 ? 128 : d.max_batch;
    B2_TRY(m->xq8.alloc((size_t)mb * (h > I ? h : I)));
    B2_TRY(m->xscale.alloc((size_t)mb * sizeof(float)));
    B2_CUDA_CHECK(cudaDeviceSynchronize());
    m->fp8_decode = true;
    return 0;
}

int b2_kv_create(b2_model* m, int max_batch, int max_seq, b2_kv** out) {
    B2_CHECK_ARG(m && out, "b2_kv_create: null argument");
    B2_CHECK_ARG(max_batch >= 1 && max_batch <= m->d.max_batch, "b2_kv_create: max_batch %d exceeds model max_batch %d",
                 max_batch, m->d.max_batch);
    B2_CHECK_ARG(max_seq >= 1, "b2_kv_create: max_seq must be positive");
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    b2_kv* kv = new b2_kv();
    kv->m = m;
    kv->max_batch = max_batch;
    kv->max_seq = max_seq;
    const size_t per = (size_t)m->d.layers * kv->layer_stride() * 2;
    int r = 0;
    if ((r = kv->k.alloc(per)) != 0 || (r = kv->v.alloc(per)) != 0 ||
        (r = kv->len_dev.alloc((size_t)max_batch * 4)) != 0 || (r = kv->tok.alloc((size_t)max_batch * 4)) != 0 ||
        (r = kv->step_counter.alloc(4)) != 0) {
        b2_kv_destroy(kv);
        return r;
    }
    kv->out_capacity = max_seq;
    const int max_split = 64;
    {
        std::vector<MegaLayer> tbl(m->d.layers);
        for (int l = 0; l < m->d.layers; ++l) {
            LlamaLayer& L = m->ll[l];
            tbl[l].ln1 = L.ln1.as<bf16>(); tbl[l].wqkv = L.wqkv.as<bf16>(); tbl[l].wo = L.wo.as<bf16>();
            tbl[l].ln2 = L.ln2.as<bf16>(); tbl[l].wgu = L.wgu.as<bf16>(); tbl[l].wd = L.wd.as<bf16>();
            tbl[l].kcache = kv->k.as<bf16>() + (size_t)l * kv->layer_stride();
            tbl[l].vcache = kv->v.as<bf16>() + (size_t)l * kv->layer_stride();
        }
        if ((r = kv->mega_layers.alloc(tbl.size() * sizeof(MegaLayer))) != 0 || (r = kv->mega_sync.alloc(64)) != 0) {
            b2_kv_destroy(kv);
            return r;
        }
        cudaMemcpy(kv->mega_layers.p, tbl.data(), tbl.size() * sizeof(MegaLayer), cudaMemcpyHostToDevice);
        cudaMemset(kv->mega_sync.p, 0, 64);
    }
    if ((r = kv->out_tokens.alloc((size_t)kv->out_capacity * max_batch * 4)) != 0 ||
        (r = kv->attn_partial.alloc(((size_t)max_batch * m->d.heads * max_split + 1024) * (128 + 2) * 4)) != 0 ||
        (r = kv->attn_counters.alloc((size_t)max_batch * m->d.heads * 4)) != 0) {
        b2_kv_destroy(kv);
        return r;
    }
    if (max_batch >= 3 && max_batch <= 128) {
        const int h = m->d.hidden, I = m->d.inter, V = m->d.vocab;
        size_t ws = 0;
        const int shapes[5][2] = {{3 * h, h}, {h, h}, {2 * I, h}, {h, I}, {V, h}};
        int nmax = 0;
        for (auto& sh : shapes) {
            const size_t b = gemm_skinny_workspace_bytes(max_batch, sh[0], sh[1]);
            ws = b > ws ? b : ws;
            nmax = sh[0] > nmax ? sh[0] : nmax;
        }
        if ((r = kv->sk_partial.alloc(ws)) != 0 || (r = kv->sk_counters.alloc(gemm_skinny_counter_bytes(nmax))) != 0) {
            b2_kv_destroy(kv);
            return r;
        }
        cudaMemset(kv->sk_counters.p, 0, gemm_skinny_counter_bytes(nmax));
    }
    kv->ring_cap = max_seq;
    if ((r = kv->sstate.alloc(sizeof(SampleState))) != 0 || (r = kv->rows_dev.alloc((size_t)max_batch * sizeof(RowState))) != 0) {
        b2_kv_destroy(kv);
        return r;
    }
    cudaMemset(kv->sstate.p, 0, sizeof(SampleState));
    cudaMemset(kv->rows_dev.p, 0, (size_t)max_batch * sizeof(RowState));
    kv->rows_host.assign(max_batch, RowState{});
    if (cudaHostAlloc(reinterpret_cast<void**>(&kv->ring_host), (size_t)kv->ring_cap * max_batch * 4, cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer(reinterpret_cast<void**>(&kv->ring_dev), kv->ring_host, 0) != cudaSuccess) {
        set_error("b2_kv_create: cannot allocate the pinned token ring (%d x %d)", kv->ring_cap, max_batch);
        cudaGetLastError();
        b2_kv_destroy(kv);
        return -2;
    }
    memset(kv->ring_host, 0, (size_t)kv->ring_cap * max_batch * 4);
    cudaMemset(kv->k.p, 0, per);
    cudaMemset(kv->v.p, 0, per);
    cudaMemset(kv->len_dev.p, 0, (size_t)max_batch * 4);
    cudaMemset(kv->tok.p, 0, (size_t)max_batch * 4);
    cudaMemset(kv->step_counter.p, 0, 4);
    cudaMemset(kv->attn_counters.p, 0, (size_t)max_batch * m->d.heads * 4);
    if ((r = kv->rope_tab.alloc((size_t)max_seq * (m->hd / 2) * sizeof(uint32_t))) != 0 ||
        (r = rope_table_build(kv->rope_tab.p, max_seq, m->hd, m->d.rope_theta, nullptr)) != 0) {
        b2_kv_destroy(kv);
        return r;
    }
    // The memsets above run on the legacy stream; the caller's stream may be a non-blocking one (torch side streams are) and
    // is NOT ordered against it: a prefill issued right after this call would race with the zero-fill of the cache it writes
    // (seen in scripts/decode_ab.py FRESHKV=1: different tokens on a cache that was created a moment earlier).
    B2_CUDA_CHECK(cudaDeviceSynchronize());
    kv->len_host.assign(max_batch, 0);
    *out = kv;
    return 0;
}

int b2_kv_reset(b2_kv* kv) {
    B2_CHECK_ARG(kv != nullptr, "b2_kv_reset: null");
    std::lock_guard<std::mutex> lk(kv->m->mu);
    DeviceGuard dg(kv->m->device);
    // Host bookkeeping only. The device-side lengths are rewritten by the next b2_prefill and the step counter by the next
    // decode call, both on the CALLER's stream; a cudaMemset here would run on the legacy stream, unordered against work
    // still in flight on a non-blocking caller stream (e.g. two forward() calls issued back to back).
    kv->len_host.assign(kv->max_batch, 0);
    {   // measurement knob: drop the captured decode graph so that the next run starts with an eager step + a new capture
        const char* e = getenv("B2_KV_RESET_GRAPH");
        if (e != nullptr && e[0] == '1') {
            if (kv->graph) { cudaGraphExecDestroy(kv->graph); kv->graph = nullptr; kv->graph_B = 0; }
            kv->warm_B = 0;
        }
    }
    return 0;
}

int b2_kv_destroy(b2_kv* kv) {
    if (kv == nullptr) return 0;
    DeviceGuard dg(kv->m->device);
    cudaDeviceSynchronize();
    if (kv->ring_host) cudaFreeHost(kv->ring_host);
    if (kv->graph) cudaGraphExecDestroy(kv->graph);
    if (kv->own_stream) cudaStreamDestroy(kv->own_stream);
    if (kv->ev_fork) cudaEventDestroy(kv->ev_fork);
    if (kv->ev_join) cudaEventDestroy(kv->ev_join);
    DevBuf* bs[] = {&kv->k, &kv->v, &kv->len_dev, &kv->tok, &kv->step_counter, &kv->out_tokens, &kv->attn_partial,
                    &kv->attn_counters, &kv->mega_layers, &kv->mega_sync, &kv->sk_partial, &kv->sk_counters, &kv->sstate, &kv->rows_dev, &kv->rope_tab};
    for (DevBuf* b : bs) b->free();
    delete kv;
    return 0;
}

int b2_kv_lengths(b2_kv* kv, int32_t* lens_host, int n) {
    B2_CHECK_ARG(kv && lens_host && n >= 0 && n <= kv->max_batch, "b2_kv_lengths: bad argument");
    for (int i = 0; i < n; ++i) lens_host[i] = kv->len_host[i];
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
int b2_vit_encode(b2_model* m, const void* pixels, int B, void* out_feats, void* stream) {
    B2_CHECK_ARG(m && pixels && out_feats && B >= 1, "b2_vit_encode: bad argument");
    B2_CHECK_ARG(m->finalized, "b2_vit_encode: model not finalized");
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    B2_TRY(ws_enter(m, st));
    const size_t img_elems = (size_t)3 * m->d.image_size * m->d.image_size;
    for (int b0 = 0; b0 < B; b0 += m->d.max_images) {
        const int n = B - b0 < m->d.max_images ? B - b0 : m->d.max_images;
        B2_TRY(vit_forward_chunk(m, reinterpret_cast<const bf16*>(pixels) + b0 * img_elems, n,
                                 reinterpret_cast<bf16*>(out_feats) + (size_t)b0 * m->P * m->d.vit_hidden, st));
    }
    return ws_leave(m, st);
}

int b2_project(b2_model* m, const void* feats, int rows, void* out, void* stream) {
    B2_CHECK_ARG(m && feats && out && rows >= 1, "b2_project: bad argument");
    B2_CHECK_ARG(m->finalized, "b2_project: model not finalized");
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    B2_TRY(ws_enter(m, st));
    B2_TRY(project_rows(m, feats, rows, out, st));
    return ws_leave(m, st);
}

int b2_encode_images(b2_model* m, const void* pixels, int B, void* out, void* stream) {
    B2_CHECK_ARG(m && pixels && out && B >= 1, "b2_encode_images: bad argument");
    B2_CHECK_ARG(m->finalized, "b2_encode_images: model not finalized");
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    B2_TRY(ws_enter(m, st));
    const size_t img_elems = (size_t)3 * m->d.image_size * m->d.image_size;
    for (int b0 = 0; b0 < B; b0 += m->d.max_images) {
        const int n = B - b0 < m->d.max_images ? B - b0 : m->d.max_images;
        B2_TRY(encode_chunk(m, reinterpret_cast<const bf16*>(pixels) + b0 * img_elems, n,
                            reinterpret_cast<bf16*>(out) + (size_t)b0 * m->P * m->d.hidden, st));
    }
    return ws_leave(m, st);
}

int b2_splice(b2_model* m, const int32_t* src_index, const void* image_feats, int n_feat_rows, int rows, void* embeds_out,
              void* stream) {
    B2_CHECK_ARG(m && src_index && embeds_out && rows >= 1 && n_feat_rows >= 0, "b2_splice: bad argument");
    B2_CHECK_ARG(image_feats != nullptr || n_feat_rows == 0, "b2_splice: n_feat_rows=%d without image_feats", n_feat_rows);
    B2_CHECK_ARG(m->finalized, "b2_splice: model not finalized");
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    return splice_embed(src_index, m->embed.p, image_feats, embeds_out, rows, m->d.hidden, m->d.vocab, n_feat_rows,
                        m->err_dev, reinterpret_cast<cudaStream_t>(stream));
}

int b2_splice_ids(b2_model* m, const int64_t* input_ids, int B, int Lt, int k_per_row, const int32_t* feat_offsets_host, int n_img,
                  const void* image_feats, int S, void* embeds_out, void* stream) {
    B2_CHECK_ARG(m && input_ids && feat_offsets_host && embeds_out, "b2_splice_ids: null argument");
    B2_CHECK_ARG(m->finalized, "b2_splice_ids: model not finalized");
    B2_CHECK_ARG(B >= 1 && S >= 1 && (size_t)B * S <= (size_t)m->d.max_batch * m->d.max_seq,
                 "b2_splice_ids: B*S=%d exceeds the workspace (%d rows)", B * S, m->d.max_batch * m->d.max_seq);
    B2_CHECK_ARG(image_feats != nullptr || n_img == 0, "b2_splice_ids: image slots without image_feats");
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    B2_TRY(ws_enter(m, st));
    B2_TRY(splice_index(reinterpret_cast<const long long*>(input_ids), B, Lt, k_per_row, feat_offsets_host, n_img, -200, S,
                        m->splice_idx.as<int32_t>(), m->err_dev, st));
    B2_TRY(splice_embed(m->splice_idx.as<int32_t>(), m->embed.p, image_feats, embeds_out, B * S, m->d.hidden, m->d.vocab,
                        n_img > 0 ? feat_offsets_host[n_img] : 0, m->err_dev, st));
    return ws_leave(m, st);
}

int b2_async_error(b2_model* m, int* code_out) {
    B2_CHECK_ARG(m && code_out, "b2_async_error: null argument");
    int code = 0;
    if (m->err_host != nullptr)
        for (int i = 0; i < 8; ++i)
            if (reinterpret_cast<volatile int*>(m->err_host)[i] != 0) {
                code |= 1 << i;
                reinterpret_cast<volatile int*>(m->err_host)[i] = 0;
            }
    *code_out = code;
    if (code != 0)
        set_error("device-side input check failed:%s%s%s", (code & B2_ERR_TOKEN_RANGE) ? " token id outside [0, vocab) (or an image placeholder without image features)" : "",
                  (code & B2_ERR_IMAGE_ROW_RANGE) ? " image-feature row outside the encoded images" : "",
                  (code & B2_ERR_SPLICE_SLOTS) ? " more image placeholders than images" : "");
    return 0;
}

int b2_prefill(b2_model* m, b2_kv* kv, const void* embeds, const int32_t* seq_lens_host, int B, int S,
               void* logits_out, int logits_mode, void* stream) {
    return b2_prefill_slots(m, kv, embeds, seq_lens_host, B, S, 0, logits_out, logits_mode, stream);
}

int b2_prefill_slots(b2_model* m, b2_kv* kv, const void* embeds, const int32_t* seq_lens_host, int B, int S, int slot0,
                     void* logits_out, int logits_mode, void* stream) {
    B2_CHECK_ARG(m && kv && embeds && kv->m == m, "b2_prefill: bad handle");
    B2_CHECK_ARG(m->finalized, "b2_prefill: model not finalized");
    B2_CHECK_ARG(B >= 1 && slot0 >= 0 && slot0 + B <= kv->max_batch && S >= 1 && S <= kv->max_seq,
                 "b2_prefill: B=%d S=%d slot0=%d exceed the KV cache (max_batch=%d max_seq=%d)", B, S, slot0, kv->max_batch, kv->max_seq);
    B2_CHECK_ARG((size_t)B * S <= (size_t)m->d.max_batch * m->d.max_seq,
                 "b2_prefill: B*S=%d exceeds the workspace (%d rows)", B * S, m->d.max_batch * m->d.max_seq);
    B2_CHECK_ARG(logits_mode == B2_LOGITS_NONE || logits_out != nullptr, "b2_prefill: logits_out is null");
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const b2_model_desc& d = m->d;
    const int h = d.hidden, I = d.inter, H = d.heads, V = d.vocab, T = B * S;

    std::vector<int32_t> lens(B), last(B);
    for (int b = 0; b < B; ++b) {
        lens[b] = seq_lens_host ? seq_lens_host[b] : S;
        B2_CHECK_ARG(lens[b] >= 1 && lens[b] <= S, "b2_prefill: seq_lens[%d]=%d out of range (S=%d)", b, lens[b], S);
        last[b] = b * S + lens[b] - 1;
    }
    B2_TRY(ws_enter(m, st));
    // lengths / last-row indices travel as kernel parameters: no pinned staging, no stream sync in this call
    B2_TRY(set_i32_pairs(kv->len_dev.as<int32_t>() + slot0, lens.data(), m->last_idx.as<int32_t>(), last.data(), B, st));
    for (int b = 0; b < B; ++b) kv->len_host[slot0 + b] = lens[b];
    const size_t slot_off = (size_t)slot0 * H * kv->max_seq * m->hd;  // cache slabs of the first slot this call fills

    B2_CUDA_CHECK(cudaMemcpyAsync(m->x.p, embeds, (size_t)T * h * 2, cudaMemcpyDeviceToDevice, st));
    // RoPE + KV write fused into the QKV GEMM where the CTA-pair kernel is the one that runs anyway (M >= 512; 256-column pair
    // tiles must hold whole 128-wide heads of ONE of q / k / v: hidden % 256 == 0). B2_ROPE_FUSED=0: standalone pass (A/B, tests).
    bool rope_fused = T >= 512 && m->hd == 128 && h % 256 == 0;
    { const char* e = getenv("B2_ROPE_FUSED"); if (e != nullptr && e[0] == '0') rope_fused = false; }
    for (int l = 0; l < d.layers; ++l) {
        LlamaLayer& L = m->ll[l];
        bf16* kc = kv->k.as<bf16>() + (size_t)l * kv->layer_stride() + slot_off;
        bf16* vc = kv->v.as<bf16>() + (size_t)l * kv->layer_stride() + slot_off;
        B2_TRY(rmsnorm_bf16(m->x.p, h, L.ln1.p, m->xn.p, T, h, d.rms_eps, st));
        if (rope_fused) {
            // QKV projection with RoPE and the cache write in its epilogue (CTA-pair kernel): q -> qkv buffer, k / v -> cache
            GemmArgs g;
            g.A = m->xn.p; g.lda = h; g.W = L.wqkv.p; g.ldw = h; g.out = m->qkv.p; g.ld_out = 3 * h;
            g.M = T; g.N = 3 * h; g.K = h; g.act = ACT_ROPE_QKV;
            g.rope.table = kv->rope_tab.p; g.rope.kcache = kc; g.rope.vcache = vc; g.rope.S = S; g.rope.H = H; g.rope.Smax = kv->max_seq;
            B2_TRY(gemm_bf16_2cta(g, st));
        } else {
            B2_TRY(gemm(m->xn.p, h, L.wqkv.p, h, nullptr, nullptr, 0, m->qkv.p, 3 * h, 0, T, 3 * h, h, ACT_NONE, st));
            B2_TRY(rope_kv_write(m->qkv.p, kc, vc, B, S, H, m->hd, kv->max_seq, d.rope_theta, st));
        }
        FlashArgs fa;
        fa.q = m->qkv.p; fa.q_bs = (int64_t)S * 3 * h; fa.q_ts = 3 * h; fa.q_hs = m->hd;
        fa.k = kc; fa.k_bs = (int64_t)H * kv->max_seq * m->hd; fa.k_ts = m->hd; fa.k_hs = (int64_t)kv->max_seq * m->hd;
        fa.v = vc; fa.v_bs = fa.k_bs; fa.v_ts = m->hd; fa.v_hs = fa.k_hs;
        fa.o = m->attn.p; fa.o_bs = (int64_t)S * h; fa.o_ts = h; fa.o_hs = m->hd;
        fa.seq_lens = kv->len_dev.as<int32_t>() + slot0;
        fa.B = B; fa.H = H; fa.S = S; fa.D = m->hd; fa.causal = 1;
        fa.scale = 1.0f / sqrtf((float)m->hd);
        B2_TRY(flash_attn_bf16(fa, st));
        B2_TRY(gemm(m->attn.p, h, L.wo.p, h, nullptr, m->x.p, h, m->x.p, h, 0, T, h, h, ACT_NONE, st));
        B2_TRY(rmsnorm_bf16(m->x.p, h, L.ln2.p, m->xn.p, T, h, d.rms_eps, st));
        B2_TRY(gemm(m->xn.p, h, L.wgu.p, h, nullptr, nullptr, 0, m->act.p, I, 0, T, 2 * I, h, ACT_SWIGLU, st));
        B2_TRY(gemm(m->act.p, I, L.wd.p, I, nullptr, m->x.p, h, m->x.p, h, 0, T, h, I, ACT_NONE, st));
    }
    if (logits_mode == B2_LOGITS_LAST) {
        // only the last valid position per sample feeds generation (the reference computes lm_head on all S)
        B2_TRY(rmsnorm_gather_bf16(m->x.p, m->last_idx.as<int32_t>(), m->final_norm.p, m->xlast.p, B, h, d.rms_eps, st));
        if (B <= 8 && gemv_fits(B, V, h, ACT_NONE))
            B2_TRY(gemv(m->xlast.p, h, m->lm_head.p, h, nullptr, 0.f, nullptr, 0, logits_out, V, 1, B, V, h, ACT_NONE, st));
        else
            B2_TRY(gemm(m->xlast.p, h, m->lm_head.p, h, nullptr, nullptr, 0, logits_out, V, 1, B, V, h, ACT_NONE, st));
    } else if (logits_mode == B2_LOGITS_ALL) {
        B2_TRY(rmsnorm_bf16(m->x.p, h, m->final_norm.p, m->xn.p, T, h, d.rms_eps, st));
        B2_TRY(gemm(m->xn.p, h, m->lm_head.p, h, nullptr, nullptr, 0, logits_out, V, 1, T, V, h, ACT_NONE, st));
    }
    return ws_leave(m, st);
}

static int copy_tokens_in(b2_kv* kv, const int32_t* tokens, int B, cudaStream_t st) {
    B2_CUDA_CHECK(cudaMemcpyAsync(kv->tok.p, tokens, (size_t)B * 4, cudaMemcpyDefault, st));
    return 0;
}

// device-resident selection state := v (pub_counter restarts at 0); launched only when something changed
static int set_sampling(b2_kv* kv, const SampleState& v, bool force, cudaStream_t st) {
    const SampleState& c = kv->samp_host;
    if (!force && kv->samp_valid && c.do_sample == v.do_sample && c.temperature == v.temperature && c.top_p == v.top_p &&
        c.top_k == v.top_k && c.seed == v.seed && c.tag == v.tag && c.per_row == v.per_row)
        return 0;
    B2_TRY(sample_state_set(kv->sstate.as<SampleState>(), v, st));
    kv->samp_host = v;
    kv->samp_valid = true;
    return 0;
}
static int set_greedy_unpublished(b2_kv* kv, cudaStream_t st) {
    SampleState v = {};
    v.temperature = 1.f; v.top_p = 1.f;
    kv->stream_B = 0;  // any streaming generation on this cache is over
    return set_sampling(kv, v, false, st);
}
static bool is_device_pointer(const void* p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

int b2_decode_step(b2_model* m, b2_kv* kv, const int32_t* tokens, int B, void* logits_out, int32_t* next_tokens_out,
                   void* stream) {
    B2_CHECK_ARG(m && kv && tokens && kv->m == m, "b2_decode_step: bad handle");
    B2_CHECK_ARG(m->finalized, "b2_decode_step: model not finalized");
    B2_CHECK_ARG(B >= 1 && B <= kv->max_batch, "b2_decode_step: B=%d exceeds cache batch %d", B, kv->max_batch);
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    for (int b = 0; b < B; ++b)
        B2_CHECK_ARG(kv->len_host[b] >= 1 && kv->len_host[b] < kv->max_seq,
                     "b2_decode_step: sample %d has cache length %d (capacity %d; prefill first)", b, kv->len_host[b],
                     kv->max_seq);
    B2_TRY(ws_enter(m, st));
    B2_TRY(copy_tokens_in(kv, tokens, B, st));
    B2_CUDA_CHECK(cudaMemsetAsync(kv->step_counter.p, 0, 4, st));
    B2_TRY(set_greedy_unpublished(kv, st));
    cudaStream_t run = nullptr;
    B2_TRY(fork_stream(kv, st, &run));
    B2_TRY(decode_step_run(m, kv, B, run));
    B2_TRY(join_stream(kv, st, run));
    for (int b = 0; b < B; ++b) kv->len_host[b]++;
    if (logits_out)
        B2_CUDA_CHECK(cudaMemcpyAsync(logits_out, m->logits.p, (size_t)B * m->d.vocab * 4, cudaMemcpyDefault, st));
    if (next_tokens_out)
        B2_CUDA_CHECK(cudaMemcpyAsync(next_tokens_out, kv->tok.p, (size_t)B * 4, cudaMemcpyDefault, st));
    B2_TRY(ws_leave(m, st));
    if (next_tokens_out && !is_device_pointer(next_tokens_out)) B2_CUDA_CHECK(cudaStreamSynchronize(st));
    return 0;
}

int b2_decode_greedy(b2_model* m, b2_kv* kv, const int32_t* first_tokens, int B, int n_steps, int32_t* out_tokens,
                     void* stream) {
    B2_CHECK_ARG(m && kv && first_tokens && out_tokens && kv->m == m, "b2_decode_greedy: bad handle");
    B2_CHECK_ARG(m->finalized, "b2_decode_greedy: model not finalized");
    B2_CHECK_ARG(B >= 1 && B <= kv->max_batch && n_steps >= 1, "b2_decode_greedy: bad B=%d n_steps=%d", B, n_steps);
    B2_CHECK_ARG(n_steps <= kv->out_capacity, "b2_decode_greedy: n_steps=%d exceeds capacity %d", n_steps,
                 kv->out_capacity);
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    for (int b = 0; b < B; ++b)
        B2_CHECK_ARG(kv->len_host[b] >= 1 && kv->len_host[b] + n_steps <= kv->max_seq,
                     "b2_decode_greedy: sample %d cache length %d + %d steps exceeds capacity %d", b, kv->len_host[b],
                     n_steps, kv->max_seq);
    B2_TRY(ws_enter(m, st));
    B2_TRY(copy_tokens_in(kv, first_tokens, B, st));
    B2_CUDA_CHECK(cudaMemsetAsync(kv->step_counter.p, 0, 4, st));
    B2_TRY(set_greedy_unpublished(kv, st));
    cudaStream_t run = nullptr;
    B2_TRY(fork_stream(kv, st, &run));
    for (int s = 0; s < n_steps; ++s) B2_TRY(decode_step_run(m, kv, B, run));
    B2_TRY(join_stream(kv, st, run));
    for (int b = 0; b < B; ++b) kv->len_host[b] += n_steps;
    B2_CUDA_CHECK(cudaMemcpyAsync(out_tokens, kv->out_tokens.p, (size_t)n_steps * B * 4, cudaMemcpyDefault, st));
    B2_TRY(ws_leave(m, st));
    // a device destination stays asynchronous on the caller's stream; a host destination must be complete on return
    if (!is_device_pointer(out_tokens)) B2_CUDA_CHECK(cudaStreamSynchronize(st));
    return 0;
}

// ---- streaming decode: the device runs ahead, the host reads tokens from mapped pinned memory ---------------------------
int b2_stream_begin(b2_model* m, b2_kv* kv, const float* logits, int B, const b2_sampling* sp, void* stream) {
    B2_CHECK_ARG(m && kv && logits && kv->m == m, "b2_stream_begin: bad handle");
    B2_CHECK_ARG(m->finalized, "b2_stream_begin: model not finalized");
    B2_CHECK_ARG(B >= 1 && B <= kv->max_batch, "b2_stream_begin: B=%d exceeds cache batch %d", B, kv->max_batch);
    SampleState v = {};
    v.temperature = 1.f; v.top_p = 1.f;
    if (sp != nullptr && sp->do_sample) {
        B2_CHECK_ARG(sp->temperature > 0.f, "b2_stream_begin: temperature must be positive when sampling (got %g)", (double)sp->temperature);
        B2_CHECK_ARG(sp->top_p > 0.f && sp->top_p <= 1.f, "b2_stream_begin: top_p must be in (0, 1] (got %g)", (double)sp->top_p);
        B2_CHECK_ARG(sp->top_k >= 0, "b2_stream_begin: top_k must be >= 0");
        v.do_sample = 1; v.temperature = sp->temperature; v.top_p = sp->top_p; v.top_k = sp->top_k; v.seed = sp->seed;
    }
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    for (int b = 0; b < B; ++b)
        B2_CHECK_ARG(kv->len_host[b] >= 1, "b2_stream_begin: sample %d has an empty cache (prefill first)", b);
    kv->epoch += 1;
    v.tag = 1 + kv->epoch % 2047;
    B2_TRY(ws_enter(m, st));
    B2_TRY(set_sampling(kv, v, true, st));
    B2_CUDA_CHECK(cudaMemsetAsync(kv->step_counter.p, 0, 4, st));
    // token 0: chosen from the prefill's last-position logits, published as ring entry 0, fed to the first decode step
    B2_TRY(sample_publish(logits, m->d.vocab, B, kv->sstate.as<SampleState>(), kv->rows_dev.as<RowState>(), kv->tok.as<int32_t>(), nullptr,
                          kv->step_counter.as<int32_t>(), kv->len_dev.as<int32_t>(), kv->ring_dev, kv->ring_cap, SP_SELECT, 0, st));
    kv->stream_B = B; kv->stream_tag = v.tag; kv->stream_scheduled = 1;
    return ws_leave(m, st);
}

int b2_stream_enqueue(b2_model* m, b2_kv* kv, int n_steps, void* stream) {
    B2_CHECK_ARG(m && kv && kv->m == m && n_steps >= 1, "b2_stream_enqueue: bad argument");
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    B2_CHECK_ARG(kv->stream_B >= 1, "b2_stream_enqueue: no streaming generation on this cache (b2_stream_begin first)");
    const int B = kv->stream_B;
    const bool per_row = kv->samp_host.per_row != 0;
    // one generation never outgrows the ring (ring_cap = max_seq); a continuously batched cache runs indefinitely and wraps
    B2_CHECK_ARG(per_row || kv->stream_scheduled + n_steps <= kv->ring_cap, "b2_stream_enqueue: %d tokens exceed the ring capacity %d",
                 kv->stream_scheduled + n_steps, kv->ring_cap);
    for (int b = 0; b < B; ++b)
        B2_CHECK_ARG((per_row && !kv->rows_host[b].active) || kv->len_host[b] + n_steps <= kv->max_seq,
                     "b2_stream_enqueue: sample %d cache length %d + %d steps exceeds capacity %d", b, kv->len_host[b], n_steps, kv->max_seq);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    B2_TRY(ws_enter(m, st));
    cudaStream_t run = nullptr;
    B2_TRY(fork_stream(kv, st, &run));
    for (int s = 0; s < n_steps; ++s) B2_TRY(decode_step_run(m, kv, B, run));
    B2_TRY(join_stream(kv, st, run));
    for (int b = 0; b < B; ++b)
        if (!per_row || kv->rows_host[b].active) kv->len_host[b] += n_steps;
    kv->stream_scheduled += n_steps;
    return ws_leave(m, st);
}

// ---- continuous batching: requests join and leave the slots of one cache between decode steps ----------------------------
int b2_batch_begin(b2_model* m, b2_kv* kv, int B, void* stream) {
    B2_CHECK_ARG(m && kv && kv->m == m && B >= 1 && B <= kv->max_batch, "b2_batch_begin: bad argument");
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    SampleState v = {};
    v.temperature = 1.f; v.top_p = 1.f; v.per_row = 1;
    kv->epoch += 1;
    v.tag = 1 + kv->epoch % 2047;
    B2_TRY(ws_enter(m, st));
    B2_TRY(set_sampling(kv, v, true, st));
    B2_CUDA_CHECK(cudaMemsetAsync(kv->step_counter.p, 0, 4, st));
    B2_CUDA_CHECK(cudaMemsetAsync(kv->rows_dev.p, 0, (size_t)kv->max_batch * sizeof(RowState), st));
    B2_CUDA_CHECK(cudaMemsetAsync(kv->len_dev.p, 0, (size_t)kv->max_batch * 4, st));
    B2_CUDA_CHECK(cudaMemsetAsync(kv->tok.p, 0, (size_t)kv->max_batch * 4, st));
    kv->rows_host.assign(kv->max_batch, RowState{});
    kv->len_host.assign(kv->max_batch, 0);
    kv->stream_B = B; kv->stream_tag = v.tag; kv->stream_scheduled = 0;
    return ws_leave(m, st);
}

int b2_batch_set_row(b2_model* m, b2_kv* kv, int slot, int active, const b2_sampling* sp, int first_token, void* stream) {
    B2_CHECK_ARG(m && kv && kv->m == m && slot >= 0 && slot < kv->stream_B, "b2_batch_set_row: bad slot %d", slot);
    B2_CHECK_ARG(kv->samp_host.per_row != 0, "b2_batch_set_row: b2_batch_begin first");
    RowState v = {};
    v.active = active ? 1 : 0; v.temperature = 1.f; v.top_p = 1.f; v.index = 1;  // draw 0 chose first_token
    if (active && sp != nullptr && sp->do_sample) {
        B2_CHECK_ARG(sp->temperature > 0.f && sp->top_p > 0.f && sp->top_p <= 1.f && sp->top_k >= 0, "b2_batch_set_row: bad sampling parameters");
        v.do_sample = 1; v.temperature = sp->temperature; v.top_p = sp->top_p; v.top_k = sp->top_k; v.seed = sp->seed;
    }
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    B2_CHECK_ARG(!active || kv->len_host[slot] >= 1, "b2_batch_set_row: slot %d has an empty cache (prefill it first)", slot);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    B2_TRY(ws_enter(m, st));
    B2_TRY(row_state_set(kv->rows_dev.as<RowState>() + slot, v, active ? kv->tok.as<int32_t>() + slot : nullptr, first_token, st));
    if (!active) {  // a freed slot restarts from an empty cache
        B2_CUDA_CHECK(cudaMemsetAsync(kv->len_dev.as<int32_t>() + slot, 0, 4, st));
        kv->len_host[slot] = 0;
    }
    kv->rows_host[slot] = v;
    return ws_leave(m, st);
}

// Blocks (the Python binding releases the GIL around it) until token `index` of the current generation is in the ring.
// Takes no lock: other threads keep issuing work on the model while this one waits.
int b2_stream_wait(b2_kv* kv, int index, int32_t* tokens_host, int timeout_ms) {
    B2_CHECK_ARG(kv && tokens_host && index >= 0, "b2_stream_wait: bad argument");
    const int B = kv->stream_B;
    B2_CHECK_ARG(B >= 1, "b2_stream_wait: no streaming generation on this cache");
    B2_CHECK_ARG(index < kv->stream_scheduled, "b2_stream_wait: token %d has not been scheduled (%d scheduled)", index,
                 kv->stream_scheduled);
    volatile int32_t* slot = kv->ring_host + (size_t)(index % kv->ring_cap) * B;
    const int tag = 1 + (kv->stream_tag - 1 + index / kv->ring_cap) % 2047;  // the tag advances when the ring wraps (sampling.cu)
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (unsigned long long spins = 0;; ++spins) {
        bool ready = true;
        for (int b = 0; b < B; ++b)
            if ((slot[b] >> 20) != tag) { ready = false; break; }
        if (ready) {
            for (int b = 0; b < B; ++b) tokens_host[b] = slot[b] & 0xFFFFF;
            return 0;
        }
        if (spins < 4096) continue;                 // ~ tens of microseconds of pure polling, then back off
        struct timespec nap = {0, 20000};
        nanosleep(&nap, nullptr);
        if ((spins & 255) == 0) {
            clock_gettime(CLOCK_MONOTONIC, &t1);
            const double ms = (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6;
            cudaError_t e = cudaPeekAtLastError();  // a sticky fault (illegal address, trap) would never publish the token
            if (e != cudaSuccess) { set_error("b2_stream_wait: CUDA error while waiting for token %d: %s", index, cudaGetErrorString(e)); return -2; }
            if (timeout_ms > 0 && ms > timeout_ms) { set_error("b2_stream_wait: token %d not published after %d ms", index, timeout_ms); return -3; }
        }
    }
}

int b2_op_preprocess_clip(const b2_preprocess_plan* pl, void* stream) {
    B2_CHECK_ARG(pl != nullptr, "b2_op_preprocess_clip: null plan");
    PreprocessArgs a;
    a.img = pl->img; a.H = pl->H; a.W = pl->W; a.pad_top = pl->pad_top; a.pad_left = pl->pad_left;
    for (int i = 0; i < 3; ++i) { a.bg[i] = pl->bg[i]; a.mean[i] = pl->mean[i]; a.stdv[i] = pl->stdv[i]; }
    a.h_bounds = pl->h_bounds; a.h_kk = pl->h_kk; a.h_ksize = pl->h_ksize; a.h_identity = pl->h_identity;
    a.v_bounds = pl->v_bounds; a.v_kk = pl->v_kk; a.v_ksize = pl->v_ksize; a.v_identity = pl->v_identity;
    a.y0 = pl->y0; a.rows = pl->rows; a.x_lo = pl->x_lo; a.y_lo = pl->y_lo; a.cols = pl->out; a.out = pl->out;
    a.tmp = pl->tmp; a.rescale = pl->rescale; a.pixels = pl->pixels; a.u8_out = pl->u8_out;
    return preprocess_clip_image(a, reinterpret_cast<cudaStream_t>(stream));
}

// standalone selection (unit tests, first-token choice outside a streaming generation): out_tokens[b] (device int32)
int b2_op_sample(const float* logits, int B, int V, const b2_sampling* sp, int index, int32_t* out_tokens, void* stream) {
    B2_CHECK_ARG(logits && out_tokens && B >= 1 && V >= 1 && index >= 0, "b2_op_sample: bad argument");
    SampleState v = {};
    v.temperature = 1.f; v.top_p = 1.f;
    if (sp != nullptr && sp->do_sample) {
        B2_CHECK_ARG(sp->temperature > 0.f && sp->top_p > 0.f && sp->top_p <= 1.f && sp->top_k >= 0, "b2_op_sample: bad sampling parameters");
        v.do_sample = 1; v.temperature = sp->temperature; v.top_p = sp->top_p; v.top_k = sp->top_k; v.seed = sp->seed;
    }
    v.pub_counter = index;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    DevBuf tmp;
    B2_TRY(tmp.alloc(sizeof(SampleState) + 64));
    int r = sample_state_set(tmp.as<SampleState>(), v, st);
    if (r == 0)
        r = sample_publish(logits, V, B, tmp.as<SampleState>(), nullptr, out_tokens, nullptr, nullptr, nullptr, nullptr, 0, SP_SELECT, 0, st);
    cudaError_t e = cudaStreamSynchronize(st);
    tmp.free();
    if (r != 0) return r;
    B2_CUDA_CHECK(e);
    return 0;
}

int b2_argmax(const float* logits, int B, int V, int32_t* out, void* stream) {
    B2_CHECK_ARG(logits && out, "b2_argmax: null argument");
    return argmax_f32(logits, B, V, out, reinterpret_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------------------------------------
// single-kernel entry points
// ---------------------------------------------------------------------------------------------------------
int b2_op_gemm(const void* A, int lda, const void* W, int ldw, const void* bias, const void* residual, int ld_res,
               void* out, int ld_out, int out_fp32, int M, int N, int K, int act, int bn_override, void* stream) {
    B2_CHECK_ARG(A && W && out, "b2_op_gemm: null argument");
    GemmArgs g;
    g.A = A; g.lda = lda; g.W = W; g.ldw = ldw; g.bias = bias; g.residual = residual; g.ld_res = ld_res;
    g.out = out; g.ld_out = ld_out; g.out_fp32 = out_fp32; g.M = M; g.N = N; g.K = K; g.act = act;
    g.bn_override = bn_override;
    return gemm_bf16(g, reinterpret_cast<cudaStream_t>(stream));
}

int b2_op_gemv(const void* x, int64_t ldx, const void* W, int ldw, const void* norm_gamma, float eps,
               const void* residual, int ld_res, void* out, int ld_out, int out_fp32, int B, int N, int K, int act,
               void* stream) {
    B2_CHECK_ARG(x && W && out, "b2_op_gemv: null argument");
    return gemv(x, ldx, W, ldw, norm_gamma, eps, residual, ld_res, out, ld_out, out_fp32, B, N, K, act,
                reinterpret_cast<cudaStream_t>(stream));
}

int b2_op_gemm_skinny(const void* x, int ldx, const void* W, int ldw, const void* residual, int ld_res, void* out,
                      int ld_out, int out_fp32, int B, int N, int K, int act, void* workspace, int64_t workspace_bytes,
                      void* counters, void* stream) {
    B2_CHECK_ARG(x && W && out && workspace && counters, "b2_op_gemm_skinny: null argument");
    SkinnyArgs g;
    g.x = x; g.ldx = ldx; g.W = W; g.ldw = ldw; g.residual = residual; g.ld_res = ld_res;
    g.out = out; g.ld_out = ld_out; g.out_fp32 = out_fp32; g.B = B; g.N = N; g.K = K; g.act = act;
    g.partial = reinterpret_cast<float*>(workspace); g.partial_bytes = (size_t)workspace_bytes;
    g.counters = reinterpret_cast<int*>(counters);
    return gemm_skinny_bf16(g, reinterpret_cast<cudaStream_t>(stream));
}
int b2_op_gemm_skinny_fp8(const void* xq, int ldx, const float* x_scale, const void* Wq, int ldw, const float* w_scale,
                          const void* residual, int ld_res, void* out, int ld_out, int out_fp32, int B, int N, int K, int act,
                          void* workspace, int64_t workspace_bytes, void* counters, void* stream) {
    B2_CHECK_ARG(xq && Wq && x_scale && w_scale && out && workspace && counters, "b2_op_gemm_skinny_fp8: null argument");
    SkinnyArgs g;
    g.x = xq; g.ldx = ldx; g.W = Wq; g.ldw = ldw; g.residual = residual; g.ld_res = ld_res;
    g.out = out; g.ld_out = ld_out; g.out_fp32 = out_fp32; g.B = B; g.N = N; g.K = K; g.act = act;
    g.partial = reinterpret_cast<float*>(workspace); g.partial_bytes = (size_t)workspace_bytes;
    g.counters = reinterpret_cast<int*>(counters);
    g.w_scale = w_scale; g.x_scale = x_scale;
    return gemm_skinny_fp8(g, reinterpret_cast<cudaStream_t>(stream));
}
int b2_op_quantize_rows_e4m3(const void* x, int64_t ldx, int rows, int K, void* q, int64_t ldq, float* scale, void* stream) {
    B2_CHECK_ARG(x && q && scale, "b2_op_quantize_rows_e4m3: null argument");
    return quantize_rows_e4m3(x, ldx, rows, K, q, ldq, scale, reinterpret_cast<cudaStream_t>(stream));
}
int b2_op_rmsnorm_quant_e4m3(const void* x, const void* gamma, void* q, float* scale, int rows, int cols, float eps,
                             void* stream) {
    B2_CHECK_ARG(x && gamma && q && scale, "b2_op_rmsnorm_quant_e4m3: null argument");
    return rmsnorm_quant_e4m3(x, cols, gamma, q, cols, scale, rows, cols, eps, reinterpret_cast<cudaStream_t>(stream));
}
int64_t b2_op_gemm_skinny_workspace_bytes(int B, int N, int K) {
    if (B < 1 || B > 128 || N < 1 || K < 1) return -1;
    return (int64_t)gemm_skinny_workspace_bytes(B, N, K);
}
int64_t b2_op_gemm_skinny_counter_bytes(int N) { return N < 1 ? -1 : (int64_t)gemm_skinny_counter_bytes(N); }

int b2_op_layernorm(const void* x, const void* gamma, const void* beta, void* y, int rows, int cols, float eps,
                    void* stream) {
    B2_CHECK_ARG(x && gamma && beta && y, "b2_op_layernorm: null argument");
    return layernorm_bf16(x, gamma, beta, y, rows, cols, eps, reinterpret_cast<cudaStream_t>(stream));
}

int b2_op_rmsnorm(const void* x, const void* gamma, void* y, int rows, int cols, float eps, void* stream) {
    B2_CHECK_ARG(x && gamma && y, "b2_op_rmsnorm: null argument");
    return rmsnorm_bf16(x, cols, gamma, y, rows, cols, eps, reinterpret_cast<cudaStream_t>(stream));
}

int b2_op_flash_attn(const void* q, const void* k, const void* v, void* o, const int32_t* seq_lens, int B, int S,
                     int H, int D, int causal, float scale, void* stream) {
    B2_CHECK_ARG(q && k && v && o, "b2_op_flash_attn: null argument");
    FlashArgs fa;
    const int64_t ts = (int64_t)H * D, bs = (int64_t)S * H * D;
    fa.q = q; fa.q_bs = bs; fa.q_ts = ts; fa.q_hs = D;
    fa.k = k; fa.k_bs = bs; fa.k_ts = ts; fa.k_hs = D;
    fa.v = v; fa.v_bs = bs; fa.v_ts = ts; fa.v_hs = D;
    fa.o = o; fa.o_bs = bs; fa.o_ts = ts; fa.o_hs = D;
    fa.seq_lens = seq_lens; fa.B = B; fa.H = H; fa.S = S; fa.D = D; fa.causal = causal; fa.scale = scale;
    return flash_attn_bf16(fa, reinterpret_cast<cudaStream_t>(stream));
}

int b2_op_rope_kv_write(void* qkv, void* kcache, void* vcache, int B, int S, int H, int D, int Smax, float theta,
                        void* stream) {
    B2_CHECK_ARG(qkv && kcache && vcache, "b2_op_rope_kv_write: null argument");
    return rope_kv_write(qkv, kcache, vcache, B, S, H, D, Smax, theta, reinterpret_cast<cudaStream_t>(stream));
}

int64_t b2_op_decode_attn_scratch_bytes(int B, int H, int nsplit) {
    return (int64_t)B * H * 4 /*counters*/ + 256 + (int64_t)B * H * nsplit * (128 + 2) * 4;
}

int b2_op_decode_attn(const void* qkv, void* kcache, void* vcache, const int32_t* cur_len, void* out, void* scratch,
                      int B, int H, int Smax, int nsplit, float theta, float scale, void* stream) {
    B2_CHECK_ARG(qkv && kcache && vcache && cur_len && out && scratch, "b2_op_decode_attn: null argument");
    DecodeAttnArgs da;
    da.qkv = qkv; da.kcache = kcache; da.vcache = vcache; da.cur_len = cur_len; da.out = out;
    da.counters = reinterpret_cast<int32_t*>(scratch);
    const size_t off = ((size_t)B * H * 4 + 255) / 256 * 256;
    da.partial = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(scratch) + off);
    da.B = B; da.H = H; da.D = 128; da.Smax = Smax; da.nsplit = nsplit; da.theta = theta; da.scale = scale;
    return decode_attn_bf16(da, reinterpret_cast<cudaStream_t>(stream));
}

int b2_op_interleave_gate_up(const void* gate, const void* up, void* out, int I, int h, void* stream) {
    B2_CHECK_ARG(gate && up && out, "b2_op_interleave_gate_up: null argument");
    return interleave_gate_up(gate, up, out, I, h, reinterpret_cast<cudaStream_t>(stream));
}

int b2_op_im2col(const void* pixels, void* out, int B, int img, int patch, int kpad, void* stream) {
    B2_CHECK_ARG(pixels && out, "b2_op_im2col: null argument");
    return vit_im2col(pixels, out, B, img, patch, kpad, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
