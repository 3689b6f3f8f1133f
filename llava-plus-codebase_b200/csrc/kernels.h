// Internal launcher declarations for the b2llava kernels (C++ side, not part of the C ABI).
// All tensors are device pointers; activations/weights are bf16 unless stated otherwise.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2 {

enum { ACT_NONE = 0, ACT_QUICK_GELU = 1, ACT_GELU_ERF = 2, ACT_SWIGLU = 3, ACT_ROPE_QKV = 4 };

// ACT_ROPE_QKV (CTA-pair kernel only): the GEMM is LLaMA's fused QKV projection of a prefill, rows = b*S + t. The epilogue
// rounds the projection to bf16, applies RoPE (HF rounding points, cos/sin from `table`) to the q and k heads, writes q to
// `out` and k / v straight into the KV cache [b][head][Smax][128] — the standalone rope_kv_write pass over qkv disappears.
struct RopeQkv {
    const void* table = nullptr;   // uint32 [Smax][64]: bf16 cos | bf16 sin << 16 of position t, frequency i (rope_table_build)
    void* kcache = nullptr;        // this layer's K slab of the first batch row written
    void* vcache = nullptr;
    int S = 0, H = 0, Smax = 0;
};
enum { DT_BF16 = 0, DT_F16 = 1, DT_F32 = 2 };

int num_sms();

// ---- tcgen05 GEMM (gemm_tcgen05.cu) -------------------------------------------------------------
// out[M, N] = act(A[M,K] · W[N,K]^T + bias) + residual      (ACT_NONE / QUICK_GELU / GELU_ERF)
// out[M, N/2] = silu(gate) * up                               (ACT_SWIGLU; W rows block-64 interleaved)
struct GemmArgs {
    const void* A = nullptr; int lda = 0;        // bf16 [M, K]
    const void* W = nullptr; int ldw = 0;        // bf16 [N, K]
    const void* bias = nullptr;                  // bf16 [N] or null
    const void* residual = nullptr; int ld_res = 0;  // bf16 [M, N] or null (may alias out)
    void* out = nullptr; int ld_out = 0; int out_fp32 = 0;
    int M = 0, N = 0, K = 0;
    int act = ACT_NONE;
    int bn_override = 0;  // 0 = cost-model heuristic, else 64/128/192/256 (2 = CTA-pair kernel)
    RopeQkv rope;         // ACT_ROPE_QKV only
};
int gemm_bf16(const GemmArgs& g, cudaStream_t stream);
// mm_projector mlp2x_gelu as ONE kernel (gemm_tcgen05.cu, projector_fused_kernel): out = gelu_erf(X W1^T + b1) W2^T + b2.
// H [M,N1] bf16 and row_done int[ceil(M/128)] are caller scratch.
int projector_fused_bf16(const void* X, int ldx, const void* W1, const void* b1, const void* W2, const void* b2, void* H,
                         void* out, int ld_out, int M, int K1, int N1, int N2, int* row_done, cudaStream_t stream);
// CTA-pair (cta_group::2, 256x256 pair tiles) variant, gemm_2cta.cu (chosen by gemm_bf16's cost model, or bn_override == 2)
int gemm_bf16_2cta(const GemmArgs& g, cudaStream_t stream);

// ---- swap-AB stream-K GEMM for decode at batch 9..128 (gemm_skinny.cu) ------------------------------------
// out[B, N] = x[B,K] · W[N,K]^T (+ residual); ACT_SWIGLU: out[B, N/2] (W rows block-64 interleaved).
struct SkinnyArgs {
    const void* x = nullptr; int ldx = 0;        // bf16 [B, K]
    const void* W = nullptr; int ldw = 0;        // bf16 [N, K]
    const void* residual = nullptr; int ld_res = 0;  // bf16 [B, N] or null (may alias out)
    void* out = nullptr; int ld_out = 0; int out_fp32 = 0;
    int B = 0, N = 0, K = 0;
    int act = ACT_NONE;                          // ACT_NONE | ACT_SWIGLU
    float* partial = nullptr; size_t partial_bytes = 0;  // >= gemm_skinny_workspace_bytes(B, N, K)
    int* counters = nullptr;                     // >= gemm_skinny_counter_bytes(N), zero-initialised once
    const float* w_scale = nullptr;              // fp8 variant: [N] per weight row; x / W then hold e4m3 bytes
    const float* x_scale = nullptr;              // fp8 variant: [B] per token
};
int gemm_skinny_bf16(const SkinnyArgs& g, cudaStream_t stream);
int gemm_skinny_fp8(const SkinnyArgs& g, cudaStream_t stream);
size_t gemm_skinny_workspace_bytes(int B, int N, int K);
size_t gemm_skinny_counter_bytes(int N);

// ---- e4m3 row quantisation for the fp8 decode path (quant_fp8.cu) ----------------------
// scale[r] = amax_r / 448 (1 for a zero row); q[r,k] = e4m3_rn_satfinite(x[r,k] * (448 / amax_r)); ld* in elements
int quantize_rows_e4m3(const void* x, int64_t ldx, int rows, int K, void* q, int64_t ldq, float* scale, cudaStream_t stream);
// RMSNorm (HF rounding points) fused with the per-token quantisation of its output
int rmsnorm_quant_e4m3(const void* x, int64_t x_row_stride, const void* gamma, void* q, int64_t ldq, float* scale, int rows,
                       int cols, float eps, cudaStream_t stream);

// ---- weight-streaming GEMV for decode (gemv.cu) ---------------------------------------------------
// out[B, N] = (rmsnorm(x) or x)[B,K] · W[N,K]^T (+ residual); B <= 8. ACT_SWIGLU: out[B, N/2].
struct GemvArgs {
    const void* x = nullptr; int64_t ldx = 0;   // bf16 [B, K], row stride ldx (elements)
    const void* W = nullptr; int ldw = 0;        // bf16 [N, K]
    const void* norm_gamma = nullptr; float eps = 0.f;  // fused RMSNorm on x when non-null
    const void* residual = nullptr; int ld_res = 0;     // bf16 [B, N] or null (may alias out)
    void* out = nullptr; int ld_out = 0; int out_fp32 = 0;
    int B = 0, N = 0, K = 0;
    int act = ACT_NONE;
};
int gemv_bf16(const GemvArgs& g, cudaStream_t stream);
bool gemv_fits(int B, int N, int K, int act);  // activation tile + partial table fit in shared memory

// ---- norms (norms.cu) ------------------------------------------------------------------------------
int layernorm_bf16(const void* x, const void* gamma, const void* beta, void* y, int rows, int cols, float eps,
                   cudaStream_t stream);
// HF LlamaRMSNorm semantics: y = gamma * bf16(x * rsqrt(mean(x^2) + eps)); x rows may be strided.
int rmsnorm_bf16(const void* x, int64_t x_row_stride, const void* gamma, void* y, int rows, int cols, float eps,
                 cudaStream_t stream);
// gather variant: row r of y is the RMSNorm of x[row_index[r]] (row_index on device)
int rmsnorm_gather_bf16(const void* x, const int32_t* row_index, const void* gamma, void* y, int rows, int cols,
                        float eps, cudaStream_t stream);

// ---- ViT helpers (vit_ops.cu) ----------------------------------------------------------------------
// pixels [B,3,img,img] bf16 -> patches [B*P, kpad] (k = c*patch*patch + i*patch + j, zero padded to kpad)
int vit_im2col(const void* pixels, void* out, int B, int img, int patch, int kpad, cudaStream_t stream);
// hidden[b, 0] = LN(cls + pos[0]); hidden[b, 1+p] = LN(patch_out[b*P+p] + pos[1+p])   (pre_layrnorm fused)
int vit_embed_ln(const void* patch_out, const void* cls, const void* pos, const void* gamma, const void* beta,
                 void* hidden, int B, int P, int D, float eps, cudaStream_t stream);
// out[b, p, :] = hidden[b, 1 + p, :]   (feature_select 'patch': drop CLS)
int vit_drop_cls(const void* hidden, void* out, int B, int P, int D, cudaStream_t stream);

// ---- attention (attention.cu) ----------------------------------------------------------------------
struct FlashArgs {
    const void* q = nullptr; int64_t q_bs = 0, q_ts = 0, q_hs = 0;  // element strides: batch, token, head
    const void* k = nullptr; int64_t k_bs = 0, k_ts = 0, k_hs = 0;
    const void* v = nullptr; int64_t v_bs = 0, v_ts = 0, v_hs = 0;
    void* o = nullptr;       int64_t o_bs = 0, o_ts = 0, o_hs = 0;
    const int32_t* seq_lens = nullptr;  // device [B] or null (=S)
    int B = 0, H = 0, S = 0, D = 0;     // S = padded/query length; D in {64, 128}
    int causal = 0;
    float scale = 1.f;
};
int flash_attn_bf16(const FlashArgs& a, cudaStream_t stream);      // dispatcher (tcgen05 unless B2_FLASH_TC=0)
int flash_attn_tc_bf16(const FlashArgs& a, cudaStream_t stream);   // attention_tc.cu: tcgen05 + TMEM + TMA
int flash_attn_mma_bf16(const FlashArgs& a, cudaStream_t stream);  // attention.cu: mma.sync variant

// prefill: RoPE on q (in place) and k inside qkv [B*S, 3*H*D]; roped k and v written to the cache
// kcache/vcache: [Bmax, H, Smax, D] for one layer. Positions are 0..S-1 (right-padded rows).
// (cos, sin) table of rope_kv_write's positions 0..Smax-1 (head_dim D): uint32 [Smax][D/2], bf16 cos | bf16 sin << 16
int rope_table_build(void* table, int Smax, int D, float theta, cudaStream_t stream);
int rope_kv_write(void* qkv, void* kcache, void* vcache, int B, int S, int H, int D, int Smax, float theta,
                  cudaStream_t stream);

// decode: q/k/v of the new token from qkv [B, 3*H*D]; RoPE at position cur_len[b]; append k,v to the cache;
// split-KV attention over cur_len[b]+1 keys; out [B, H*D] bf16.
struct DecodeAttnArgs {
    const void* qkv = nullptr;
    void* kcache = nullptr; void* vcache = nullptr;  // layer base, [Bmax, H, Smax, D]
    const int32_t* cur_len = nullptr;                // device [B]
    void* out = nullptr;
    float* partial = nullptr;                        // workspace [B*H*nsplit*(D+2)] fp32
    int32_t* counters = nullptr;                     // workspace [B*H], zero-initialised, self-resetting
    int B = 0, H = 0, D = 128, Smax = 0, nsplit = 1;
    float theta = 10000.f, scale = 1.f;
};
int decode_attn_ctas_per_sm();
int decode_attn_bf16(const DecodeAttnArgs& a, cudaStream_t stream);

// ---- persistent decode-step megakernel (decode_mega.cu), batch <= 8 ------------------------------------
struct MegaLayer {
    const __nv_bfloat16 *ln1, *wqkv, *wo, *ln2, *wgu, *wd;
    __nv_bfloat16 *kcache, *vcache;  // this layer's [Bmax, H, Smax, 128] slabs
};
struct MegaParams {
    const MegaLayer* layers = nullptr;  // device array [L]
    int L = 0, h = 0, I = 0, H = 0, V = 0, B = 0, Smax = 0, nsplit = 1;
    const __nv_bfloat16 *embed = nullptr, *final_norm = nullptr, *lm_head = nullptr;
    int32_t *tok = nullptr, *cur_len = nullptr, *out_tokens = nullptr, *step_counter = nullptr;
    __nv_bfloat16 *x = nullptr, *qkv = nullptr, *attn = nullptr, *act = nullptr;
    float* logits = nullptr;
    float* attn_partial = nullptr;      // [B*H*nsplit*(128+2)]
    int32_t* attn_counters = nullptr;   // [B*H], zero-initialised, self-resetting
    unsigned int *bar_count = nullptr, *done_count = nullptr;  // zero-initialised
    unsigned int bar_base = 0;  // barrier-counter value before this launch = launches so far * (5L+2) * grid
    float eps = 1e-5f, theta = 10000.f, scale_log2 = 1.f;
    int l2_ahead = 0;       // weight tiles pulled into L2 in front of the shared-memory ring (0 = off), multiple of 4
    int l2_mode = 1;        // 1 = prefetch.global.L2 lines (LSU), 2 = cp.async.bulk.prefetch.L2 (TMA queue)
    int fast_prologue = 0;  // single-pass activation staging with pre-barrier RMSNorm-weight loads
    int gamma_smem = 0;     // next phase's RMSNorm weights staged in shared memory (cp.async) in front of the grid barrier
    long long* trace = nullptr;  // optional [n_phases+2][4] SM-clock timestamps of CTA 0 (B2_MEGA_TRACE=1)
    // token publication to the host ring (sampling.cu): non-null only for greedy streaming; with do_sample the separate
    // sample_publish kernel that follows the launch overrides the fused argmax and publishes instead
    struct SampleState* sstate = nullptr;
    int32_t* ring = nullptr;
    int ring_cap = 0;
    const struct RowState* rows = nullptr;  // continuous batching: only active slots advance their cache length
};
// one launch = embed -> all layers -> lm_head -> argmax -> token store; cur_len/step_counter advance on device
int decode_mega(const MegaParams& p, cudaStream_t stream);
bool decode_mega_fits(int B, int h, int I);

// ---- token selection + host-ring publication (sampling.cu) --------------------------------------------------
// Device-resident state of one generation (lives in the b2_kv): read by the kernels, so a captured CUDA graph of the decode
// step stays valid when the sampling parameters change.
struct SampleState {
    int do_sample;            // 0 = greedy argmax
    float temperature, top_p;
    int top_k;                // 0 = off
    unsigned long long seed;  // Philox key
    int tag;                  // generation epoch 1..2047 published with every token; 0 = do not publish to the host ring
    int pub_counter;          // index of the next token of this generation
    unsigned int done;        // rows finished in the current launch (self-resetting)
    int per_row;              // continuous batching: selection parameters and liveness come from RowState[b] instead
};
// One cache slot of a continuously batched decode (llava/_b2/batching.py): requests join and leave between steps, each with
// its own sampling parameters and its own Philox draw index; a slot that is not active keeps its cache length and token.
struct RowState {
    int active;
    int do_sample; float temperature, top_p; int top_k;
    unsigned long long seed;
    int index;                // draws made for this request so far
};
enum { SP_SELECT = 1,     // choose from `logits` (argmax or sample) and write tok[b]; otherwise tok[b] is already chosen
       SP_WRITE_OUT = 2,  // out_tokens[(*step_counter + step_offset) * B + b] = token
       SP_BUMP = 4 };     // last row: *step_counter += 1, cur_len[b] += 1
int sample_state_set(SampleState* st_dev, const SampleState& v, cudaStream_t stream);
int sample_publish(const float* logits, int V, int B, SampleState* st_dev, RowState* rows_dev, int32_t* tok, int32_t* out_tokens,
                   int32_t* step_counter, int32_t* cur_len, int32_t* ring_dev, int ring_cap, int flags, int step_offset,
                   cudaStream_t stream);
int row_state_set(RowState* row_dev, const RowState& v, int32_t* tok_dev, int token, cudaStream_t stream);

// ---- image preprocessing (preprocess.cu): uint8 HWC -> CLIP pixel_values, PIL-exact bicubic resize ------------------------
struct PreprocessArgs {
    const uint8_t* img = nullptr; int H = 0, W = 0;        // device, RGB HWC
    int pad_top = 0, pad_left = 0; uint8_t bg[3] = {0, 0, 0};  // virtual expand2square: reads outside the image return bg
    const int32_t *h_bounds = nullptr, *h_kk = nullptr; int h_ksize = 0, h_identity = 0;  // tables over the resized WIDTH
    const int32_t *v_bounds = nullptr, *v_kk = nullptr; int v_ksize = 0, v_identity = 0;  // tables over the resized HEIGHT
    int y0 = 0, rows = 0;        // source rows the vertical pass reads: [y0, y0 + rows)
    int x_lo = 0, y_lo = 0;      // centre-crop origin inside the resized image
    int cols = 0, out = 0;       // cols == out: crop width / output size
    uint8_t* tmp = nullptr;      // [rows, cols, 3] scratch
    float mean[3] = {0, 0, 0}, stdv[3] = {1, 1, 1}, rescale = 1.f / 255.f;
    void* pixels = nullptr;      // bf16 [3, out, out] or null
    uint8_t* u8_out = nullptr;   // uint8 [out, out, 3] (the resized + cropped image before normalisation) or null
};
int preprocess_clip_image(const PreprocessArgs& a, cudaStream_t stream);

// ---- misc (misc_ops.cu) ----------------------------------------------------------------------------
// out[r, :] = src_index[r] >= 0 ? table[src_index[r]] : (src_index[r] == INT32_MIN ? 0 : feats[-src_index[r]-1]).
// Rows whose index is outside [0, vocab) / [0, n_feat_rows) are written as zeros and reported through *err_flag
// (mapped host memory, B2_ERR_* codes, may be null): nothing is ever read out of bounds.
int splice_embed(const int32_t* src_index, const void* table, const void* feats, void* out, int rows, int h, int vocab,
                 int n_feat_rows, int* err_flag, cudaStream_t stream);
// device-built source index for equal-length unpadded rows with k_per_row image placeholders each (misc_ops.cu):
// ids int64 [B, Lt] on the device; feat_offsets_host[n_img + 1] = prefix sums of the feature rows of the image slots
int splice_index(const long long* ids, int B, int Lt, int k_per_row, const int32_t* feat_offsets_host, int n_img, int image_token,
                 int S, int32_t* src_index, int* err_flag, cudaStream_t stream);
int embed_tokens(const int32_t* tokens, const void* table, void* out, int rows, int h, int vocab, int* err_flag,
                 cudaStream_t stream);
enum { B2_ERR_TOKEN_RANGE = 1, B2_ERR_IMAGE_ROW_RANGE = 2, B2_ERR_SPLICE_SLOTS = 4 };
int argmax_f32(const float* logits, int B, int V, int32_t* out, cudaStream_t stream);
struct I32Pack { int32_t v[128]; };
// dst_a[i] = a_host[i] (and dst_b[i] = b_host[i] when dst_b != null), i < n: values travel as kernel parameters
int set_i32_pairs(int32_t* dst_a, const int32_t* a_host, int32_t* dst_b, const int32_t* b_host, int n, cudaStream_t stream);
int add_i32(int32_t* x, int n, int delta, cudaStream_t stream);
int convert_to_bf16(const void* src, int src_dtype, void* dst, int64_t n, cudaStream_t stream);
// out[2I, h]: within each 128-row group g: rows [0,64) = gate[g*64 .. +64), rows [64,128) = up[g*64 .. +64)
int interleave_gate_up(const void* gate, const void* up, void* out, int I, int h, cudaStream_t stream);
int store_token(const int32_t* src, int32_t* dst_base, const int32_t* step_counter, int B, cudaStream_t stream);

}  // namespace b2
