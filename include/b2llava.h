/* b2llava.h — C ABI of the B200-native LLaVA multimodal forward path (libb2llava.so).
 *
 * The reference (LLaVA-VL/LLaVA-Plus-Codebase) has NO native boundary on this path: its hot path is a Python
 * class surface (llava/model/language_model/llava_llama.py:56-108, llava/model/llava_arch.py:94-240,
 * llava/model/multimodal_encoder/clip_encoder.py:39-51, llava/model/multimodal_projector/builder.py:33-51)
 * sitting directly on HuggingFace transformers + ATen. This header is the boundary we introduce where
 * HF/ATen sit today: each entry point names the reference function whose arithmetic it
 * replaces. The Python package `llava` in this repo binds it with ctypes.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types. cudaStream_t is passed as void*.
 *   - All tensor pointers are DEVICE pointers unless the parameter name ends in _host or the doc says
 *     "host or device". Activations and weights are bf16 (uint16 storage); logits are fp32.
 *   - Return 0 on success, <0 on error (-1 bad argument, -2 CUDA failure, -3 bad state); the message is
 *     available from b2_last_error() (thread-local). Nothing aborts the process; no exceptions cross the ABI.
 *   - The caller owns every input/output buffer and the stream. The library owns weights (copied and
 *     repacked at set_weight/finalize), workspaces and KV caches.
 *   - A b2_model / b2_kv handle may be used from any host thread, one call at a time (calls lock the handle).
 */
#ifndef B2LLAVA_H_
#define B2LLAVA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b2_model b2_model;
typedef struct b2_kv b2_kv;

/* dtype codes for b2_model_set_weight */
enum { B2_DT_BF16 = 0, B2_DT_F16 = 1, B2_DT_F32 = 2 };
/* activation codes for b2_op_gemm */
enum { B2_ACT_NONE = 0, B2_ACT_QUICK_GELU = 1, B2_ACT_GELU_ERF = 2, B2_ACT_SWIGLU = 3 };
/* logits modes for b2_prefill */
enum { B2_LOGITS_NONE = 0, B2_LOGITS_LAST = 1, B2_LOGITS_ALL = 2 };

typedef struct b2_model_desc {
    /* CLIP vision tower (transformers CLIPVisionConfig; reference clip_encoder.py:22-27) */
    int32_t image_size;       /* 336 */
    int32_t patch_size;       /* 14 */
    int32_t vit_hidden;       /* 1024 */
    int32_t vit_inter;        /* 4096 */
    int32_t vit_layers;       /* 24 (total in the checkpoint) */
    int32_t vit_heads;        /* 16 (head_dim must be 64) */
    int32_t vit_select_layer; /* -2: hidden_states index, reference clip_encoder.py:30 */
    float vit_ln_eps;         /* 1e-5 */
    /* LLaMA decoder (LlavaConfig(LlamaConfig), reference llava_llama.py:29-30) */
    int32_t hidden;           /* 4096 | 5120 */
    int32_t inter;            /* 11008 | 13824 */
    int32_t layers;           /* 32 | 40 */
    int32_t heads;            /* 32 | 40 (head_dim must be 128; kv_heads == heads) */
    int32_t vocab;            /* 32000 */
    float rms_eps;            /* 1e-5 */
    float rope_theta;         /* 10000 */
    /* workspace sizing */
    int32_t max_batch;        /* largest B for prefill/decode */
    int32_t max_seq;          /* largest spliced prefill length S (B*S rows of workspace) */
    int32_t max_images;       /* images encoded per ViT pass (larger batches are chunked) */
} b2_model_desc;

/* Token selection of the decode loop (HF GenerationMixin as the reference calls it: llava/serve/model_worker.py:155-185
 * passes do_sample / temperature / top_p; top_k comes from the GenerationConfig default). do_sample == 0 -> greedy argmax. */
typedef struct b2_sampling {
    int32_t do_sample;
    float temperature;        /* > 0 */
    float top_p;              /* (0, 1]; 1 = off */
    int32_t top_k;            /* 0 = off */
    unsigned long long seed;  /* Philox key; draw t of row b is a pure function of (logits, seed, t, b) */
} b2_sampling;

/* ---- lifecycle ------------------------------------------------------------------------------------------ */
int b2_init(int device);                 /* cudaSetDevice + capability check (needs sm_100) */
const char* b2_last_error(void);         /* thread-local message of the last failing call */
int b2_version(void);
unsigned long long b2_launch_count(void);/* number of kernels this library has launched (bench.py gpu_launches) */

int b2_model_create(const b2_model_desc* desc, b2_model** out);
/* Copy one checkpoint tensor into the model. `hf_key` uses the reference's state-dict names:
 * model.embed_tokens.weight, model.layers.{i}.self_attn.{q,k,v,o}_proj.weight,
 * model.layers.{i}.mlp.{gate,up,down}_proj.weight, model.layers.{i}.{input,post_attention}_layernorm.weight,
 * model.norm.weight, lm_head.weight, model.mm_projector.{0,2}.{weight,bias},
 * [model.vision_tower.vision_tower.]vision_model.* (CLIPVisionModel names). `ptr` may be host or device memory,
 * borrowed for the duration of the call. q/k/v are fused into one [3h,h] matrix, gate/up are block-64
 * interleaved for the fused SwiGLU epilogue. Unknown keys return -1; keys that are dead on this path
 * (vision post_layernorm, CLIP layers above the selected one, rotary inv_freq buffers) are accepted and ignored. */
int b2_model_set_weight(b2_model* m, const char* hf_key, const void* ptr, const int64_t* shape, int ndim, int dtype);
int b2_model_finalize(b2_model* m);      /* checks every tensor arrived, allocates workspaces */
int b2_model_destroy(b2_model* m);
/* BASELINE configs[4] ("fp8-weight tcgen05 path"): after finalize, quantise the decoder's Linear weights to e4m3 with one
 * fp32 scale per output channel; decode steps at batch >= 7 then run e4m3 x e4m3 tcgen05 GEMMs (activations quantised
 * per token on the fly, KV cache stays bf16). Prefill and small-batch decode keep the bf16 weights. The reference has no
 * fp8 path; oracle/fp8_oracle.py defines the arithmetic and the tolerance (tests/test_fp8_gpu.py). Off unless this call is
 * made: it changes the numerics of the decode step (W8A8), so it is never a default. */
int b2_model_enable_fp8_decode(b2_model* m);

int b2_kv_create(b2_model* m, int max_batch, int max_seq, b2_kv** out); /* KV cache [L][2][B][H][Smax][128] bf16 */
int b2_kv_reset(b2_kv* kv);
int b2_kv_destroy(b2_kv* kv);
int b2_kv_lengths(b2_kv* kv, int32_t* lens_host, int n);  /* current cache length per sample */

/* ---- hot path ------------------------------------------------------------------------------------------- */
/* CLIPVisionTower.forward + feature_select (reference clip_encoder.py:29-51; HF modeling_clip.py:202-217,
 * 300-384, 667-691): pixels [B,3,img,img] bf16 -> patch features [B, P, vit_hidden] bf16 of
 * hidden_states[vit_select_layer] with the CLS token dropped. Layers above the selected one are not computed. */
int b2_vit_encode(b2_model* m, const void* pixels, int B, void* out_feats, void* stream);
/* mm_projector mlp2x_gelu (reference multimodal_projector/builder.py:39-46): [rows, vit_hidden] -> [rows, hidden] */
int b2_project(b2_model* m, const void* feats, int rows, void* out, void* stream);
/* LlavaMetaForCausalLM.encode_images (reference llava_arch.py:94-97) = b2_vit_encode then b2_project:
 * pixels [B,3,img,img] -> [B, P, hidden] */
int b2_encode_images(b2_model* m, const void* pixels, int B, void* out, void* stream);
/* The device half of prepare_inputs_labels_for_multimodal (reference llava_arch.py:150-225). src_index[r]
 * (device int32, one per output row of the padded [B,S] layout) is: >= 0 -> embed_tokens row (token id);
 * < 0 and != INT32_MIN -> row (-src-1) of image_feats [n_img*P, hidden]; INT32_MIN -> zero row (padding).
 * The index is built on the host from input_ids (llava/model/llava_arch.py in this repo). image_feats holds n_feat_rows
 * rows (0 with a NULL pointer for text-only batches). An index outside the embedding table or the feature rows never reads
 * out of bounds: the row is zero-filled and the problem is reported by b2_async_error(). */
int b2_splice(b2_model* m, const int32_t* src_index, const void* image_feats, int n_feat_rows, int rows, void* embeds_out,
              void* stream);
/* The whole splice on the device, for the layout every generation caller of the reference uses (equal-length rows without
 * padding, k_per_row IMAGE_TOKEN_INDEX placeholders per row; llava/serve/model_worker.py:163, llava/eval/model_vqa_loader.py:98):
 * input_ids int64 [B,Lt] stays on the device (no D2H of the ids, no host loop over rows — reference llava_arch.py:143-187);
 * image slot j holds feature rows [feat_offsets_host[j], feat_offsets_host[j+1]) of image_feats, slots are consumed in
 * row-major order. The caller derives S = Lt - k + rows-per-row from shapes alone. A row whose placeholder count is not
 * k_per_row is flagged (b2_async_error code 4) and the caller redoes the splice on the exact host path (b2_splice). */
int b2_splice_ids(b2_model* m, const int64_t* input_ids, int B, int Lt, int k_per_row, const int32_t* feat_offsets_host, int n_img,
                  const void* image_feats, int S, void* embeds_out, void* stream);
/* Input problems that only a kernel can see (ids live on the device): returns in *code_out the OR of 1 = token id outside
 * [0, vocab) or an image placeholder without features, 2 = image-feature row out of range, 4 = more placeholders than
 * images, accumulated since the last call, and clears it. Meaningful after the stream has been synchronised (the codes
 * are written to mapped host memory by the kernels); b2_last_error() then holds the text. */
int b2_async_error(b2_model* m, int* code_out);
/* LlamaModel.forward prefill over inputs_embeds [B,S,hidden] (HF modeling_llama.py:375-425 and :303-332 per
 * layer), right-padded rows with seq_lens_host[b] valid tokens (NULL => all S). Fills the KV cache from
 * position 0. logits_out: B2_LOGITS_LAST -> fp32 [B,vocab] at each sample's last valid position;
 * B2_LOGITS_ALL -> fp32 [B,S,vocab] (the reference's lm_head over all positions, llava_llama.py:88-99). */
int b2_prefill(b2_model* m, b2_kv* kv, const void* embeds, const int32_t* seq_lens_host, int B, int S,
               void* logits_out, int logits_mode, void* stream);
/* b2_prefill into the cache slots [slot0, slot0 + B) — the other slots of the cache are not touched (continuous batching:
 * a new request is prefilled while the rest of the batch keeps its context). b2_prefill == slot0 0. */
int b2_prefill_slots(b2_model* m, b2_kv* kv, const void* embeds, const int32_t* seq_lens_host, int B, int S, int slot0,
                     void* logits_out, int logits_mode, void* stream);
/* One autoregressive step (reference decode branch llava_arch.py:103-112 + HF one-token forward): tokens [B]
 * int32 (host or device) are embedded, run through the decoder against the cache (appending one K/V row per
 * layer), logits_out fp32 [B,vocab] (nullable), next_tokens_out int32 [B] = argmax (nullable, host or device). */
int b2_decode_step(b2_model* m, b2_kv* kv, const int32_t* tokens, int B, void* logits_out, int32_t* next_tokens_out,
                   void* stream);
/* n_steps greedy steps with device-resident token feedback, replayed from a CUDA graph (no host sync between
 * steps): the decode half of HF generate() greedy search as invoked by the reference (model_worker.py:174-185).
 * first_tokens [B] (host or device) is the token fed to step 0; out_tokens [n_steps,B] (host or device) receives
 * the token produced by every step. */
int b2_decode_greedy(b2_model* m, b2_kv* kv, const int32_t* first_tokens, int B, int n_steps, int32_t* out_tokens,
                     void* stream);
int b2_argmax(const float* logits, int B, int V, int32_t* out, void* stream);

/* Streaming decode — what generate() needs when somebody watches every token (the reference always passes a streamer and a
 * stopping criterion: llava/serve/model_worker.py:166-188, llava/serve/cli.py:91-102). The device loop is the same as
 * b2_decode_greedy (token feedback stays on the device, argmax or the temperature/top-k/top-p draw is a kernel), but every
 * step also publishes its token into a ring in mapped pinned host memory, tagged with the generation's epoch, so the host
 * reads token t while step t+k is already running: no D2H copy and no stream synchronisation per token.
 *   b2_stream_begin   selects token 0 from `logits` (device fp32 [B,vocab], the prefill's last-position logits) and
 *                     publishes it as index 0;
 *   b2_stream_enqueue queues n_steps more decode steps (token indices continue from the last one scheduled);
 *   b2_stream_wait    blocks until token `index` is visible and copies it to tokens_host[B] (host). Takes no lock;
 *                     timeout_ms <= 0 waits forever; -3 on timeout, -2 if the device faulted.
 * Steps that were queued past the point where the host decides to stop simply run to completion (rows never interact). */
int b2_stream_begin(b2_model* m, b2_kv* kv, const float* logits, int B, const b2_sampling* sampling, void* stream);
int b2_stream_enqueue(b2_model* m, b2_kv* kv, int n_steps, void* stream);
int b2_stream_wait(b2_kv* kv, int index, int32_t* tokens_host, int timeout_ms);

/* Continuous batching (the reference's worker runs up to limit_model_concurrency generate() threads on one
 * model, llava/serve/model_worker.py:230-243, each a batch-1 HF loop; here they share ONE batched decode step). The B slots of
 * a cache are a pool: b2_batch_begin puts the cache in per-slot mode (every slot idle, streaming ring armed);
 * b2_prefill_slots fills one slot; b2_batch_set_row(active=1) arms it with its own sampling parameters and the token chosen
 * from its prefill logits; b2_stream_enqueue(n) then advances ALL slots by n steps in one batched step each (idle slots keep
 * their length and are ignored) and b2_stream_wait hands the step's B tokens to the host; b2_batch_set_row(active=0) frees a
 * slot. Token selection is per slot: greedy or temperature/top-k/top-p with the slot's own Philox stream. */
int b2_batch_begin(b2_model* m, b2_kv* kv, int B, void* stream);
int b2_batch_set_row(b2_model* m, b2_kv* kv, int slot, int active, const b2_sampling* sampling, int first_token, void* stream);

/* ---- single-kernel entry points (unit-level parity tests; same kernels the hot path launches) ----------- */
int b2_op_gemm(const void* A, int lda, const void* W, int ldw, const void* bias, const void* residual, int ld_res,
               void* out, int ld_out, int out_fp32, int M, int N, int K, int act, int bn_override, void* stream);
int b2_op_gemv(const void* x, int64_t ldx, const void* W, int ldw, const void* norm_gamma, float eps,
               const void* residual, int ld_res, void* out, int ld_out, int out_fp32, int B, int N, int K, int act,
               void* stream);
/* decode Linear at batch 9..128 (swap-AB stream-K tcgen05 GEMM, csrc/gemm_skinny.cu): out[B,N] = x[B,K]·W[N,K]^T
 * (+ residual); act = B2_ACT_NONE | B2_ACT_SWIGLU (out [B,N/2], W rows block-64 interleaved). `workspace` (fp32,
 * >= b2_op_gemm_skinny_workspace_bytes) and `counters` (int32, >= b2_op_gemm_skinny_counter_bytes, zero-filled once
 * by the caller; the kernel leaves them zero) are caller-owned scratch. */
int b2_op_gemm_skinny(const void* x, int ldx, const void* W, int ldw, const void* residual, int ld_res, void* out,
                      int ld_out, int out_fp32, int B, int N, int K, int act, void* workspace, int64_t workspace_bytes,
                      void* counters, void* stream);
/* fp8 variant of b2_op_gemm_skinny: xq [B,K] / Wq [N,K] e4m3 bytes (ld in bytes), x_scale [B], w_scale [N] fp32:
 * out = (xq·Wq^T) * x_scale[b] * w_scale[n] (+ residual). Same scratch contract. */
int b2_op_gemm_skinny_fp8(const void* xq, int ldx, const float* x_scale, const void* Wq, int ldw, const float* w_scale,
                          const void* residual, int ld_res, void* out, int ld_out, int out_fp32, int B, int N, int K, int act,
                          void* workspace, int64_t workspace_bytes, void* counters, void* stream);
/* scale[r] = amax_r / 448 (1 for a zero row); q[r,k] = e4m3_rn_satfinite(x[r,k] * (448 / amax_r)); x bf16, ld in elements */
int b2_op_quantize_rows_e4m3(const void* x, int64_t ldx, int rows, int K, void* q, int64_t ldq, float* scale, void* stream);
/* LlamaRMSNorm (HF rounding points) fused with the per-token quantisation of its output; contiguous rows */
int b2_op_rmsnorm_quant_e4m3(const void* x, const void* gamma, void* q, float* scale, int rows, int cols, float eps,
                             void* stream);
int64_t b2_op_gemm_skinny_workspace_bytes(int B, int N, int K);
int64_t b2_op_gemm_skinny_counter_bytes(int N);
int b2_op_layernorm(const void* x, const void* gamma, const void* beta, void* y, int rows, int cols, float eps,
                    void* stream);
int b2_op_rmsnorm(const void* x, const void* gamma, void* y, int rows, int cols, float eps, void* stream);
/* q,k,v,o: [B,S,H,D] bf16 contiguous; seq_lens device int32 [B] or NULL */
int b2_op_flash_attn(const void* q, const void* k, const void* v, void* o, const int32_t* seq_lens, int B, int S,
                     int H, int D, int causal, float scale, void* stream);
/* qkv [B*S, 3*H*D] (q roped in place); kcache/vcache [B,H,Smax,D] */
int b2_op_rope_kv_write(void* qkv, void* kcache, void* vcache, int B, int S, int H, int D, int Smax, float theta,
                        void* stream);
/* qkv [B,3*H*128]; caches [B,H,Smax,128]; cur_len device int32 [B]; out [B,H*128]; scratch from b2_op_decode_attn_scratch */
int b2_op_decode_attn(const void* qkv, void* kcache, void* vcache, const int32_t* cur_len, void* out, void* scratch,
                      int B, int H, int Smax, int nsplit, float theta, float scale, void* stream);
int64_t b2_op_decode_attn_scratch_bytes(int B, int H, int nsplit); /* caller zero-fills the scratch once */
int b2_op_interleave_gate_up(const void* gate, const void* up, void* out, int I, int h, void* stream);
int b2_op_im2col(const void* pixels, void* out, int B, int img, int patch, int kpad, void* stream);
/* Image preprocessing on the device (csrc/preprocess.cu): the reference's llava/mm_utils.py:16-44 (`expand2square` +
 * CLIPImageProcessor.preprocess = PIL bicubic shortest-edge resize, centre crop, rescale, normalise) for ONE uint8 RGB image.
 * The caller (llava/_b2/preprocess.py) supplies the resize plan: PIL's fixed-point coefficient tables for both axes
 * (bounds[2*i] = first tap, bounds[2*i+1] = tap count, kk[i*ksize + t] = 22-bit fixed-point weight; *_identity != 0 when
 * the axis is not resized), the virtual padding, the crop origin and the source-row window of the vertical pass. All
 * pointers are device pointers. pixels: bf16 [3,out,out] (nullable); u8_out: uint8 [out,out,3] before normalisation (nullable). */
typedef struct b2_preprocess_plan {
    const uint8_t* img; int32_t H, W;
    int32_t pad_top, pad_left; uint8_t bg[4];
    const int32_t *h_bounds, *h_kk; int32_t h_ksize, h_identity;
    const int32_t *v_bounds, *v_kk; int32_t v_ksize, v_identity;
    int32_t y0, rows, x_lo, y_lo, out;
    uint8_t* tmp;                /* scratch, >= rows*out*3 bytes */
    float mean[3], stdv[3], rescale;
    void* pixels; uint8_t* u8_out;
} b2_preprocess_plan;
int b2_op_preprocess_clip(const b2_preprocess_plan* plan, void* stream);
/* one selection per row from fp32 logits [B,V] (csrc/sampling.cu): out_tokens device int32 [B]; `index` is the draw index
 * that keys the Philox stream (token position within a generation). Synchronises the stream. */
int b2_op_sample(const float* logits, int B, int V, const b2_sampling* sampling, int index, int32_t* out_tokens, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B2LLAVA_H_ */
