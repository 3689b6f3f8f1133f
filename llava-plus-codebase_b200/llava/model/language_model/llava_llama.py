"""LlavaLlamaForCausalLM on the B200 engine.

Same public surface as the reference's llava/model/language_model/llava_llama.py (LlavaConfig :29-30,
LlavaLlamaModel :33-38, LlavaLlamaForCausalLM.forward :56-99, prepare_inputs_for_generation :101-108) and
the same state-dict key layout, so `load_pretrained_model`, `llava.serve.model_worker`,
`llava.serve.cli` and the eval scripts can drive it unchanged. The modules here only HOLD checkpoint tensors;
all arithmetic of the path (CLIP ViT, projector, splice, LLaMA prefill + KV-cache decode, lm_head, greedy
argmax) runs in libb2llava.so through the C ABI in include/b2llava.h. There is no PyTorch fallback: without
the CUDA library / an sm_100 device every compute entry point raises.

Compute dtype is bf16 (north-star). Callers that ask for fp16 (`.half()`, torch_dtype=float16 in the
reference's builder.py:43) keep their tensor dtypes at the boundary, the engine computes in bf16.
"""
import json
import os
import threading
import weakref
from typing import List, Optional, Tuple, Union

import torch
import torch.nn as nn
from transformers import AutoConfig, LlamaConfig
from transformers.modeling_outputs import CausalLMOutputWithPast

from ..llava_arch import LlavaMetaModel, LlavaMetaForCausalLM
from ..multimodal_encoder.clip_encoder import _Holder, _read_checkpoint_dir
from ..._b2 import Engine, KVCache, LOGITS_ALL, LOGITS_LAST, ERR_SPLICE_SLOTS, last_error, make_sampling


class LlavaConfig(LlamaConfig):
    model_type = "llava"


class _KVPool:
    """KV caches of one engine. Every generate() works on a cache of its own for its whole duration — the reference's
    model_worker runs up to `limit_model_concurrency` generate() threads on one model object
    (llava/serve/model_worker.py:174-185, :230-243), and a shared cache would let request B's prefill overwrite request
    A's context mid-decode. Caches are created on demand (up to `cap`, then acquire() blocks) and reused."""

    def __init__(self, engine, max_batch, max_seq, cap):
        self.engine, self.max_batch, self.max_seq, self.cap = engine, max_batch, max_seq, max(1, int(cap))
        self._cond = threading.Condition()
        self._free, self._made = [], 0

    def acquire(self):
        with self._cond:
            while not self._free and self._made >= self.cap:
                self._cond.wait()
            if self._free:
                return self._free.pop()
            self._made += 1
        try:
            return self.engine.new_kv(self.max_batch, self.max_seq)
        except Exception:
            with self._cond:
                self._made -= 1
                self._cond.notify()
            raise

    def release(self, kv):
        with self._cond:
            self._free.append(kv)
            self._cond.notify()


class PastKeyValues:
    """What forward(use_cache=True) returns as `past_key_values`: a lease on one of the model's forward caches. A later
    forward() may recycle the underlying cache (`config.b2_forward_caches` of them exist, default 2); using a recycled
    lease raises instead of silently decoding against somebody else's context."""

    def __init__(self, kv, serial):
        self.kv, self.serial = kv, serial

    @property
    def valid(self):
        return getattr(self.kv, "_lease_serial", None) == self.serial

    def get_seq_length(self, layer_idx=0):
        return self.kv.get_seq_length()


def _empty_param(*shape, dtype=None, device=None):
    return nn.Parameter(torch.empty(*shape, dtype=dtype, device=device), requires_grad=False)


class _Weight(nn.Module):
    """`<name>.weight` holder (Linear without bias / RMSNorm / Embedding), never initialised on construction."""

    def __init__(self, *shape, dtype=None, device=None):
        super().__init__()
        self.weight = _empty_param(*shape, dtype=dtype, device=device)


class _DecoderWeights(nn.Module):
    """State-dict layout of transformers LlamaModel: embed_tokens, layers.N.{self_attn,mlp,*_layernorm}, norm."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        h, I, V = config.hidden_size, config.intermediate_size, config.vocab_size
        dt = getattr(config, "_b2_param_dtype", torch.bfloat16)
        dev = getattr(config, "_b2_param_device", None)
        self.embed_tokens = _Weight(V, h, dtype=dt, device=dev)
        layers = []
        for _ in range(config.num_hidden_layers):
            L = _Holder()
            L.self_attn = _Holder()
            for n in ("q_proj", "k_proj", "v_proj", "o_proj"):
                setattr(L.self_attn, n, _Weight(h, h, dtype=dt, device=dev))
            L.mlp = _Holder()
            L.mlp.gate_proj = _Weight(I, h, dtype=dt, device=dev)
            L.mlp.up_proj = _Weight(I, h, dtype=dt, device=dev)
            L.mlp.down_proj = _Weight(h, I, dtype=dt, device=dev)
            L.input_layernorm = _Weight(h, dtype=dt, device=dev)
            L.post_attention_layernorm = _Weight(h, dtype=dt, device=dev)
            layers.append(L)
        self.layers = nn.ModuleList(layers)
        self.norm = _Weight(h, dtype=dt, device=dev)


class LlavaLlamaModel(LlavaMetaModel, _DecoderWeights):
    config_class = LlavaConfig

    def __init__(self, config):
        super(LlavaLlamaModel, self).__init__(config)


class LlavaLlamaForCausalLM(nn.Module, LlavaMetaForCausalLM):
    config_class = LlavaConfig

    def __init__(self, config, device=None, dtype=torch.bfloat16, max_batch=None, max_seq=None, max_images=None):
        nn.Module.__init__(self)
        if getattr(config, "num_key_value_heads", config.num_attention_heads) != config.num_attention_heads:
            raise NotImplementedError("grouped-query attention is not part of the LLaVA-1.5 path (kv_heads == heads)")
        if getattr(config, "pretraining_tp", 1) != 1:
            raise NotImplementedError("pretraining_tp != 1")
        self.config = config
        config._b2_param_dtype, config._b2_param_device = dtype, device
        try:
            self.model = LlavaLlamaModel(config)
        finally:
            del config._b2_param_dtype, config._b2_param_device
        self.vocab_size = config.vocab_size
        self.lm_head = _Weight(config.vocab_size, config.hidden_size, dtype=dtype, device=device)
        self._engine = None
        self._batcher = None       # config.b2_continuous_batching = N: concurrent generate() calls share batched decode steps
        self._pool = None          # generate(): one exclusive cache per call
        self._fwd_kvs = []         # forward(use_cache=True): small LRU ring of leased caches
        self._fwd_lock = threading.Lock()
        self._engine_lock = threading.RLock()
        self._limits = {"max_batch": max_batch, "max_seq": max_seq, "max_images": max_images}
        self._attach()

    # ------------------------------------------------------------------ plumbing
    def _attach(self):
        ref = weakref.ref(self)
        self.model._owner = ref
        if getattr(self.model, "mm_projector", None) is not None:
            self.model.mm_projector._owner = ref
        vt = self.model.get_vision_tower()
        if vt is not None:
            vt._owner = ref

    def get_model(self):
        return self.model

    def get_input_embeddings(self):
        return self.model.embed_tokens

    def get_output_embeddings(self):
        return self.lm_head

    @property
    def device(self):
        return self.lm_head.weight.device

    @property
    def dtype(self):
        return self.lm_head.weight.dtype

    def invalidate_engine(self):
        """Weights changed (load_state_dict, resize, tower load): rebuild the device engine lazily."""
        with self._engine_lock:
            if self._batcher is not None:
                self._batcher.close()
                self._batcher = None
            self._pool = None
            self._fwd_kvs = []
            if self._engine is not None:
                self._engine.close()
            self._engine = None

    def load_state_dict(self, state_dict, strict=True, assign=False):
        r = super().load_state_dict(state_dict, strict=strict, assign=assign)
        self.invalidate_engine()
        return r

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self.invalidate_engine()
        return r

    def resize_token_embeddings(self, new_num_tokens=None):
        old = self.model.embed_tokens.weight
        if new_num_tokens is None or new_num_tokens == old.shape[0]:
            return self.model.embed_tokens
        for mod in (self.model.embed_tokens, self.lm_head):
            w = mod.weight.data
            nw = torch.zeros(new_num_tokens, w.shape[1], dtype=w.dtype, device=w.device)
            n = min(new_num_tokens, w.shape[0])
            nw[:n] = w[:n]
            if new_num_tokens > n:
                nw[n:].normal_(mean=0.0, std=getattr(self.config, "initializer_range", 0.02))
            mod.weight = nn.Parameter(nw, requires_grad=False)
        self.config.vocab_size = self.vocab_size = new_num_tokens
        self.invalidate_engine()
        return self.model.embed_tokens

    def engine_limits(self, max_batch=None, max_seq=None, max_images=None):
        """Workspace sizing of the device engine (KV cache + activations); call before the first forward."""
        for k, v in (("max_batch", max_batch), ("max_seq", max_seq), ("max_images", max_images)):
            if v is not None and v != self._limits[k]:
                self._limits[k] = v
                self.invalidate_engine()

    def _ensure_engine(self) -> Engine:
        eng = self._engine
        if eng is not None:
            return eng
        with self._engine_lock:
            return self._build_engine()

    def _build_engine(self) -> Engine:
        if self._engine is not None:
            return self._engine
        vt = self.get_vision_tower()
        if vt is None or not vt.is_loaded:
            raise RuntimeError("vision tower is not loaded: call model.get_vision_tower().load_model() first")
        if self.device.type != "cuda":
            raise RuntimeError("LlavaLlamaForCausalLM weights must be on a CUDA (sm_100a) device: model.to('cuda'); "
                               "there is no CPU path")
        c, vc = self.config, vt.config
        max_seq = self._limits["max_seq"] or min(getattr(c, "max_position_embeddings", 4096), 4096)
        desc = dict(
            image_size=vc.image_size, patch_size=vc.patch_size, vit_hidden=vc.hidden_size,
            vit_inter=vc.intermediate_size, vit_layers=vc.num_hidden_layers, vit_heads=vc.num_attention_heads,
            vit_select_layer=vt.select_layer, vit_ln_eps=vc.layer_norm_eps,
            hidden=c.hidden_size, inter=c.intermediate_size, layers=c.num_hidden_layers,
            heads=c.num_attention_heads, vocab=self.lm_head.weight.shape[0], rms_eps=c.rms_norm_eps,
            rope_theta=float(getattr(c, "rope_theta", None) or (getattr(c, "rope_parameters", None) or {}).get("rope_theta", 10000.0)),
            max_batch=max(self._limits["max_batch"] or 1, int(getattr(c, "b2_continuous_batching", 0) or 0)),
            max_seq=max_seq, max_images=self._limits["max_images"] or 8,
        )
        if getattr(vc, "hidden_act", "quick_gelu") != "quick_gelu":
            raise NotImplementedError("CLIP tower activation must be quick_gelu (openai/clip-vit-large-patch14-336)")
        eng = Engine(desc, self.device)
        torch.cuda.current_stream(self.device).synchronize()  # set_weight copies on the library's stream: order it after
        for k, v in self.state_dict().items():                # whatever produced the parameters on the caller's stream
            if "position_ids" in k or "inv_freq" in k:
                continue
            eng.set_weight(k, v.to(self.device))
        eng.finalize()
        # BASELINE configs[4] opt-in: e4m3 decoder weights for batch >= 7 decode (config.b2_fp8_decode or B2_FP8_DECODE=1)
        if getattr(c, "b2_fp8_decode", False) or os.environ.get("B2_FP8_DECODE") == "1":
            eng.enable_fp8_decode()
        self._pool = _KVPool(eng, desc["max_batch"], desc["max_seq"],
                             getattr(c, "b2_max_concurrent_generations", None) or int(os.environ.get("B2_MAX_GENERATIONS", "8")))
        self._fwd_kvs = []
        self._engine = eng
        return eng

    def _get_batcher(self, eng):
        slots = int(getattr(self.config, "b2_continuous_batching", 0) or 0)
        if slots < 2:
            return None
        with self._engine_lock:
            if self._batcher is None:
                from ..._b2.batching import ContinuousBatcher
                self._batcher = ContinuousBatcher(eng, slots, eng.desc.max_seq)
            return self._batcher

    def _check_limits(self, eng, B, need_seq):
        lim_b, lim_s = eng.desc.max_batch, eng.desc.max_seq
        if B > lim_b or need_seq > lim_s:
            raise ValueError(f"batch {B} / sequence {need_seq} exceed the engine limits (max_batch={lim_b}, max_seq={lim_s}); "
                             f"call model.engine_limits(max_batch=..., max_seq=...) before the first forward")

    def _lease_forward_kv(self, eng, B, need_seq):
        """Cache for one forward(use_cache=True): least recently leased of a small ring; earlier leases of it go stale."""
        self._check_limits(eng, B, need_seq)
        with self._fwd_lock:
            cap = max(1, int(getattr(self.config, "b2_forward_caches", 2)))
            if len(self._fwd_kvs) < cap:
                kv = eng.new_kv(eng.desc.max_batch, eng.desc.max_seq)
                kv._lease_serial = 0
            else:
                kv = self._fwd_kvs.pop(0)
            kv._lease_serial += 1
            self._fwd_kvs.append(kv)
            return PastKeyValues(kv, kv._lease_serial)

    # ------------------------------------------------------------------ forward (ref llava_llama.py:56-99)
    def forward(
        self,
        input_ids: torch.LongTensor = None,
        attention_mask: Optional[torch.Tensor] = None,
        position_ids: Optional[torch.LongTensor] = None,
        past_key_values=None,
        inputs_embeds: Optional[torch.FloatTensor] = None,
        labels: Optional[torch.LongTensor] = None,
        use_cache: Optional[bool] = None,
        output_attentions: Optional[bool] = None,
        output_hidden_states: Optional[bool] = None,
        images: Optional[torch.FloatTensor] = None,
        return_dict: Optional[bool] = None,
    ) -> Union[Tuple, CausalLMOutputWithPast]:
        if output_attentions or output_hidden_states:
            raise NotImplementedError("attention maps / hidden states are never materialised on the fused path")
        engine = self._ensure_engine()

        # ---- decode step: [B,1] ids against an engine cache ----
        if inputs_embeds is None and past_key_values is not None and input_ids is not None and input_ids.shape[1] == 1:
            if isinstance(past_key_values, PastKeyValues):
                if not past_key_values.valid:
                    raise RuntimeError("this past_key_values was recycled by a later forward(use_cache=True): the model keeps "
                                       "config.b2_forward_caches (default 2) forward caches alive at a time")
                kv = past_key_values.kv
            elif isinstance(past_key_values, KVCache):
                kv = past_key_values
            else:
                raise ValueError("past_key_values must be the cache object returned by a previous forward of this model")
            logits = engine.decode_step(kv, input_ids.reshape(-1))
            logits = logits.unsqueeze(1)
            return self._output(logits, past_key_values, labels, return_dict)

        if inputs_embeds is None:
            (input_ids, position_ids, attention_mask, past_key_values, inputs_embeds, labels) = \
                self.prepare_inputs_labels_for_multimodal(input_ids, position_ids, attention_mask, past_key_values, labels, images)
            if inputs_embeds is None:  # text-only (no images / no tower): plain embedding lookup on the engine
                ids = input_ids.to(torch.int32).reshape(-1).to(engine.device)
                inputs_embeds = engine.splice(ids, None, input_ids.shape[0], input_ids.shape[1])
        if past_key_values is not None:
            raise NotImplementedError("chunked prefill against an existing cache is not part of the reference's call pattern")

        B, S = inputs_embeds.shape[0], inputs_embeds.shape[1]
        lens, left = None, False
        if attention_mask is not None:
            m = attention_mask.to(device=inputs_embeds.device).bool()
            lens = m.sum(dim=1).tolist()
            left = bool((~m[:, 0]).any()) and bool(m[:, -1].all())
            if left:  # engine rows are right-padded: rotate valid tokens to the front (ref pads left only on request)
                inputs_embeds = torch.stack([torch.roll(inputs_embeds[b], shifts=-(S - lens[b]), dims=0) for b in range(B)])
        lease = self._lease_forward_kv(engine, B, S)
        lease.kv.reset()
        logits = engine.prefill(lease.kv, inputs_embeds, lens, LOGITS_ALL)
        if left:
            logits = torch.stack([torch.roll(logits[b], shifts=(S - lens[b]), dims=0) for b in range(B)])
        return self._output(logits, lease if use_cache is not False else None, labels, return_dict)

    def _output(self, logits, kv, labels, return_dict):
        loss = None
        if labels is not None:  # HF LlamaForCausalLM loss: shift by one, ignore_index=-100
            sl = logits[..., :-1, :].contiguous().float()
            tl = labels[..., 1:].contiguous().to(sl.device)
            loss = nn.functional.cross_entropy(sl.view(-1, sl.size(-1)), tl.view(-1), ignore_index=-100)
        if return_dict is False:
            out = (logits, kv)
            return ((loss,) + out) if loss is not None else out
        return CausalLMOutputWithPast(loss=loss, logits=logits, past_key_values=kv)

    def prepare_inputs_for_generation(self, input_ids, past_key_values=None, inputs_embeds=None, **kwargs):
        """ref llava_llama.py:101-108 (kept for API parity; generate() below does not need it)."""
        images = kwargs.pop("images", None)
        if past_key_values is not None:
            input_ids = input_ids[:, -1:]
        model_inputs = {"input_ids": input_ids, "past_key_values": past_key_values,
                        "use_cache": kwargs.get("use_cache"), "attention_mask": kwargs.get("attention_mask")}
        if inputs_embeds is not None and past_key_values is None:
            model_inputs = {"inputs_embeds": inputs_embeds, **{k: v for k, v in model_inputs.items() if k != "input_ids"}}
        if images is not None:
            model_inputs["images"] = images
        return model_inputs

    # ------------------------------------------------------------------ generate
    # generation arguments of HF generate() that change WHAT is generated and that this loop does not implement: raise
    # instead of silently decoding something else (value = the setting that means "off")
    _UNSUPPORTED_GENERATION_ARGS = {
        "repetition_penalty": 1.0, "encoder_repetition_penalty": 1.0, "length_penalty": 1.0, "no_repeat_ngram_size": 0,
        "encoder_no_repeat_ngram_size": 0, "min_new_tokens": None, "min_length": 0, "bad_words_ids": None,
        "force_words_ids": None, "num_beam_groups": 1, "diversity_penalty": 0.0, "penalty_alpha": None, "typical_p": 1.0,
        "epsilon_cutoff": 0.0, "eta_cutoff": 0.0, "num_return_sequences": 1, "logits_processor": None,
        "prefix_allowed_tokens_fn": None, "constraints": None, "suppress_tokens": None, "begin_suppress_tokens": None,
        "forced_bos_token_id": None, "forced_eos_token_id": None, "assistant_model": None, "min_p": None,
    }
    # accepted and without effect on this path
    _IGNORED_GENERATION_ARGS = {"synced_gpus", "early_stopping", "output_attentions", "output_hidden_states",
                                "generation_config", "position_ids", "past_key_values", "renormalize_logits", "max_time"}

    @torch.no_grad()
    def generate(self, inputs=None, images=None, do_sample=False, temperature=1.0, top_p=None, top_k=None,
                 num_beams=1, max_new_tokens=None, max_length=None, use_cache=True, streamer=None,
                 stopping_criteria=None, eos_token_id=None, pad_token_id=None, attention_mask=None,
                 input_ids=None, output_scores=False, return_dict_in_generate=False, **kwargs):
        """Own decoding loop with the side-protocols the reference's callers rely on: prompt ids (with
        IMAGE_TOKEN_INDEX) echoed in the result, `streamer.put/end`, `stopping_criteria` called as
        crit(ids_so_far, scores) -> bool | bool tensor, eos stop, temperature / top-k / top-p sampling.

        Every variant runs the same device-resident loop: token feedback, argmax or the sampling draw are kernels, each
        step publishes its token into pinned host memory, and this thread (usually a worker `Thread`,
        llava/serve/model_worker.py:174-185) reads token t — streamer, eos, stopping criteria — while the device is
        already `config.b2_run_ahead` (default 8) steps further. Steps queued beyond the stop point are discarded."""
        if inputs is None:
            inputs = input_ids
        if inputs is None:
            raise ValueError("generate() needs input ids")
        if num_beams != 1:
            raise NotImplementedError("beam search is not used on the LLaVA path (num_beams=1 everywhere)")
        if return_dict_in_generate or output_scores:
            raise NotImplementedError("generate() returns the id tensor only")
        for k, v in kwargs.items():
            if k in self._IGNORED_GENERATION_ARGS:
                continue
            if k in self._UNSUPPORTED_GENERATION_ARGS:
                if v is None or v == self._UNSUPPORTED_GENERATION_ARGS[k]:
                    continue
                raise NotImplementedError(f"generate({k}={v!r}) is not implemented on the B200 path")
            raise NotImplementedError(f"generate() got an unsupported argument {k!r}")
        prompt = inputs if inputs.dim() == 2 else inputs.unsqueeze(0)
        B, Lt = prompt.shape
        engine = self._ensure_engine()
        if eos_token_id is None:
            eos_token_id = getattr(self.config, "eos_token_id", None)
        eos_ids = set(eos_token_id) if isinstance(eos_token_id, (list, tuple)) else ({eos_token_id} if eos_token_id is not None else set())
        if max_new_tokens is None:
            max_new_tokens = (max_length - Lt) if max_length is not None else 20
        if max_new_tokens <= 0:
            raise ValueError("max_new_tokens must be positive")
        greedy = (not do_sample) or (temperature is not None and temperature <= 1e-5)
        if greedy:
            sampling = make_sampling()
        else:
            if top_p is not None and not (0.0 < top_p <= 1.0):
                raise ValueError(f"top_p must be in (0, 1], got {top_p}")
            # HF GenerationConfig defaults: top_k = 50, top_p = 1.0; the draw is seeded from torch's CPU generator so that
            # torch.manual_seed() makes a run repeatable
            seed = int(torch.randint(0, 2**62, (1,), dtype=torch.int64).item())
            sampling = make_sampling(True, temperature, 1.0 if top_p is None else top_p, 50 if top_k is None else top_k, seed)
        prof = _StageTimer() if os.environ.get("B2_PROFILE_GENERATE") else None

        batcher = self._get_batcher(engine) if B == 1 else None
        if batcher is not None:
            # continuous batching: this thread splices its prompt and runs its own host loop (streamer, eos, criteria); the
            # decode steps are shared with every other generate() in flight (llava/_b2/batching.py)
            from ..._b2.batching import RequestStream
            embeds = lens = None
            if images is not None and self.get_vision_tower() is not None:
                embeds, lens, speculative = self._spliced_embeds(prompt, attention_mask, images)
                if speculative:
                    torch.cuda.current_stream(engine.device).synchronize()
                    if engine.take_async_error() & ERR_SPLICE_SLOTS:
                        embeds, lens, _ = self._spliced_embeds(prompt, attention_mask, images, force_host=True)
            if embeds is None:
                embeds = engine.splice(prompt.to(torch.int32).reshape(-1).to(engine.device), None, 1, Lt)
                lens = [Lt]
            self._check_limits(engine, 1, lens[0] + max_new_tokens)
            req = batcher.submit(embeds, lens[0], sampling, max_new_tokens)
            if streamer is not None:
                streamer.put(prompt.cpu())
            pad = pad_token_id if pad_token_id is not None else (next(iter(eos_ids)) if eos_ids else 0)
            try:
                new_tokens = _stream_decode(RequestStream(req), None, None, sampling, 1, max_new_tokens, eos_ids, pad, prompt,
                                            streamer, stopping_criteria, begun=True)
            finally:
                req.cancel()
            if streamer is not None:
                streamer.end()
            return torch.cat([prompt, new_tokens.to(device=prompt.device, dtype=prompt.dtype)], dim=1)

        kv = self._pool.acquire()  # exclusive for this call (concurrent generate() threads each get their own)
        try:
            # ---- prefill: splice + decoder, last-position logits only; token 0 is chosen on the device ----
            def prefill(force_host):
                embeds, lens, speculative = None, None, False
                if images is not None and self.get_vision_tower() is not None:
                    embeds, lens, speculative = self._spliced_embeds(prompt, attention_mask, images, force_host=force_host)
                if embeds is not None:
                    if getattr(self.config, "tokenizer_padding_side", "right") == "left" and len(set(lens)) > 1:
                        S = embeds.shape[1]
                        embeds = torch.stack([torch.roll(embeds[b], shifts=-(S - lens[b]), dims=0) for b in range(B)])
                else:  # text-only prompt (or a [B,1] prompt, which the multimodal splice passes through)
                    if attention_mask is not None and not bool(attention_mask.bool().all()):
                        raise NotImplementedError("padded text-only batches: pass equal-length prompts")
                    ids = prompt.to(torch.int32).reshape(-1).to(engine.device)
                    embeds = engine.splice(ids, None, B, Lt)
                    lens = [Lt] * B
                if prof: prof.mark("encode_images+splice")
                self._check_limits(engine, B, max(lens) + max_new_tokens)
                kv.reset()
                logits = engine.prefill(kv, embeds, lens, LOGITS_LAST)
                engine.stream_begin(kv, logits, sampling)
                engine.stream_wait(kv, 0, B)          # first sync of this call: every input check has run by now
                if prof: prof.mark("prefill + first token")
                return speculative

            speculative = prefill(False)
            err = engine.take_async_error()
            if speculative and (err & ERR_SPLICE_SLOTS):
                # the rows do not hold n_images / B placeholders each: the shape-only output length was wrong. Redo on the
                # host path, which reproduces the reference's behaviour for that input (extra images ignored / IndexError)
                speculative = prefill(True)
                err = engine.take_async_error()
            if err:
                raise ValueError(last_error())

            if streamer is not None:
                streamer.put(prompt.cpu())
            pad = pad_token_id if pad_token_id is not None else (next(iter(eos_ids)) if eos_ids else 0)
            new_tokens = _stream_decode(engine, kv, None, sampling, B, max_new_tokens, eos_ids, pad, prompt,
                                        streamer, stopping_criteria,
                                        run_ahead=int(getattr(self.config, "b2_run_ahead", 8)), begun=True)
            engine.check_async_error()
            if prof: prof.mark("decode")
        finally:
            self._pool.release(kv)
        if streamer is not None:
            streamer.end()
        out = torch.cat([prompt, new_tokens.to(device=prompt.device, dtype=prompt.dtype)], dim=1)
        if prof: prof.mark("ids to caller"); prof.report()
        return out

    # ------------------------------------------------------------------ checkpoints
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, *model_args, config=None, torch_dtype=None,
                        low_cpu_mem_usage=True, device_map=None, device=None, **kwargs):
        """Minimal HF-style loader: `config.json` + safetensors / pytorch_model*.bin shards in a local directory
        (ref builder.py:105-106 calls this). 8-bit/4-bit bitsandbytes loading is not part of the B200 path."""
        if kwargs.get("load_in_8bit") or kwargs.get("load_in_4bit") or kwargs.get("quantization_config") is not None:
            raise NotImplementedError("bitsandbytes quantised loading is not supported on the B200 path")
        path = pretrained_model_name_or_path
        if not os.path.isdir(path):
            raise ValueError(f"{path!r} is not a local checkpoint directory (no network access on this path)")
        if config is None:
            with open(os.path.join(path, "config.json")) as f:
                config = LlavaConfig(**json.load(f))
        dev = device
        if dev is None and isinstance(device_map, (str, torch.device)) and str(device_map) not in ("auto",):
            dev = device_map
        if dev is None:
            dev = "cuda"
        model = cls(config, device=dev, dtype=torch.bfloat16)
        sd = _read_checkpoint_dir(path)
        if not sd:
            raise RuntimeError(f"no weight files found under {path!r}")
        own = model.state_dict()
        sd = {k: v for k, v in sd.items() if k in own}
        missing = [k for k in own if k not in sd and "vision_tower" not in k]
        if missing:
            raise RuntimeError(f"checkpoint is missing tensors: {missing[:8]} ...")
        model.load_state_dict(sd, strict=False)
        return model

    def eval(self):
        return super().eval()


def _stream_decode(engine, kv, logits, sampling, B, max_new_tokens, eos_ids, pad, prompt, streamer, stopping_criteria,
                   run_ahead=8, begun=False):
    """Host half of the decode loop. The device chooses token 0 from the prefill `logits` and then runs up to `run_ahead`
    steps in front of this loop; `engine.stream_wait(kv, t, B)` hands over token t as soon as its kernel has written it
    to pinned memory. Per token, in the order HF's loop uses: finished rows show `pad`; streamer.put; eos bookkeeping;
    stopping criteria on cat(prompt, tokens so far) (the reference's KeywordsStoppingCriteria looks at the tail of that
    tensor, llava/mm_utils.py:92-114). Returns a CPU int64 tensor [B, n], 1 <= n <= max_new_tokens."""
    Lt = prompt.shape[1]
    crit_buf = None
    if stopping_criteria:
        crit_buf = torch.empty(B, Lt + max_new_tokens, dtype=torch.long)
        crit_buf[:, :Lt] = prompt.to("cpu", torch.long)
    run_ahead = max(1, int(run_ahead))
    if not begun:  # generate() begins the stream itself (it confirms its input checks on token 0 before any output)
        engine.stream_begin(kv, logits, sampling)
    scheduled = 1                                   # tokens whose kernels have been queued (token 0 = begin)
    finished = [False] * B
    cols = []
    for t in range(max_new_tokens):
        # keep the device `run_ahead` tokens in front of the host (queued in blocks: one call per ~run_ahead/2 tokens)
        if scheduled < max_new_tokens and scheduled - t <= (run_ahead + 1) // 2:
            n = min(run_ahead - (scheduled - t) + 1, max_new_tokens - scheduled)
            if n > 0:
                engine.stream_enqueue(kv, n)
                scheduled += n
        toks = engine.stream_wait(kv, t, B)
        col = [pad if finished[b] else int(toks[b]) for b in range(B)]
        cols.append(col)
        if streamer is not None:
            streamer.put(torch.tensor(col, dtype=torch.long))
        for b in range(B):
            if col[b] in eos_ids:
                finished[b] = True
        stop = bool(eos_ids) and all(finished)
        if crit_buf is not None:
            crit_buf[:, Lt + t] = torch.tensor(col, dtype=torch.long)
            view = crit_buf[:, :Lt + t + 1]
            for crit in stopping_criteria:
                r = crit(view, None)
                stop = stop or (bool(r.all()) if torch.is_tensor(r) else bool(r))
        if stop:
            break
    return torch.tensor(cols, dtype=torch.long).t().contiguous()


class _StageTimer:
    """B2_PROFILE_GENERATE=1: wall-clock per stage of generate() with a device sync at every mark (debug aid; the
    syncs serialise host and device, so the stages add up to MORE than an unprofiled call)."""

    def __init__(self):
        import time
        self._t, self._clock, self._rows = time.perf_counter(), time.perf_counter, []

    def mark(self, name):
        torch.cuda.synchronize()
        t = self._clock()
        self._rows.append((name, (t - self._t) * 1e3))
        self._t = t

    def report(self):
        import sys
        print("generate stages (ms): " + ", ".join(f"{n} {ms:.2f}" for n, ms in self._rows), file=sys.stderr)


try:  # the installed transformers may already ship a "llava" model type (ref llava_llama.py:110-111)
    AutoConfig.register("llava", LlavaConfig, exist_ok=True)
except Exception:  # pragma: no cover - registry differences across transformers versions
    pass
