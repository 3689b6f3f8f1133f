"""GPU: FULL-DEPTH parity at the BASELINE configs: the 32-layer LLaVA-1.5-7B stack on configs[1]'s
inputs (one 336 px image + 128 text tokens, S = 704) and the 40-layer 13B stack, engine (bf16, fp32 accumulation) against
the fp32 oracle on the host cores: vision features, last-position prefill logits and 16 teacher-forced decode steps, with
the measured errors written to gpurun_out/full_depth_parity.json; plus STRICT greedy-id equality on the well-conditioned
weight set (oracle.condition_weights) across the three decode paths (megakernel B=1, GEMV graph B=4, stream-K GEMM B=12).

Stated tolerance at full depth (north_star: "logits within a stated fp tolerance"): max |logit - ref| <= 10 % and mean
<= 2 % of the reference logit std against the fp32 oracle (bf16 rounding of 32-40 residual updates; the same bound the
reference's own bf16 path stays in, see test_model_gpu for the measured bf16-vs-fp32 noise of the restated reference)."""
import json
import os
import time

import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import make_engine, rel_err, synth_inputs  # noqa: E402
from llava import _b2  # noqa: E402
from llava.model.llava_arch import build_source_index  # noqa: E402
from oracle import llava_oracle as O  # noqa: E402

DEV, BF = "cuda", torch.bfloat16
TOL_MAX, TOL_MEAN = 0.10, 0.02   # measured on B200 (profiles/r2c_full_depth_parity.json): 7B 0.063 / 0.0104, 13B 0.081 / 0.015
REPORT = {}


def _device_weights(cfg, seed):
    """bf16 weights generated on the device (a 7B set on the host generator takes minutes), mirrored to the host for the oracle."""
    g = torch.Generator(device=DEV).manual_seed(seed)
    w = {}
    for key, shape, kind in O.weight_shapes(cfg):
        t = torch.randn(*shape, generator=g, device=DEV) * O.init_std(kind, shape)
        if kind == "g":
            t = t + 1.0
        w[key] = t.to(BF)
    # the oracle copy is converted to fp32 ONCE (27 GB for 7B, 52 GB for 13B host memory): converting bf16 -> fp32 inside every
    # oracle matmul made the 41 forward passes of this file cost 5 minutes
    return w, {k: v.cpu().float() for k, v in w.items()}


def _report(name, **kv):
    REPORT[name] = kv
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.path.join("gpurun_out", "full_depth_parity.json"), "w") as f:
        json.dump(REPORT, f, indent=1)
    print(name, json.dumps(kv))


def _teacher_forced(eng, kv, eng_prefill_logits, w_cpu, cfg, embeds_cpu, steps):
    """oracle fp32 prefill + `steps` greedy steps; the engine is fed the ORACLE's tokens. Returns per-step (max, mean) errors
    of the engine logits (step 0 = prefill), the fraction of steps with identical argmax, and the largest oracle margin (in
    logit std) at which the argmax differed."""
    logits, okv = O.llama_forward(w_cpu, embeds_cpu, cfg, last_only=True)
    errs, same, worst_margin = [], 0, None
    eng_logits = eng_prefill_logits
    for t in range(steps + 1):
        ref = logits[:, -1]
        errs.append(rel_err(eng_logits, ref))
        nxt = ref.argmax(-1)
        agree = bool((eng_logits.float().cpu().argmax(-1) == nxt).all())
        same += agree
        if not agree:
            top2 = ref.topk(2, dim=-1).values
            m = float((top2[:, 0] - top2[:, 1]).min() / ref.std())
            worst_margin = m if worst_margin is None else max(worst_margin, m)
        if t == steps:
            break
        e = w_cpu["model.embed_tokens.weight"].float()[nxt][:, None]
        logits, okv = O.llama_forward(w_cpu, e, cfg, kv=okv, last_only=True)
        eng_logits = eng.decode_step(kv, nxt.to(torch.int32))
    return errs, same / (steps + 1), worst_margin


def test_7b_full_depth_configs1_inputs_vs_fp32_oracle():
    torch.set_num_threads(min(32, os.cpu_count() or 1))
    cfg = O.CONFIGS["llava-1.5-7b"]
    t0 = time.time()
    w_dev, w_cpu = _device_weights(cfg, seed=0)
    eng = make_engine(cfg, w_dev, max_batch=12, max_seq=760, max_images=1)
    ids, images = synth_inputs(cfg, B=1, Lt=129, seed=1)          # BASELINE configs[1]: 576 image tokens + 128 text
    images = images.to(BF).float()
    # ---- vision tower + projector ----
    feats_ref = O.encode_images(w_cpu, images, cfg)
    feats = eng.encode_images(images.to(DEV))
    fe = rel_err(feats, feats_ref)
    # ---- prefill (S = 704) + 16 teacher-forced decode steps ----
    embeds_ref, _, _, _ = O.prepare_multimodal(w_cpu, ids, None, cfg, image_features=list(feats_ref))
    src, _, _, _, lens = build_source_index(ids.numpy(), torch.ones_like(ids, dtype=torch.bool).numpy(),
                                            torch.full_like(ids, -100).numpy(), 576, [576], None, "right")
    embeds = eng.splice(torch.from_numpy(src.reshape(-1)).to(DEV), feats.view(-1, cfg["hidden"]), 1, src.shape[1])
    assert src.shape[1] == 704
    kv = eng.new_kv(1, 760)
    errs, agree, worst = _teacher_forced(eng, kv, eng.prefill(kv, embeds, lens, _b2.LOGITS_LAST), w_cpu, cfg, embeds_ref, steps=16)
    mx, mn = max(e[0] for e in errs), max(e[1] for e in errs)
    _report("7b_32_layers_S704", encode_images_err=fe, prefill_err=errs[0], decode_err_max_over_16_steps=(mx, mn),
            argmax_agreement=agree, largest_margin_of_a_disagreement_in_std=worst, seconds=round(time.time() - t0, 1))
    assert fe[0] <= TOL_MAX and fe[1] <= TOL_MEAN, fe
    assert mx <= TOL_MAX and mn <= TOL_MEAN, (mx, mn)
    assert worst is None or worst <= 2 * mx, f"argmax differs at a margin of {worst} std with logit error {mx}"
    kv.close()

    # ---- strict greedy ids on the well-conditioned weight set, three decode paths ----
    wc_cpu = O.condition_weights(w_cpu, cfg, seed=0)
    changed = [k for k in wc_cpu if wc_cpu[k] is not w_cpu[k]]
    eng.close()
    for k in changed:
        w_dev[k] = wc_cpu[k].to(DEV, BF)
    eng = make_engine(cfg, w_dev, max_batch=12, max_seq=128, max_images=1)
    g = torch.Generator().manual_seed(5)
    prompt = torch.randint(3, cfg["vocab"], (1, 24), generator=g)
    emb = wc_cpu["model.embed_tokens.weight"].float()[prompt]
    logits, okv = O.llama_forward(wc_cpu, emb, cfg, last_only=True)
    want, margins = [], []
    for t in range(16):
        ref = logits[:, -1]
        top2 = ref.topk(2, dim=-1).values
        margins.append(float((top2[:, 0] - top2[:, 1]).min() / ref.std()))
        nxt = ref.argmax(-1)
        want.append(int(nxt))
        logits, okv = O.llama_forward(wc_cpu, wc_cpu["model.embed_tokens.weight"].float()[nxt][:, None], cfg, kv=okv, last_only=True)
    got = {}
    for B in (1, 4, 12):
        kvb = eng.new_kv(B, 128)
        ids_b = prompt.repeat(B, 1).to(torch.int32).reshape(-1).to(DEV)
        lg = eng.prefill(kvb, eng.splice(ids_b, None, B, 24), None, _b2.LOGITS_LAST)
        first = eng.argmax(lg)
        rest = eng.decode_greedy(kvb, first, 15).cpu()
        got[B] = torch.cat([first.cpu()[None], rest]).t().tolist()
        kvb.close()
    _report("7b_32_layers_conditioned_ids", oracle_ids=want, min_margin_in_std=min(margins),
            paths={"megakernel_B1": got[1][0] == want, "gemv_graph_B4": all(r == want for r in got[4]),
                   "stream_k_gemm_B12": all(r == want for r in got[12])})
    assert min(margins) > 1.0
    for B in (1, 4, 12):
        for r in got[B]:
            assert r == want, (B, r, want)
    eng.close()


def test_13b_full_depth_vs_fp32_oracle():
    torch.set_num_threads(min(32, os.cpu_count() or 1))
    cfg = dict(O.CONFIGS["llava-1.5-13b"], vit_layers=2)  # the tower is depth-tested at 7B; 13B differs in the decoder only
    t0 = time.time()
    w_dev, w_cpu = _device_weights(cfg, seed=1)
    eng = make_engine(cfg, w_dev, max_batch=1, max_seq=96, max_images=1)
    g = torch.Generator().manual_seed(2)
    prompt = torch.randint(3, cfg["vocab"], (1, 48), generator=g)
    embeds_ref = w_cpu["model.embed_tokens.weight"].float()[prompt]
    kv = eng.new_kv(1, 96)
    embeds = eng.splice(prompt.to(torch.int32).reshape(-1).to(DEV), None, 1, 48)
    errs, agree, worst = _teacher_forced(eng, kv, eng.prefill(kv, embeds, None, _b2.LOGITS_LAST), w_cpu, cfg, embeds_ref, steps=8)
    mx, mn = max(e[0] for e in errs), max(e[1] for e in errs)
    _report("13b_40_layers_S48", prefill_err=errs[0], decode_err_max_over_8_steps=(mx, mn), argmax_agreement=agree,
            largest_margin_of_a_disagreement_in_std=worst, seconds=round(time.time() - t0, 1))
    assert mx <= TOL_MAX and mn <= TOL_MEAN, (mx, mn)
    assert worst is None or worst <= 2 * mx
    kv.close()
    eng.close()
