"""Sweep of the BASELINE.json configurations that are not bench.py's headline line (configs[2], [3] and the
batch ladder of configs[1]): device-resident timings with CUDA events on the launching stream, one model build
per size. Writes one JSON object per measurement to stdout (and to gpurun_out/config_sweep.jsonl).

    python scripts/config_sweep.py --model 7b --vit 1,16,64,256 --prefill 1,8,32 --decode 1,2,4,8,16,32
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (path setup + model builder)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def timed(fn, warmup, iters):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="7b")
    ap.add_argument("--vit", default="1,16,64,256")
    ap.add_argument("--prefill", default="1,8,32")
    ap.add_argument("--decode", default="1,2,4,8,16,32")
    ap.add_argument("--prompt", type=int, default=128)
    ap.add_argument("--new", type=int, default=64)
    ap.add_argument("--max-images", type=int, default=32)
    ap.add_argument("--out", default="gpurun_out/config_sweep.jsonl")
    a = ap.parse_args()
    from llava import _b2
    from llava.model.llava_arch import build_source_index

    m = bench.MODELS[a.model]
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    lists = {k: [int(x) for x in getattr(a, k).split(",") if x] for k in ("vit", "prefill", "decode")}
    maxB = max(lists["prefill"] + lists["decode"] + [1])
    S, N = a.prompt + bench.P_IMG, a.new
    hbm_peak, tf_peak, _ = bench.peaks()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    fout = open(a.out, "a")

    def emit(d):
        d = dict(model=m["name"], **d)
        s = json.dumps(d)
        print(s, flush=True)
        fout.write(s + "\n")
        fout.flush()

    model = build(m, dev, maxB, S + N + 8, a.max_images)
    engine = model._ensure_engine()
    stream = torch.cuda.Stream(device=dev)
    g = torch.Generator().manual_seed(1)
    with torch.cuda.stream(stream), torch.no_grad():
        # ---- configs[2]: encode_images only -------------------------------------------------------------
        for B in lists["vit"]:
            pixels = torch.randn(B, 3, 336, 336, generator=g).to(dev, torch.bfloat16)
            for what, fn, gf in (("vit_encode", lambda: engine.vit_encode(pixels), bench.VIT_GF),
                                 ("encode_images", lambda: engine.encode_images(pixels),
                                  bench.VIT_GF + 2e-9 * bench.P_IMG * (1024 * m["hidden"] + m["hidden"] ** 2))):
                ms = timed(fn, 2, 3 if B >= 64 else 10)
                tf = B * gf / ms  # GFLOP / ms = TFLOP/s
                emit(dict(what=what, B=B, ms=ms, images_per_s=B / ms * 1e3, tflops=tf, frac_bf16_peak=tf / tf_peak))
            del pixels
        # ---- prefill + decode ladder -----------------------------------------------------------------------
        for B in sorted(set(lists["prefill"] + lists["decode"])):
            work = bench.algorithmic_work(m, B, S, N + 1)
            ids = torch.randint(3, bench.VOCAB, (B, a.prompt + 1), generator=g)
            ids[:, 0] = 1
            ids[:, 5] = bench.IMAGE_TOKEN
            ids_np = ids.numpy().astype(np.int64)
            src, _, _, _, lens = build_source_index(ids_np, np.ones_like(ids_np, bool), np.full_like(ids_np, -100),
                                                    B * bench.P_IMG, [bench.P_IMG] * B, None, "right")
            src_dev = torch.from_numpy(src.reshape(-1)).to(dev)
            feats = (torch.randn(B * bench.P_IMG, m["hidden"], generator=g) * 0.5).to(dev, torch.bfloat16)
            embeds = engine.splice(src_dev, feats, B, S)
            kv = engine.new_kv(B, S + N + 8)
            state = {}

            def prefill():
                kv.reset()
                state["first"] = engine.argmax(engine.prefill(kv, embeds, lens, _b2.LOGITS_LAST))

            if B in lists["prefill"]:
                ms = timed(prefill, 2, 5)
                tf = work["prefill_flops"] / ms / 1e9
                emit(dict(what="prefill", B=B, S=S, ms=ms, tok_per_s=B * S / ms * 1e3, tflops=tf, frac_bf16_peak=tf / tf_peak))
            if B in lists["decode"]:
                out_tokens = torch.empty(N, B, dtype=torch.int32, device=dev)
                best = None
                for _ in range(3):
                    prefill()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    engine.decode_greedy(kv, state["first"], N, out=out_tokens)
                    e1.record()
                    torch.cuda.synchronize()
                    t = e0.elapsed_time(e1) / N
                    best = t if best is None else min(best, t)
                gbs = work["decode_bytes_per_step"] / best / 1e6
                emit(dict(what="decode", B=B, ctx=f"{S}->{S + N}", ms_per_step=best, tok_per_s=B / best * 1e3,
                          algorithmic_gb_per_step=work["decode_bytes_per_step"] / 1e9, achieved_gbs=gbs,
                          frac_hbm_peak=gbs / hbm_peak))
            kv.close()
            del kv, embeds, feats


def build(m, device, max_batch, max_seq, max_images):
    from helpers import write_clip_config_dir, make_llava_config
    from llava.model import LlavaLlamaForCausalLM
    from oracle.llava_oracle import make_config, weight_shapes, init_std  # shapes/init table only (no compute)

    cfg = make_config(hidden=m["hidden"], inter=m["inter"], layers=m["layers"], heads=m["heads"])
    clip_dir = write_clip_config_dir(cfg)
    model = LlavaLlamaForCausalLM(make_llava_config(cfg, clip_dir), device=device, max_batch=max_batch,
                                  max_seq=max_seq, max_images=max_images)
    model.get_vision_tower().load_model(random_init=True)
    model.to(device=device, dtype=torch.bfloat16)
    gen = torch.Generator(device=device).manual_seed(0)
    kinds = {k: (shape, kind) for k, shape, kind in weight_shapes(cfg)}
    with torch.no_grad():
        for k, p in model.state_dict().items():
            shape, kind = kinds[k]
            p.normal_(0.0, init_std(kind, shape), generator=gen)
            if kind == "g":
                p.add_(1.0)
    model.invalidate_engine()
    return model.eval()


if __name__ == "__main__":
    main()
