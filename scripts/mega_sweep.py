"""Decode-megakernel knob sweep inside ONE process (7B, bs=1, ctx 704 -> 704+N): the knobs are environment variables
that model.cu re-reads on every launch. Every setting must generate exactly the tokens of the baseline setting.

    python scripts/mega_sweep.py [--new 128] [--configs "ahead,mode,fast;..."]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import bench  # noqa: E402
import config_sweep  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="7b")
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--new", type=int, default=128)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--configs", default="0,1,0;0,1,1;4,1,0;8,1,0;12,1,0;16,1,0;8,2,0;8,1,1;12,1,1")
    ap.add_argument("--batches", default="", help="after the knob sweep: decode at these batch sizes with the skinny GEMM on/off")
    ap.add_argument("--out", default="gpurun_out/mega_sweep.jsonl")
    a = ap.parse_args()
    from llava import _b2
    from llava.model.llava_arch import build_source_index

    m = bench.MODELS[a.model]
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    B, N, S = a.batch, a.new, 128 + bench.P_IMG
    hbm_peak, _, _ = bench.peaks()
    batches = [int(x) for x in a.batches.split(",") if x]
    model = config_sweep.build(m, dev, max([B] + batches), S + N + 8, 8)
    engine = model._ensure_engine()
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(3, bench.VOCAB, (B, 129), generator=g)
    ids[:, 0] = 1
    ids[:, 5] = bench.IMAGE_TOKEN
    ids_np = ids.numpy().astype(np.int64)
    src, _, _, _, lens = build_source_index(ids_np, np.ones_like(ids_np, bool), np.full_like(ids_np, -100),
                                            B * bench.P_IMG, [bench.P_IMG] * B, None, "right")
    work = bench.algorithmic_work(m, B, S, N + 1)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    fout = open(a.out, "a")
    stream = torch.cuda.Stream(device=dev)
    base_tokens = None
    with torch.cuda.stream(stream), torch.no_grad():
        src_dev = torch.from_numpy(src.reshape(-1)).to(dev)
        feats = (torch.randn(B * bench.P_IMG, m["hidden"], generator=g) * 0.5).to(dev, torch.bfloat16)
        embeds = engine.splice(src_dev, feats, B, S)
        kv = engine.new_kv(B, S + N + 8)
        out_tokens = torch.empty(N, B, dtype=torch.int32, device=dev)
        for cfg in a.configs.split(";"):
            fields = [int(x) for x in cfg.split(",")] + [0]
            ahead, mode, fast, gsm = fields[:4]
            os.environ["B2_MEGA_L2_AHEAD"] = str(ahead)
            os.environ["B2_MEGA_L2_MODE"] = str(mode)
            os.environ["B2_MEGA_FAST_PROLOGUE"] = str(fast)
            os.environ["B2_MEGA_GAMMA_SMEM"] = str(gsm)
            times = []
            for _ in range(a.reps + 1):
                kv.reset()
                first = engine.argmax(engine.prefill(kv, embeds, lens, _b2.LOGITS_LAST))
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                engine.decode_greedy(kv, first, N, out=out_tokens)
                e1.record()
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1) / N)
            toks = out_tokens.cpu()
            if base_tokens is None:
                base_tokens = toks.clone()
            same = bool((toks == base_tokens).all())
            best = min(times[1:])
            gbs = work["decode_bytes_per_step"] / best / 1e6
            d = dict(l2_ahead=ahead, l2_mode=mode, fast_prologue=fast, gamma_smem=gsm, ms_per_token=best, all_ms=times,
                     achieved_gbs=gbs, frac_hbm_peak=gbs / hbm_peak, tokens_equal_baseline=same)
            s = json.dumps(d)
            print(s, flush=True)
            fout.write(s + "\n")
            fout.flush()
        kv.close()
        # ---- batch ladder: multi-kernel graph decode, skinny (swap-AB stream-K) GEMM on / off --------------------
        for Bb in batches:
            workb = bench.algorithmic_work(m, Bb, S, N + 1)
            idsb = torch.randint(3, bench.VOCAB, (Bb, 129), generator=g)
            idsb[:, 0] = 1
            idsb[:, 5] = bench.IMAGE_TOKEN
            idsb_np = idsb.numpy().astype(np.int64)
            srcb, _, _, _, lensb = build_source_index(idsb_np, np.ones_like(idsb_np, bool), np.full_like(idsb_np, -100),
                                                      Bb * bench.P_IMG, [bench.P_IMG] * Bb, None, "right")
            featsb = (torch.randn(Bb * bench.P_IMG, m["hidden"], generator=g) * 0.5).to(dev, torch.bfloat16)
            embedsb = engine.splice(torch.from_numpy(srcb.reshape(-1)).to(dev), featsb, Bb, S)
            outb = torch.empty(N, Bb, dtype=torch.int32, device=dev)
            ref = None
            for sk in ("0", "1"):
                os.environ["B2_DECODE_SKINNY"] = sk
                kvb = engine.new_kv(Bb, S + N + 8)  # a fresh cache -> a freshly captured decode graph
                times = []
                for _ in range(3):
                    kvb.reset()
                    first = engine.argmax(engine.prefill(kvb, embedsb, lensb, _b2.LOGITS_LAST))
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    engine.decode_greedy(kvb, first, N, out=outb)
                    e1.record()
                    torch.cuda.synchronize()
                    times.append(e0.elapsed_time(e1) / N)
                toks = outb.cpu()
                if ref is None:
                    ref = toks.clone()
                best = min(times[1:])
                gbs = workb["decode_bytes_per_step"] / best / 1e6
                d = dict(what="decode_batch", B=Bb, skinny=int(sk), ms_per_step=best, tok_per_s=Bb / best * 1e3,
                         achieved_gbs=gbs, frac_hbm_peak=gbs / hbm_peak,
                         token_agreement_vs_tile_gemm=float((toks == ref).float().mean()))
                s = json.dumps(d)
                print(s, flush=True)
                fout.write(s + "\n")
                fout.flush()
                kvb.close()


if __name__ == "__main__":
    main()

