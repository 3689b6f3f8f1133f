"""Per-kernel timings of the batched decode step (no model build): every Linear of a 7B/13B layer at batch B through
the swap-AB stream-K GEMM (gemm_skinny.cu) and through the tile GEMM (gemm_tcgen05.cu, batch as M), and the split-KV
decode attention at that batch — each over operands larger than L2 (weights rotate through copies), CUDA events on
the launching stream. Prints achieved GB/s of the ALGORITHMIC bytes (weights once, K/V once).

    python scripts/op_bench.py --batch 32 --model 7b
"""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "llava-plus-codebase_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from llava import _b2  # noqa: E402

DEV, BF = "cuda", torch.bfloat16
P, S = _b2.ptr, _b2.stream_ptr


def timed(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="7b")
    ap.add_argument("--batch", default="32")
    ap.add_argument("--ctx", type=int, default=736)
    ap.add_argument("--out", default="gpurun_out/op_bench.jsonl")
    a = ap.parse_args()
    _b2.init(0)
    lib = _b2.load_library()
    h, I, H = (4096, 11008, 32) if a.model == "7b" else (5120, 13824, 40)
    V = 32000
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    fout = open(a.out, "a")

    def emit(d):
        s = json.dumps(d)
        print(s, flush=True)
        fout.write(s + "\n")
        fout.flush()

    for B in [int(x) for x in a.batch.split(",")]:
        for name, N, K, act in (("qkv", 3 * h, h, 0), ("o_proj", h, h, 0), ("gate_up", 2 * I, h, 3), ("down", h, I, 0),
                                ("lm_head", V, h, 0)):
            copies = max(2, math.ceil(300e6 / (N * K * 2)))
            Ws = [(torch.randn(N, K, device=DEV) * K ** -0.5).to(BF) for _ in range(copies)]
            x = torch.randn(B, K, device=DEV).to(BF)
            n_out = N // 2 if act == 3 else N
            out = torch.empty(B, n_out, device=DEV, dtype=BF)
            ws = torch.empty(int(lib.b2_op_gemm_skinny_workspace_bytes(B, N, K)) // 4, device=DEV, dtype=torch.float32)
            cnt = torch.zeros(int(lib.b2_op_gemm_skinny_counter_bytes(N)) // 4, device=DEV, dtype=torch.int32)
            state = {"i": 0}

            def run_skinny():
                W = Ws[state["i"] % copies]
                state["i"] += 1
                _b2.check(lib.b2_op_gemm_skinny(P(x), K, P(W), K, None, 0, P(out), n_out, 0, B, N, K, act, P(ws),
                                                ws.numel() * 4, P(cnt), S()), "skinny")

            def run_tile():
                W = Ws[state["i"] % copies]
                state["i"] += 1
                _b2.check(lib.b2_op_gemm(P(x), K, P(W), K, None, None, 0, P(out), n_out, 0, B, N, K, act, 0, S()), "gemm")

            # parity spot check against fp32 torch on copy 0
            state["i"] = 0
            run_skinny()
            ref = x.float() @ Ws[0].float().t()
            if act == 3:
                v = ref.view(B, N // 128, 2, 64)
                ref = (torch.nn.functional.silu(v[:, :, 0]) * v[:, :, 1]).reshape(B, N // 2)
            err = float((out.float() - ref).abs().max() / (ref.abs().mean() + 1e-6))
            us_s = timed(run_skinny, 4 * copies)
            us_t = timed(run_tile, 4 * copies)
            gb = N * K * 2 / 1e9
            emit(dict(op=name, B=B, N=N, K=K, skinny_us=us_s, tile_gemm_us=us_t, skinny_gbs=gb / us_s * 1e6,
                      tile_gemm_gbs=gb / us_t * 1e6, skinny_max_err_over_mean=err))
            del Ws, x, out, ws, cnt
            torch.cuda.empty_cache()
        # ---- decode attention at this batch -----------------------------------------------------------------
        Smax = a.ctx + 40
        hd = H * 128
        qkv = torch.randn(B, 3 * hd, device=DEV).to(BF)
        kc = torch.randn(B, H, Smax, 128, device=DEV).to(BF)
        vc = torch.randn(B, H, Smax, 128, device=DEV).to(BF)
        cur = torch.full((B,), a.ctx, device=DEV, dtype=torch.int32)
        out = torch.empty(B, hd, device=DEV, dtype=BF)
        for nsplit in sorted({max(1, min(32, (4 * 148 + B * H - 1) // (B * H))), 1, 2, 4}):
            scratch = torch.zeros(int(lib.b2_op_decode_attn_scratch_bytes(B, H, nsplit)), device=DEV, dtype=torch.uint8)

            def run_attn():
                _b2.check(lib.b2_op_decode_attn(P(qkv), P(kc), P(vc), P(cur), P(out), P(scratch), B, H, Smax, nsplit,
                                                10000.0, 1 / math.sqrt(128), S()), "decode_attn")

            us = timed(run_attn, 10)
            gb = B * H * (a.ctx + 1) * 128 * 2 * 2 / 1e9
            emit(dict(op="decode_attn", B=B, H=H, ctx=a.ctx, nsplit=nsplit, us=us, gbs=gb / us * 1e6))


if __name__ == "__main__":
    main()
