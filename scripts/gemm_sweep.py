"""GPU microbenchmark of the tcgen05 GEMM over the LLaVA shapes and tile-N choices (CUDA events, L2-cold inputs
rotated over several weight copies). Run on the GPU box: python scripts/gemm_sweep.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "llava-plus-codebase_b200"))
from llava import _b2  # noqa: E402

_b2.init(0)
lib = _b2.load_library()
dev = "cuda"
SHAPES = [  # (name, M, N, K, act)
    ("7B prefill qkv", 704, 12288, 4096, 0), ("7B prefill o", 704, 4096, 4096, 0),
    ("7B prefill gate/up swiglu", 704, 22016, 4096, 3), ("7B prefill down", 704, 4096, 11008, 0),
    ("ViT b1 qkv", 577, 3072, 1024, 0), ("ViT b1 fc1", 577, 4096, 1024, 1), ("ViT b1 fc2", 577, 1024, 4096, 0),
    ("ViT b16 qkv", 9232, 3072, 1024, 0), ("ViT b16 fc1", 9232, 4096, 1024, 1), ("ViT b16 fc2", 9232, 1024, 4096, 0),
    ("ViT b16 out", 9232, 1024, 1024, 0), ("square 8192", 8192, 8192, 8192, 0),
    ("13B prefill gate/up", 704, 27648, 5120, 3), ("13B prefill o", 704, 5120, 5120, 0),
    ("projector fc1 b1", 576, 4096, 1024, 2), ("projector fc2 b1", 576, 4096, 4096, 0),
    ("ViT b64 qkv", 36928, 3072, 1024, 0), ("ViT b64 fc1", 36928, 4096, 1024, 1), ("ViT b64 fc2", 36928, 1024, 4096, 0),
    ("7B prefill b8 qkv", 5632, 12288, 4096, 0), ("7B prefill b8 gate/up swiglu", 5632, 22016, 4096, 3),
    ("7B prefill b8 down", 5632, 4096, 11008, 0),
]
res = []
for name, M, N, K, act in SHAPES:
    ncopy = max(1, min(8, int(300e6 // (N * K * 2)) + 1))
    A = (torch.randn(M, K, device=dev)).to(torch.bfloat16)
    Ws = [(torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16) for _ in range(ncopy)]
    bias = torch.zeros(N, device=dev, dtype=torch.bfloat16) if act == 1 else None
    out = torch.empty(M, N // 2 if act == 3 else N, device=dev, dtype=torch.bfloat16)
    row = {"shape": name, "M": M, "N": N, "K": K}
    for bn in (64, 128, 192, 256, 2):  # 2 = CTA-pair kernel (cta_group::2, 256x256 pair tiles)
        if act == 3 and bn % 128 != 0 and bn != 2:
            continue
        def run(i):
            W = Ws[i % ncopy]
            _b2.check(lib.b2_op_gemm(_b2.ptr(A), K, _b2.ptr(W), K, _b2.ptr(bias), None, 0, _b2.ptr(out), out.stride(0), 0,
                                     M, N, K, act, bn, _b2.stream_ptr()))
        for i in range(3):
            run(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        e0.record()
        for i in range(iters):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        row[f"bn{bn}_us"] = round(ms * 1e3, 1)
        row[f"bn{bn}_tflops"] = round(2.0 * M * N * K / (ms * 1e-3) / 1e12, 1)
    # library reference for context (cuBLAS through torch) — never on our hot path
    Wt = Ws[0]
    for i in range(3):
        torch.matmul(A, Wt.t())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        torch.matmul(A, Ws[i % ncopy].t())
    e1.record()
    torch.cuda.synchronize()
    row["cublas_tflops"] = round(2.0 * M * N * K / (e0.elapsed_time(e1) / 20 * 1e-3) / 1e12, 1)
    res.append(row)
    print(json.dumps(row), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "gemm_sweep.json"), "w"), indent=1)
