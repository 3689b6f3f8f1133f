"""GPU: the fp8 decode path (BASELINE configs[4]) against oracle/fp8_oracle.py, through the C ABI (first executed on a
B200 in round 2: profiles/r2a_fp8_2cta_first_run.txt)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from llava import _b2  # noqa: E402
from oracle import fp8_oracle as F  # noqa: E402
from oracle import llava_oracle as O  # noqa: E402
from helpers import make_engine, rel_err  # noqa: E402

DEV, BF = "cuda", torch.bfloat16
P, S = _b2.ptr, _b2.stream_ptr


@pytest.fixture(scope="module", autouse=True)
def _init():
    _b2.init(0)
    torch.manual_seed(0)


def quantize(x):
    rows, K = x.shape
    q = torch.empty(rows, K, device=DEV, dtype=torch.uint8)
    s = torch.empty(rows, device=DEV, dtype=torch.float32)
    _b2.check(_b2.load_library().b2_op_quantize_rows_e4m3(P(x), x.stride(0), rows, K, P(q), K, P(s), S()), "quantize")
    return q, s


@pytest.mark.parametrize("rows,K", [(1, 4096), (32, 4096), (7, 11008), (128, 5120), (300, 264)])
def test_quantize_rows_bit_exact(rows, K):
    x = (torch.randn(rows, K, device=DEV) * 3).to(BF)
    x[0, :8] = 0
    if rows > 2:
        x[2] = 0
    q, s = quantize(x)
    qr, sr = F.quantize_rows_e4m3(x.cpu())
    assert torch.equal(s.cpu(), sr)
    qg, qc = q.cpu(), qr.view(torch.uint8)
    bad = (qg != qc).nonzero()
    if len(bad):  # say WHAT differs: value, scaled value, both codes
        inv = torch.where(sr > 0, 448.0 / (sr * 448.0), torch.ones_like(sr))
        lines = [f"x={float(x[r, c]):.9g} scaled={float(x[r, c].float().cpu() * (448.0 / x[r].float().abs().max().cpu())):.9g} "
                 f"gpu={int(qg[r, c])} cpu={int(qc[r, c])}" for r, c in bad[:12].tolist()]
        pytest.fail(f"{len(bad)} of {qg.numel()} e4m3 codes differ:\n" + "\n".join(lines))


def test_rmsnorm_quant_matches_restatement():
    x = (torch.randn(32, 4096, device=DEV) * 2).to(BF)
    gamma = (1 + 0.1 * torch.randn(4096, device=DEV)).to(BF)
    q = torch.empty(32, 4096, device=DEV, dtype=torch.uint8)
    s = torch.empty(32, device=DEV, dtype=torch.float32)
    _b2.check(_b2.load_library().b2_op_rmsnorm_quant_e4m3(P(x), P(gamma), P(q), P(s), 32, 4096, 1e-5, S()), "rmsnorm_quant")
    y = F.rmsnorm_hf(x.cpu(), gamma.cpu(), 1e-5)
    qr, sr = F.quantize_rows_e4m3(y)
    # rstd differs in the last fp32 bit with the summation order: a few bf16 roundings (and so a few e4m3 codes) may flip
    torch.testing.assert_close(s.cpu(), sr, rtol=2e-2, atol=0)
    deq, deq_r = q.cpu().view(torch.float8_e4m3fn).float() * s.cpu()[:, None], qr.float() * sr[:, None]
    assert float(((deq - deq_r).abs() > 0.13 * deq_r.abs() + 1e-6).float().mean()) < 2e-3


def skinny_fp8(qx, sx, qw, sw, residual=None, act=_b2.ACT_NONE, out_fp32=False):
    B, K = qx.shape
    N = qw.shape[0]
    lib = _b2.load_library()
    n_out = N // 2 if act == _b2.ACT_SWIGLU else N
    out = torch.empty(B, n_out, device=DEV, dtype=torch.float32 if out_fp32 else BF)
    ws = torch.empty(int(lib.b2_op_gemm_skinny_workspace_bytes(B, N, K)) // 4, device=DEV, dtype=torch.float32)
    cnt = torch.zeros(int(lib.b2_op_gemm_skinny_counter_bytes(N)) // 4, device=DEV, dtype=torch.int32)
    _b2.check(lib.b2_op_gemm_skinny_fp8(P(qx), K, P(sx), P(qw), K, P(sw), P(residual),
                                        residual.stride(0) if residual is not None else 0, P(out), out.stride(0),
                                        int(out_fp32), B, N, K, act, P(ws), ws.numel() * 4, P(cnt), S()), "skinny_fp8")
    assert int(cnt.abs().sum()) == 0
    return out


@pytest.mark.parametrize("B,N,K", [(32, 128, 128), (32, 4096, 4096), (8, 12288, 4096), (64, 4096, 11008),
                                   (100, 5120, 13824), (20, 200, 272)])
def test_gemm_skinny_fp8_vs_restatement(B, N, K):
    x, w = (torch.randn(B, K, device=DEV)).to(BF), (torch.randn(N, K, device=DEV) * K ** -0.5).to(BF)
    qx, sx = quantize(x)
    qw, sw = quantize(w)
    want = F.linear_w8a8(qx.cpu().view(torch.float8_e4m3fn), sx.cpu(), qw.cpu().view(torch.float8_e4m3fn), sw.cpu())
    got = skinny_fp8(qx, sx, qw, sw, out_fp32=True)
    torch.testing.assert_close(got.cpu(), want, rtol=2e-3, atol=2e-3 * float(want.abs().mean()))
    # and the quantised product stays within the derived W8A8 error of the exact bf16 product
    ref = x.float() @ w.float().t()
    rel = float((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt())
    assert rel < 1.25 * F.expected_relative_error(K), rel


def test_gemm_skinny_fp8_swiglu_and_residual():
    B, I, h = 32, 1024, 512
    x = torch.randn(B, h, device=DEV).to(BF)
    wg, wu = (torch.randn(I, h, device=DEV) * h ** -0.5).to(BF), (torch.randn(I, h, device=DEV) * h ** -0.5).to(BF)
    wgu = torch.empty(2 * I, h, device=DEV, dtype=BF)
    _b2.check(_b2.load_library().b2_op_interleave_gate_up(P(wg), P(wu), P(wgu), I, h, S()))
    qx, sx = quantize(x)
    qw, sw = quantize(wgu)
    got = skinny_fp8(qx, sx, qw, sw, act=_b2.ACT_SWIGLU)
    want = F.swiglu_w8a8(x.cpu(), wg.cpu(), wu.cpu())
    # Error model: gate and up each carry an accumulation error delta ~ 2e-3 of their RMS (the plain-GEMM test's bound: the
    # tensor core's fp32 accumulation of e4m3 products is not bit-identical to a sequential fp32 sum), which
    # silu(g) * u turns into |silu'(g)| |u| delta + |silu(g)| delta — NOT small relative to the output where u or silu(g)
    # is small while the other factor is large — plus one bf16 rounding of the result.
    g_ref, u_ref = F.linear_fake_quant(x.cpu(), wg.cpu()), F.linear_fake_quant(x.cpu(), wu.cpu())
    delta = 2e-3 * float(g_ref.pow(2).mean().sqrt())
    tol = delta * (1.1 * u_ref.abs() + torch.nn.functional.silu(g_ref).abs()) + 2 ** -7 * want.float().abs() + 1e-6
    err = (got.cpu().float() - want.float()).abs()
    assert bool((err <= tol).all()), f"max err/tol {float((err / tol).max()):.2f} at {int((err / tol).argmax())}"
    r = torch.randn(B, 2 * I, device=DEV).to(BF)
    got_r = skinny_fp8(qx, sx, qw, sw, residual=r, out_fp32=True)
    want_r = F.linear_w8a8(qx.cpu().view(torch.float8_e4m3fn), sx.cpu(), qw.cpu().view(torch.float8_e4m3fn), sw.cpu(), r.cpu())
    torch.testing.assert_close(got_r.cpu(), want_r, rtol=2e-3, atol=2e-3)


def test_engine_fp8_decode_close_to_bf16_decode():
    """7B layer shapes, 2 layers, B=16: decode logits with e4m3 weights/activations vs the bf16 engine on the same cache.
    Tolerance from tests/test_fp8_oracle.py (one layer + head moves logits by < 0.35 max / 0.08 mean of std)."""
    cfg = O.make_config(hidden=4096, inter=11008, layers=2, heads=32, vit_layers=2)
    w = O.make_weights(cfg, seed=13)
    g = torch.Generator().manual_seed(6)
    B, S_ = 16, 64
    embeds = (torch.randn(B, S_, 4096, generator=g) * 0.5).to(BF)
    logits = {}
    for mode in ("bf16", "fp8"):
        eng = make_engine(cfg, w, max_batch=B, max_seq=128, max_images=1)
        if mode == "fp8":
            eng.enable_fp8_decode()
        kv = eng.new_kv(B, 128)
        last = eng.prefill(kv, embeds.to(DEV), None, _b2.LOGITS_LAST)   # prefill is bf16 in both
        toks = last.argmax(-1).to(torch.int32)
        logits[mode] = [eng.decode_step(kv, toks).cpu() for _ in range(1)][0]
        kv.close()
        eng.close()
    mx, mn = rel_err(logits["fp8"], logits["bf16"])
    print(f"fp8 vs bf16 decode logits: max {mx:.4f} mean {mn:.4f} of std")
    assert mx < 0.6 and mn < 0.12, (mx, mn)
