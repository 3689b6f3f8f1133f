"""CPU: the fp8 (e4m3 W8A8) restatement that defines BASELINE configs[4]'s arithmetic — quantiser invariants, the error
model the GPU tolerance is taken from, and the end-to-end effect on a decode step of the small oracle config."""
import torch

from oracle import fp8_oracle as F
from oracle import llava_oracle as O


def test_quantiser_invariants():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(37, 512, generator=g).to(torch.bfloat16)
    x[3] = 0                      # all-zero row
    x[5, 7] = 3.0e4               # outlier row: everything else collapses towards zero
    q, s = F.quantize_rows_e4m3(x)
    assert q.dtype == torch.float8_e4m3fn and s.dtype == torch.float32 and s.shape == (37,)
    assert not torch.isnan(q.float()).any(), "the cast must never hit the e4m3fn NaN encoding"
    assert float(s[3]) == 1.0 and float(q[3].float().abs().max()) == 0.0
    amax = x.float().abs().amax(-1)
    nz = amax > 0
    assert torch.equal(q.float().abs().amax(-1)[nz], torch.full_like(amax[nz], 448.0)), "the row maximum maps to +-448"
    deq = q.float() * s[:, None]
    # round-to-nearest on a 3-bit mantissa: |err| <= 2^-4 * |x| for normal values; tiny values (below 2^-6 * amax/448 * ...)
    # fall into the subnormal range where the absolute step is 2^-9 * scale
    err = (deq - x.float()).abs()
    bound = torch.maximum(x.float().abs() * 2.0 ** -4, (s * 2.0 ** -10)[:, None])
    assert bool((err <= bound * 1.0001).all())
    # idempotent: quantising the dequantised row reproduces the same codes
    q2, s2 = F.quantize_rows_e4m3(deq)
    assert torch.equal(q2.view(torch.uint8), q.view(torch.uint8)) and torch.allclose(s2, s)


def test_linear_error_model_matches_measurement():
    g = torch.Generator().manual_seed(1)
    for K in (256, 4096, 11008):
        x = torch.randn(16, K, generator=g).to(torch.bfloat16)
        w = (torch.randn(96, K, generator=g) * K ** -0.5).to(torch.bfloat16)
        ref = x.float() @ w.float().t()
        got = F.linear_fake_quant(x, w)
        rel = float((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt())
        model = F.expected_relative_error(K)
        assert 0.25 * model < rel < 1.25 * model, (K, rel, model)


def test_per_channel_scales_follow_interleaved_gate_up_rows():
    """The engine quantises the block-64 interleaved gate/up matrix row by row: the scale vector must follow the
    PHYSICAL rows, i.e. interleave(quantise(gate), quantise(up)) == quantise(interleave(gate, up))."""
    g = torch.Generator().manual_seed(2)
    I, h = 256, 128
    wg, wu = torch.randn(I, h, generator=g).to(torch.bfloat16), torch.randn(I, h, generator=g).to(torch.bfloat16)
    inter = torch.stack([wg.view(I // 64, 64, h), wu.view(I // 64, 64, h)], 1).reshape(2 * I, h)
    q, s = F.quantize_rows_e4m3(inter)
    qg, sg = F.quantize_rows_e4m3(wg)
    qu, su = F.quantize_rows_e4m3(wu)
    s_ref = torch.stack([sg.view(-1, 64), su.view(-1, 64)], 1).reshape(-1)
    q_ref = torch.stack([qg.view(torch.uint8).view(I // 64, 64, h), qu.view(torch.uint8).view(I // 64, 64, h)], 1).reshape(2 * I, h)
    assert torch.equal(s, s_ref) and torch.equal(q.view(torch.uint8), q_ref)


def test_decode_step_logit_shift_is_bounded():
    """One decoder layer + lm_head of the small oracle config with every Linear through W8A8: the logits move by a few
    per cent of their standard deviation — the tolerance the GPU test of the fp8 path states (max 0.35, mean 0.08 of std)."""
    cfg = O.CONFIGS["small"] if "small" in O.CONFIGS else O.CONFIGS["tiny"]
    w = O.make_weights(cfg, seed=3)
    h = cfg["hidden"]
    g = torch.Generator().manual_seed(4)
    x = (torch.randn(8, h, generator=g) * 0.5).to(torch.bfloat16)
    pre = "model.layers.0."
    xn = F.rmsnorm_hf(x, w[pre + "post_attention_layernorm.weight"], cfg["rms_eps"])
    wg, wu, wd = (w[pre + f"mlp.{n}_proj.weight"].to(torch.bfloat16) for n in ("gate", "up", "down"))
    ref_act = (torch.nn.functional.silu(xn.float() @ wg.float().t()) * (xn.float() @ wu.float().t())).to(torch.bfloat16)
    ref = x.float() + ref_act.float() @ wd.float().t()
    act = F.swiglu_w8a8(xn, wg, wu)
    got = x.float() + F.linear_fake_quant(act, wd)
    head = w["lm_head.weight"].to(torch.bfloat16)
    lr = F.rmsnorm_hf(ref.to(torch.bfloat16), w["model.norm.weight"], cfg["rms_eps"]).float() @ head.float().t()
    lg = F.linear_fake_quant(F.rmsnorm_hf(got.to(torch.bfloat16), w["model.norm.weight"], cfg["rms_eps"]), head)
    d = (lg - lr).abs() / lr.std()
    assert float(d.max()) < 0.35 and float(d.mean()) < 0.08, (float(d.max()), float(d.mean()))
