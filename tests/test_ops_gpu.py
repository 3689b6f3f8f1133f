"""GPU: every CUDA kernel of the path, called through the C ABI (include/b2llava.h b2_op_*), against a plain
PyTorch fp32 reference of the same op on the same bf16-rounded inputs. Shapes include the LLaVA-1.5 full
sizes (BASELINE configs) and the ragged/edge cases (partial tiles, K tail, batch 1..8, ragged seq_lens)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from llava import _b2  # noqa: E402

DEV = "cuda"
BF = torch.bfloat16


@pytest.fixture(scope="module", autouse=True)
def _init():
    _b2.init(0)
    torch.manual_seed(0)


def P(t):
    return _b2.ptr(t)


def S():
    return _b2.stream_ptr()


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device=DEV) * scale).to(BF)


def gemm(A, W, bias=None, residual=None, act=_b2.ACT_NONE, out_fp32=False, bn=0):
    M, K = A.shape
    N = W.shape[0]
    n_out = N // 2 if act == _b2.ACT_SWIGLU else N
    out = torch.empty(M, n_out, device=DEV, dtype=torch.float32 if out_fp32 else BF)
    lib = _b2.load_library()
    _b2.check(lib.b2_op_gemm(P(A), A.stride(0), P(W), W.stride(0), P(bias), P(residual),
                             residual.stride(0) if residual is not None else 0, P(out), out.stride(0),
                             int(out_fp32), M, N, K, act, bn, S()), "b2_op_gemm")
    return out


def ref_linear(A, W, bias=None):
    y = A.float() @ W.float().t()
    return y + bias.float() if bias is not None else y


def assert_close(got, want, rtol=1.6e-2, atol=None):
    got, want = got.float(), want.float()
    atol = atol if atol is not None else 1.6e-2 * float(want.abs().mean() + 1e-6)
    torch.testing.assert_close(got, want, rtol=rtol, atol=atol)


# ---------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K,bn", [
    (128, 128, 64, 128),        # single tile, single k-block
    (128, 256, 512, 128),       # pipeline wrap (8 k-blocks > 6 stages)
    (704, 4096, 4096, 0),       # LLaVA-7B prefill o_proj (BASELINE config 2), heuristic tile
    (704, 12288, 4096, 128),    # fused QKV
    (577, 1024, 1024, 64),      # ViT B=1: ragged M (577 = 4*128 + 65)
    (577, 3072, 1024, 128),
    (576, 1024, 592, 64),       # patch embedding: K = 588 padded to 592 (K tail, not a multiple of 64)
    (1154, 4096, 1024, 256),    # BN=256 variant, 2 images
    (64, 32000, 4096, 128),     # decode-as-GEMM (B=64), fp32 logits shape
    (300, 136, 264, 64),        # everything ragged: M, N (not a tile multiple), K
    (704, 4096, 4096, 192),     # BN=192: 132 tiles fill one wave of 148 SMs (last n-tile is 64 wide)
    (577, 4096, 1024, 192),     # ViT fc1 at B=1
    (300, 136, 264, 192),       # ragged everything through the 192-wide tile (N < BN)
])
def test_gemm_plain(M, N, K, bn):
    A, W = rnd(M, K), rnd(N, K, scale=K ** -0.5)
    assert_close(gemm(A, W, bn=bn), ref_linear(A, W))


def test_gemm_persistent_many_tiles():
    # > 148 tiles per wave * several waves: exercises the TMEM double buffer and the tile scheduler wrap
    A, W = rnd(2048, 1024), rnd(4096, 1024, scale=1 / 32)
    assert_close(gemm(A, W, bn=128), ref_linear(A, W))
    assert_close(gemm(A, W, bn=64), ref_linear(A, W))
    assert_close(gemm(A, W, bn=256), ref_linear(A, W))
    assert_close(gemm(A, W, bn=192), ref_linear(A, W))


@pytest.mark.parametrize("act", [_b2.ACT_NONE, _b2.ACT_QUICK_GELU, _b2.ACT_GELU_ERF])
def test_gemm_bias_act_residual(act):
    M, N, K = 577, 4096, 1024
    A, W, b, r = rnd(M, K), rnd(N, K, scale=K ** -0.5), rnd(N), rnd(M, N)
    y = ref_linear(A, W, b)
    if act == _b2.ACT_QUICK_GELU:
        y = y * torch.sigmoid(1.702 * y)
    elif act == _b2.ACT_GELU_ERF:
        y = torch.nn.functional.gelu(y)
    assert_close(gemm(A, W, bias=b, act=act), y)
    assert_close(gemm(A, W, bias=b, residual=r, act=act), y + r.float())
    # in-place residual (out aliases residual), as the decoder layers use it
    r2 = r.clone()
    lib = _b2.load_library()
    _b2.check(lib.b2_op_gemm(P(A), K, P(W), K, P(b), P(r2), N, P(r2), N, 0, M, N, K, act, 0, S()))
    assert_close(r2, y + r.float())


def test_gemm_fp32_out():
    A, W = rnd(8, 4096), rnd(32000, 4096, scale=1 / 64)
    out = gemm(A, W, out_fp32=True)
    torch.testing.assert_close(out, ref_linear(A, W), rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("M,I,h,bn", [(704, 11008, 4096, 128), (100, 512, 256, 128), (300, 1024, 512, 256)])
def test_gemm_swiglu_interleaved(M, I, h, bn):
    x, Wg, Wu = rnd(M, h), rnd(I, h, scale=h ** -0.5), rnd(I, h, scale=h ** -0.5)
    Wgu = torch.empty(2 * I, h, device=DEV, dtype=BF)
    lib = _b2.load_library()
    _b2.check(lib.b2_op_interleave_gate_up(P(Wg), P(Wu), P(Wgu), I, h, S()))
    # layout contract: group g rows [0,64) = gate[g*64..], rows [64,128) = up[g*64..]
    v = Wgu.view(I // 64, 2, 64, h)
    assert torch.equal(v[:, 0].reshape(I, h), Wg) and torch.equal(v[:, 1].reshape(I, h), Wu)
    want = torch.nn.functional.silu(ref_linear(x, Wg)) * ref_linear(x, Wu)
    assert_close(gemm(x, Wgu, act=_b2.ACT_SWIGLU, bn=bn), want)


def test_gemm_linearity_and_determinism():
    """size-independent properties at full size: C(A1+A2) = C(A1)+C(A2) (fp32 out), bitwise repeatable."""
    A1, A2, W = rnd(704, 4096), rnd(704, 4096), rnd(4096, 4096, scale=1 / 64)
    c1, c2 = gemm(A1, W, out_fp32=True), gemm(A2, W, out_fp32=True)
    c12 = gemm((A1.float() + A2.float()).to(BF), W, out_fp32=True)
    ref12 = ref_linear((A1.float() + A2.float()).to(BF), W)
    torch.testing.assert_close(c12, ref12, rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(c1 + c2, ref_linear(A1, W) + ref_linear(A2, W), rtol=2e-3, atol=4e-3)
    assert torch.equal(gemm(A1, W, out_fp32=True), c1)


# ------------------------------------------------------------------------- CTA-pair (cta_group::2) GEMM
@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (256, 256, 512), (704, 4096, 4096), (9232, 3072, 1024),
                                   (9232, 1024, 4096), (5632, 12288, 4096), (300, 136, 264)])
def test_gemm_2cta_plain(M, N, K):
    A, W = rnd(M, K), rnd(N, K, scale=K ** -0.5)
    assert_close(gemm(A, W, bn=2), ref_linear(A, W))


def test_gemm_2cta_epilogues_match_the_1cta_kernel():
    M, N, K = 1154, 4096, 1024
    A, W, b, r = rnd(M, K), rnd(N, K, scale=K ** -0.5), rnd(N), rnd(M, N)
    for act in (_b2.ACT_NONE, _b2.ACT_QUICK_GELU, _b2.ACT_GELU_ERF):  # same operands, same fp32 accumulation: <= 1 bf16 ulp apart
        assert_close(gemm(A, W, bias=b, residual=r, act=act, bn=2), gemm(A, W, bias=b, residual=r, act=act, bn=256), rtol=8e-3)
    torch.testing.assert_close(gemm(A, W, out_fp32=True, bn=2), gemm(A, W, out_fp32=True, bn=256), rtol=1e-4, atol=1e-4)
    x, Wg, Wu = rnd(704, 4096), rnd(11008, 4096, scale=1 / 64), rnd(11008, 4096, scale=1 / 64)
    Wgu = torch.empty(2 * 11008, 4096, device=DEV, dtype=BF)
    _b2.check(_b2.load_library().b2_op_interleave_gate_up(P(Wg), P(Wu), P(Wgu), 11008, 4096, S()))
    assert_close(gemm(x, Wgu, act=_b2.ACT_SWIGLU, bn=2), gemm(x, Wgu, act=_b2.ACT_SWIGLU, bn=256), rtol=8e-3)


# ------------------------------------------------------------------------- skinny (swap-AB stream-K) GEMM
def skinny(x, W, residual=None, act=_b2.ACT_NONE, out_fp32=False, out=None, scratch=None):
    B, K = x.shape
    N = W.shape[0]
    lib = _b2.load_library()
    n_out = N // 2 if act == _b2.ACT_SWIGLU else N
    if out is None:
        out = torch.empty(B, n_out, device=DEV, dtype=torch.float32 if out_fp32 else BF)
    if scratch is None:
        ws = torch.empty(int(lib.b2_op_gemm_skinny_workspace_bytes(B, N, K)) // 4, device=DEV, dtype=torch.float32)
        cnt = torch.zeros(int(lib.b2_op_gemm_skinny_counter_bytes(N)) // 4, device=DEV, dtype=torch.int32)
    else:
        ws, cnt = scratch
    _b2.check(lib.b2_op_gemm_skinny(P(x), x.stride(0), P(W), W.stride(0), P(residual),
                                    residual.stride(0) if residual is not None else 0, P(out), out.stride(0),
                                    int(out_fp32), B, N, K, act, P(ws), ws.numel() * 4, P(cnt), S()), "b2_op_gemm_skinny")
    assert int(cnt.abs().sum()) == 0, "stream-K tile counters must be left at zero"
    return out


@pytest.mark.parametrize("B,N,K", [
    (32, 128, 64),          # one tile, one k-block, one CTA
    (32, 256, 1024),        # 2 tiles x 16 k-blocks over 32 CTAs: every tile split 16 ways
    (9, 4096, 4096),        # 7B o_proj: 32 tiles on 148 SMs (stream-K splits each tile ~4.6 ways)
    (32, 12288, 4096),      # 7B fused QKV
    (32, 4096, 11008),      # 7B down_proj (K = 172 k-blocks)
    (17, 5120, 13824),      # 13B down_proj, ragged batch
    (33, 4096, 4096),       # BN = 64 variant
    (64, 15360, 5120),      # 13B QKV at bs 64
    (100, 4096, 4096),      # BN = 128 variant, ragged batch
    (20, 200, 264),         # ragged N (not a tile multiple) and K tail (264 = 4*64 + 8)
])
def test_gemm_skinny_plain(B, N, K):
    x, W = rnd(B, K), rnd(N, K, scale=K ** -0.5)
    assert_close(skinny(x, W), ref_linear(x, W))


def test_gemm_skinny_residual_inplace_and_fp32_logits():
    B, N, K = 32, 4096, 11008
    x, W, r = rnd(B, K), rnd(N, K, scale=K ** -0.5), rnd(B, N)
    want = ref_linear(x, W) + r.float()
    assert_close(skinny(x, W, residual=r), want)
    r2 = r.clone()
    skinny(x, W, residual=r2, out=r2)  # out aliases residual, as the decoder layers use it
    assert_close(r2, want)
    x, W = rnd(32, 4096), rnd(32000, 4096, scale=1 / 64)  # lm_head: 250 tiles, fp32 logits
    torch.testing.assert_close(skinny(x, W, out_fp32=True), ref_linear(x, W), rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("B,I,h", [(32, 11008, 4096), (12, 512, 256), (64, 13824, 5120), (128, 1024, 512)])
def test_gemm_skinny_swiglu(B, I, h):
    x, Wg, Wu = rnd(B, h), rnd(I, h, scale=h ** -0.5), rnd(I, h, scale=h ** -0.5)
    Wgu = torch.empty(2 * I, h, device=DEV, dtype=BF)
    lib = _b2.load_library()
    _b2.check(lib.b2_op_interleave_gate_up(P(Wg), P(Wu), P(Wgu), I, h, S()))
    want = torch.nn.functional.silu(ref_linear(x, Wg)) * ref_linear(x, Wu)
    assert_close(skinny(x, Wgu, act=_b2.ACT_SWIGLU), want)


def test_gemm_skinny_matches_tile_gemm_and_is_deterministic():
    """Same operands through the prefill GEMM (batch as M) and the decode GEMM (batch as N): fp32 outputs agree to
    accumulation-order noise; the stream-K reduction is bitwise repeatable (fixed slot order) with reused scratch."""
    x, W = rnd(32, 4096), rnd(12288, 4096, scale=1 / 64)
    lib = _b2.load_library()
    ws = torch.empty(int(lib.b2_op_gemm_skinny_workspace_bytes(32, 12288, 4096)) // 4, device=DEV, dtype=torch.float32)
    cnt = torch.zeros(int(lib.b2_op_gemm_skinny_counter_bytes(12288)) // 4, device=DEV, dtype=torch.int32)
    a = skinny(x, W, out_fp32=True, scratch=(ws, cnt))
    for _ in range(3):
        assert torch.equal(skinny(x, W, out_fp32=True, scratch=(ws, cnt)), a)
    torch.testing.assert_close(a, gemm(x, W, out_fp32=True), rtol=1e-3, atol=1e-3)


def test_gemm_skinny_rejects_bad_arguments():
    lib = _b2.load_library()
    x, W = rnd(200, 64), rnd(128, 64)
    assert lib.b2_op_gemm_skinny_workspace_bytes(200, 128, 64) == -1
    ws, cnt = torch.empty(1 << 20, device=DEV), torch.zeros(64, device=DEV, dtype=torch.int32)
    with pytest.raises(ValueError):
        skinny(x, W, scratch=(ws, cnt))  # B > 128
    with pytest.raises(ValueError):
        skinny(rnd(16, 60), rnd(128, 60), scratch=(ws, cnt))  # K % 8 != 0
    with pytest.raises(ValueError):
        skinny(rnd(16, 64), rnd(128, 64), scratch=(ws[:16], cnt))  # workspace too small


def test_gemm_rejects_bad_arguments():
    A, W = rnd(16, 60), rnd(16, 60)
    with pytest.raises(ValueError):
        gemm(A, W)  # K % 8 != 0


# ---------------------------------------------------------------------------------------------- norms
@pytest.mark.parametrize("rows,cols", [(577, 1024), (1, 1024), (1154, 256)])
def test_layernorm(rows, cols):
    x, g, b = rnd(rows, cols), (1 + 0.1 * torch.randn(cols, device=DEV)).to(BF), rnd(cols, scale=0.1)
    y = torch.empty_like(x)
    _b2.check(_b2.load_library().b2_op_layernorm(P(x), P(g), P(b), P(y), rows, cols, 1e-5, S()))
    want = torch.nn.functional.layer_norm(x.float(), (cols,), g.float(), b.float(), 1e-5)
    assert_close(y, want, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("rows,cols", [(704, 4096), (3, 5120), (1, 256)])
def test_rmsnorm(rows, cols):
    x, g = rnd(rows, cols), (1 + 0.1 * torch.randn(cols, device=DEV)).to(BF)
    y = torch.empty_like(x)
    _b2.check(_b2.load_library().b2_op_rmsnorm(P(x), P(g), P(y), rows, cols, 1e-5, S()))
    xf = x.float()
    want = g.float() * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5)).to(BF).float()
    assert_close(y, want, rtol=1e-2, atol=1e-2)


# ---------------------------------------------------------------------------------------------- attention
def ref_attention(q, k, v, causal, lens=None):
    B, S_, H, D = q.shape
    qf, kf, vf = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))
    s = qf @ kf.transpose(-1, -2) / math.sqrt(D)
    mask = torch.zeros(B, 1, S_, S_, device=q.device, dtype=torch.bool)
    if causal:
        mask |= torch.triu(torch.ones(S_, S_, device=q.device, dtype=torch.bool), 1)
    if lens is not None:
        ar = torch.arange(S_, device=q.device)
        mask |= (ar[None, :] >= lens[:, None])[:, None, None, :]
    s = s.masked_fill(mask, float("-inf"))
    return (torch.softmax(s, -1) @ vf).permute(0, 2, 1, 3)


@pytest.mark.parametrize("B,S_,H,D,causal", [
    (1, 577, 16, 64, 0),     # CLIP ViT-L/14-336 (577 = 9*64 + 1 tokens)
    (3, 17, 4, 64, 0),       # tiny ViT, single partial tile
    (1, 704, 32, 128, 1),    # LLaVA-7B prefill, BASELINE config 2
    (2, 130, 2, 128, 1),
    (1, 64, 2, 128, 1),
])
def test_flash_attention(B, S_, H, D, causal):
    q, k, v = rnd(B, S_, H, D), rnd(B, S_, H, D), rnd(B, S_, H, D)
    o = torch.empty_like(q)
    _b2.check(_b2.load_library().b2_op_flash_attn(P(q), P(k), P(v), P(o), None, B, S_, H, D, causal,
                                                  1 / math.sqrt(D), S()))
    assert_close(o, ref_attention(q, k, v, causal), rtol=2e-2, atol=2e-2)


def test_flash_attention_ragged_lengths():
    B, S_, H, D = 3, 200, 4, 128
    q, k, v = rnd(B, S_, H, D), rnd(B, S_, H, D), rnd(B, S_, H, D)
    lens = torch.tensor([200, 1, 77], device=DEV, dtype=torch.int32)
    o = torch.empty_like(q)
    _b2.check(_b2.load_library().b2_op_flash_attn(P(q), P(k), P(v), P(o), P(lens), B, S_, H, D, 1,
                                                  1 / math.sqrt(D), S()))
    want = ref_attention(q, k, v, True, lens)
    for b in range(B):
        n = int(lens[b])
        assert_close(o[b, :n], want[b, :n], rtol=2e-2, atol=2e-2)
    assert torch.isfinite(o.float()).all()  # padded query rows stay finite


def ref_rope(x, pos, theta=10000.0):
    """HF apply_rotary_pos_emb in bf16 (cos/sin cast to bf16, bf16 products)."""
    D = x.shape[-1]
    inv = 1.0 / (theta ** (torch.arange(0, D, 2, device=x.device, dtype=torch.float32) / D))
    f = pos.float()[:, None] * inv[None]
    emb = torch.cat([f, f], -1)
    cos, sin = emb.cos().to(BF), emb.sin().to(BF)
    x1, x2 = x[..., : D // 2], x[..., D // 2:]
    rot = torch.cat([-x2, x1], -1)
    return x * cos[:, None, :] + rot * sin[:, None, :]


def test_rope_kv_write():
    B, S_, H, D, Smax = 2, 70, 4, 128, 96
    qkv = rnd(B * S_, 3 * H * D)
    orig = qkv.clone().view(B, S_, 3, H, D)
    kc = torch.zeros(B, H, Smax, D, device=DEV, dtype=BF)
    vc = torch.zeros_like(kc)
    _b2.check(_b2.load_library().b2_op_rope_kv_write(P(qkv), P(kc), P(vc), B, S_, H, D, Smax, 10000.0, S()))
    pos = torch.arange(S_, device=DEV)
    got = qkv.view(B, S_, 3, H, D)
    for b in range(B):
        torch.testing.assert_close(got[b, :, 0].float(), ref_rope(orig[b, :, 0], pos).float(), rtol=2e-2, atol=2e-2)
        torch.testing.assert_close(kc[b, :, :S_].permute(1, 0, 2).float(), ref_rope(orig[b, :, 1], pos).float(),
                                   rtol=2e-2, atol=2e-2)
        assert torch.equal(vc[b, :, :S_].permute(1, 0, 2), orig[b, :, 2])
    assert torch.equal(got[:, :, 2], orig[:, :, 2])  # v untouched in qkv


@pytest.mark.parametrize("B,H,lens,nsplit", [(1, 32, [703], 19), (4, 2, [0, 5, 64, 129], 4), (2, 4, [33, 7], 1),
                                             (8, 32, [1088] * 8, 3)])
def test_decode_attention(B, H, lens, nsplit):
    D, Smax = 128, max(lens) + 8
    hd = H * D
    qkv = rnd(B, 3 * hd)
    kc, vc = rnd(B, H, Smax, D), rnd(B, H, Smax, D)
    kc0, vc0 = kc.clone(), vc.clone()
    cur = torch.tensor(lens, device=DEV, dtype=torch.int32)
    lib = _b2.load_library()
    scratch = torch.zeros(lib.b2_op_decode_attn_scratch_bytes(B, H, nsplit), device=DEV, dtype=torch.uint8)
    out = torch.empty(B, hd, device=DEV, dtype=BF)
    for _ in range(2):  # second launch checks the self-resetting split counters
        kc.copy_(kc0), vc.copy_(vc0)
        _b2.check(lib.b2_op_decode_attn(P(qkv), P(kc), P(vc), P(cur), P(out), P(scratch), B, H, Smax, nsplit,
                                        10000.0, 1 / math.sqrt(D), S()))
    v3 = qkv.view(B, 3, H, D)
    for b in range(B):
        n = lens[b]
        pos = torch.tensor([n], device=DEV)
        q = ref_rope(v3[b, 0][None], pos)[0]            # [H, D]
        knew = ref_rope(v3[b, 1][None], pos)[0]
        K = torch.cat([kc0[b, :, :n], knew[:, None]], 1).float()   # [H, n+1, D]
        V = torch.cat([vc0[b, :, :n], v3[b, 2][:, None]], 1).float()
        s = torch.einsum("hd,hnd->hn", q.float(), K) / math.sqrt(D)
        want = torch.einsum("hn,hnd->hd", torch.softmax(s, -1), V).reshape(-1)
        assert_close(out[b], want, rtol=2e-2, atol=2e-2)
        # cache append: exactly row n written, everything else untouched
        torch.testing.assert_close(kc[b, :, n].float(), knew.float(), rtol=1e-2, atol=1e-2)
        assert torch.equal(vc[b, :, n], v3[b, 2])
        assert torch.equal(kc[b, :, :n], kc0[b, :, :n]) and torch.equal(kc[b, :, n + 1:], kc0[b, :, n + 1:])


# ---------------------------------------------------------------------------------------------- GEMV (decode)
def gemv(x, W, gamma=None, residual=None, act=_b2.ACT_NONE, out_fp32=False, eps=1e-5):
    B, K = x.shape
    N = W.shape[0]
    n_out = N // 2 if act == _b2.ACT_SWIGLU else N
    out = torch.empty(B, n_out, device=DEV, dtype=torch.float32 if out_fp32 else BF)
    _b2.check(_b2.load_library().b2_op_gemv(P(x), x.stride(0), P(W), K, P(gamma), eps, P(residual),
                                            residual.stride(0) if residual is not None else 0, P(out), n_out,
                                            int(out_fp32), B, N, K, act, S()), "b2_op_gemv")
    return out


@pytest.mark.parametrize("B", [1, 2, 3, 4, 5, 8])
@pytest.mark.parametrize("N,K", [(4096, 4096), (12288, 4096), (4096, 11008), (768, 256)])
def test_gemv_plain_and_residual(B, N, K):
    x, W, r = rnd(B, K), rnd(N, K, scale=K ** -0.5), rnd(B, N)
    assert_close(gemv(x, W), ref_linear(x, W))
    assert_close(gemv(x, W, residual=r), ref_linear(x, W) + r.float())


@pytest.mark.parametrize("B", [1, 4, 8])
def test_gemv_fused_rmsnorm_and_fp32_logits(B):
    K, N = 4096, 32000
    x, W = rnd(B, K), rnd(N, K, scale=K ** -0.5)
    g = (1 + 0.1 * torch.randn(K, device=DEV)).to(BF)
    xf = x.float()
    xn = (g.float() * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5)).to(BF).float()).to(BF)
    out = gemv(x, W, gamma=g, out_fp32=True)
    torch.testing.assert_close(out, ref_linear(xn, W), rtol=3e-3, atol=3e-3)


@pytest.mark.parametrize("B,I,h", [(1, 11008, 4096), (8, 13824, 5120), (2, 512, 256)])
def test_gemv_swiglu(B, I, h):
    x, Wg, Wu = rnd(B, h), rnd(I, h, scale=h ** -0.5), rnd(I, h, scale=h ** -0.5)
    Wgu = torch.empty(2 * I, h, device=DEV, dtype=BF)
    _b2.check(_b2.load_library().b2_op_interleave_gate_up(P(Wg), P(Wu), P(Wgu), I, h, S()))
    want = torch.nn.functional.silu(ref_linear(x, Wg)) * ref_linear(x, Wu)
    assert_close(gemv(x, Wgu, act=_b2.ACT_SWIGLU), want)


def test_gemv_matches_gemm_path():
    """decode (GEMV) and prefill (tcgen05 GEMM) paths must agree on the same operands."""
    x, W = rnd(8, 4096), rnd(4096, 4096, scale=1 / 64)
    torch.testing.assert_close(gemv(x, W, out_fp32=True), gemm(x, W, out_fp32=True), rtol=1e-3, atol=1e-3)


# ---------------------------------------------------------------------------------------------- misc
def test_argmax_first_occurrence_and_random():
    lib = _b2.load_library()
    x = torch.randn(5, 32000, device=DEV)
    x[1, 7] = x[1, 31999] = 100.0   # tie -> lowest index
    x[2, 31999] = 50.0              # last element
    x[3, 0] = 50.0                  # first element
    out = torch.empty(5, device=DEV, dtype=torch.int32)
    _b2.check(lib.b2_argmax(P(x), 5, 32000, P(out), S()))
    assert out.tolist() == x.argmax(-1).tolist()
    assert out[1].item() == 7


def test_im2col_matches_conv():
    B, img, ps, D = 2, 56, 14, 256
    kpad = 592
    pix, Wc = rnd(B, 3, img, img), rnd(D, 3, ps, ps, scale=0.05)
    col = torch.empty(B * (img // ps) ** 2, kpad, device=DEV, dtype=BF)
    _b2.check(_b2.load_library().b2_op_im2col(P(pix), P(col), B, img, ps, kpad, S()))
    assert (col[:, 588:] == 0).all()
    Wp = torch.zeros(D, kpad, device=DEV, dtype=BF)
    Wp[:, :588] = Wc.reshape(D, -1)
    got = gemm(col, Wp, bn=64).view(B, -1, D)
    want = torch.nn.functional.conv2d(pix.float(), Wc.float(), stride=ps).flatten(2).transpose(1, 2)
    assert_close(got, want)
