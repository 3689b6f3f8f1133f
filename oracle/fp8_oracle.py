"""CPU restatement of the fp8 decode path (BASELINE.json configs[4]: "fp8-weight tcgen05 path") — TEST INFRASTRUCTURE ONLY.

The reference has no fp8 path ("no reference exists for it at all; tolerance must be derived"), so this
file DEFINES the arithmetic the CUDA kernels implement and derives the tolerance against the bf16 oracle
(oracle/llava_oracle.py, which restates llava/model/language_model/llava_llama.py:56-99 + HF modeling_llama.py):

  quantize_rows_e4m3   scale[r] = amax_r / 448 (1 for a zero row); q = e4m3_rne(x * (448 / amax_r))        (csrc/quant_fp8.cu)
  linear_w8a8          y[b,n] = (sum_k qx[b,k] * qw[n,k]) * sx[b] * sw[n]   fp32 accumulation                (csrc/gemm_skinny.cu, FP8)
  decode_layer_w8a8    LlamaDecoderLayer one-token step with the four Linears (and lm_head) replaced by linear_w8a8;
                       RMSNorm, RoPE, attention, residuals and the KV cache stay bf16 exactly as in the bf16 oracle

Weights: one scale per OUTPUT channel; activations: one scale per TOKEN, computed on the fly. torch.float8_e4m3fn casts
round to nearest even; every value fed to the cast is <= 448 * (1 + 2^-23) in magnitude, which rounds to 448, so the
cast never produces the NaN encoding and equals the hardware's satfinite conversion.
"""
import torch

E4M3_MAX = 448.0


def quantize_rows_e4m3(x: torch.Tensor):
    """x [..., K] (any float dtype; arithmetic in fp32) -> (q float8_e4m3fn [..., K], scale fp32 [...])."""
    xf = x.float()
    amax = xf.abs().amax(dim=-1)
    pos = amax > 0
    # tensor / tensor: IEEE division like the kernel's `448.0f / amax` (torch evaluates `scalar / tensor` as scalar * reciprocal,
    # which is 1 ulp off for e.g. amax = 12 and flips exact ties such as 4.5 * 448/12 = 168 between the codes 160 and 176)
    inv = torch.where(pos, torch.full_like(amax, E4M3_MAX) / amax, torch.ones_like(amax))
    scale = torch.where(pos, amax / E4M3_MAX, torch.ones_like(amax))
    q = (xf * inv.unsqueeze(-1)).to(torch.float8_e4m3fn)
    return q, scale


def linear_w8a8(qx, sx, qw, sw, residual=None):
    """qx [B,K] / qw [N,K] float8_e4m3fn, sx [B], sw [N] fp32 -> fp32 [B,N]."""
    y = (qx.float() @ qw.float().t()) * sx[:, None] * sw[None, :]
    return y if residual is None else y + residual.float()


def linear_fake_quant(x, w):
    """bf16 activations x [B,K], bf16 weight w [N,K] -> fp32 [B,N] through per-token / per-channel e4m3 quantisation."""
    qx, sx = quantize_rows_e4m3(x)
    qw, sw = quantize_rows_e4m3(w)
    return linear_w8a8(qx, sx, qw, sw)


def rmsnorm_hf(x, gamma, eps):
    """LlamaRMSNorm (HF modeling_llama.py:62-67) with its bf16 rounding points: gamma * bf16(x * rstd) -> bf16."""
    xf = x.float()
    rstd = torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return (gamma.float() * (xf * rstd).to(torch.bfloat16).float()).to(torch.bfloat16)


def swiglu_w8a8(x_bf16, wg, wu):
    """silu(gate(x)) * up(x) with both projections through linear_fake_quant; returns bf16 like the kernel's epilogue."""
    g = linear_fake_quant(x_bf16, wg)
    u = linear_fake_quant(x_bf16, wu)
    return (torch.nn.functional.silu(g) * u).to(torch.bfloat16)


def expected_relative_error(K: int) -> float:
    """Model of the W8A8 error of one Linear relative to the output's RMS: e4m3 keeps 3 mantissa bits, so each operand
    carries a uniform relative rounding error of RMS 2^-4 / sqrt(3) ~ 3.6 %; with independent errors on x and w every
    product is off by ~5.1 % RMS of its own magnitude and the K products add incoherently, as do the exact terms — the
    relative error of the SUM therefore stays ~5 % of the output RMS independent of K (it does not average down,
    because signal and noise both grow as sqrt(K)). Used by the tests as the scale of the tolerance."""
    del K
    return 2.0 ** -4 / 3 ** 0.5 * 2 ** 0.5
