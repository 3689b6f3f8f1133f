"""CPU restatement (numpy) of the token selection of the decode loop — TEST INFRASTRUCTURE ONLY (never imported by the
product; tests/ and __graft_entry__.smoke() are the only callers).

What it restates: HF `GenerationMixin` token selection as the reference's callers invoke it
(/root/reference/llava/serve/model_worker.py:155-185: do_sample = temperature > 0.001, temperature, top_p;
llava/eval/model_vqa_loader.py:98-106: greedy when temperature == 0). The arithmetic lives in the third-party
`transformers` package (reference pins 4.31.0, pyproject.toml:17; installed here: 5.5.0):
  TemperatureLogitsWarper  scores / temperature
  TopKLogitsWarper         remove scores < k-th largest score (ties with the k-th value survive)
  TopPLogitsWarper         sort ascending, softmax, remove where cumsum <= 1 - top_p, keep >= 1 token
  multinomial draw from softmax of what is left.
`kept_set_hf` below calls those installed warpers directly; tests/test_sampling_oracle.py pins `kept_mask` (this file's own
statement of the same rule: keep token i iff the probability mass strictly above it is < top_p) against them.

The draw itself cannot be compared with torch.multinomial (different RNG); the product (csrc/sampling.cu) defines it as an
inverse CDF in token-index order over 2^-40 fixed-point masses with a Philox4x32-10 stream keyed by (seed; token index,
row). `sample_row` restates exactly that, so a GPU draw is checked against the CDF interval of the token it returned.
"""
import numpy as np

FIXED_ONE = float(2 ** 40)
M32 = 0xFFFFFFFF


def philox_u64(seed: int, index: int, row: int) -> int:
    """Philox4x32-10, counter (index, row, 0, 0), key = seed split in two words; returns (c0 << 32) | c1."""
    c = [index & M32, row & M32, 0, 0]
    k0, k1 = seed & M32, (seed >> 32) & M32
    for _ in range(10):
        p0 = 0xD2511F53 * c[0]
        p1 = 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c[3] ^ k1) & M32, p0 & M32]
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return (c[0] << 32) | c[1]


def kept_mask(logits_row, temperature=1.0, top_k=0, top_p=1.0):
    """Survivors of temperature -> top-k -> top-p, and their softmax numerators e_i = exp(x_i - max) (float32).

    top-k: keep x_i >= (k-th largest x). top-p: keep i iff sum_{j: e_j > e_i} e_j < top_p * sum_j e_j, the largest always
    survives (HF: ascending cumsum <= 1 - top_p is removed; min_tokens_to_keep = 1)."""
    x = (np.asarray(logits_row, dtype=np.float32) * np.float32(1.0 / np.float32(temperature))).astype(np.float32)
    V = x.shape[0]
    keep = np.ones(V, dtype=bool)
    if top_k and 0 < top_k < V:
        kth = np.sort(x)[V - top_k]
        keep &= x >= kth
    mx = x[keep].max()
    e = np.where(keep, np.exp((x - mx).astype(np.float32)), np.float32(0)).astype(np.float32)
    if top_p is not None and top_p < 1.0:
        mass = np.floor(e.astype(np.float64) * FIXED_ONE)
        total = mass.sum()
        limit = max(np.floor(total * float(np.float32(top_p))), 1.0)  # the kernel holds top_p as fp32
        order = np.argsort(-e, kind="stable")
        es, ms = e[order], mass[order]
        # mass strictly above each element: cumulative mass of strictly larger values (ties share the same "above")
        cum = np.concatenate([[0.0], np.cumsum(ms)[:-1]])
        first_of_value = np.concatenate([[True], es[1:] != es[:-1]])
        above = np.maximum.accumulate(np.where(first_of_value, cum, 0.0))
        k2 = np.zeros(V, dtype=bool)
        k2[order] = above < limit
        keep &= k2
        e = np.where(keep, e, np.float32(0)).astype(np.float32)
    return keep, e


def cdf_intervals(e):
    """Index-order fixed-point CDF of the surviving masses: (lo[V], hi[V], total) as Python-int-exact float64."""
    mass = np.floor(e.astype(np.float64) * FIXED_ONE)
    hi = np.cumsum(mass)
    return hi - mass, hi, float(hi[-1])


def sample_row(logits_row, temperature, top_k, top_p, seed, index, row):
    """The token csrc/sampling.cu draws for (row, index), plus (target, lo, hi, total) of the decision."""
    keep, e = kept_mask(logits_row, temperature, top_k, top_p)
    lo, hi, total = cdf_intervals(e)
    target = (int(total) * philox_u64(seed, index, row)) >> 64
    tok = int(np.searchsorted(hi, target, side="right"))
    return tok, dict(target=float(target), lo=lo, hi=hi, total=total, keep=keep)


def greedy(logits_row):
    """torch.argmax semantics: first occurrence of the maximum, NaN skipped."""
    x = np.asarray(logits_row, dtype=np.float32)
    x = np.where(np.isnan(x), -np.inf, x)
    return int(np.argmax(x))


def kept_set_hf(logits_row, temperature=1.0, top_k=0, top_p=1.0):
    """The same survivors through the INSTALLED transformers warpers (the reference's actual arithmetic)."""
    import torch
    from transformers.generation.logits_process import TemperatureLogitsWarper, TopKLogitsWarper, TopPLogitsWarper

    s = torch.tensor(np.asarray(logits_row, dtype=np.float32))[None]
    ids = torch.zeros(1, 1, dtype=torch.long)
    if temperature != 1.0:
        s = TemperatureLogitsWarper(float(temperature))(ids, s)
    if top_k and top_k > 0:
        s = TopKLogitsWarper(int(top_k))(ids, s)
    if top_p is not None and top_p < 1.0:
        s = TopPLogitsWarper(float(top_p))(ids, s)
    return torch.isfinite(s[0]).numpy()
