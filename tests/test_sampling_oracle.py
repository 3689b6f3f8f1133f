"""CPU: pins oracle/sampling_oracle.py — Philox4x32-10 against the Random123 known-answer vectors, the temperature /
top-k / top-p survivor rule against the INSTALLED transformers warpers (the arithmetic the reference's generate() runs,
llava/serve/model_worker.py:155-185), and the inverse-CDF draw against its own distribution."""
import numpy as np
import pytest

from oracle import sampling_oracle as S


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10, counter (0,0,0,0) key (0,0) -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8
    # (philox_u64 returns output words 0 and 1; only counters of the form (index, row, 0, 0) are used)
    assert S.philox_u64(0, 0, 0) == 0x6627E8D5E169C58D
    # distinct (index, row, seed) give distinct streams
    vals = {S.philox_u64(s, i, r) for s in (0, 1, 2**40 + 7) for i in range(8) for r in range(4)}
    assert len(vals) == 3 * 8 * 4


@pytest.mark.parametrize("seed", range(6))
def test_survivors_equal_hf_warpers(seed):
    rng = np.random.default_rng(seed)
    for _ in range(40):
        V = int(rng.integers(10, 4000))
        x = (rng.standard_normal(V) * rng.uniform(0.5, 6)).astype(np.float32)
        T = float(rng.choice([0.2, 0.7, 1.0, 1.5]))
        k = int(rng.choice([0, 1, 5, 50, V + 5]))
        p = float(rng.choice([0.05, 0.5, 0.7, 0.9, 0.99, 1.0]))
        ours, _ = S.kept_mask(x, T, k, p)
        assert np.array_equal(ours, S.kept_set_hf(x, T, k, p)), (V, T, k, p)
        assert ours.any()


def test_survivors_edge_cases():
    x = np.array([1.0, 3.0, 3.0, 2.0, -np.inf], dtype=np.float32)
    keep, e = S.kept_mask(x, 1.0, 1, 1.0)        # top-1 with a tie at the top: both maxima survive (scores >= k-th)
    assert keep.tolist() == [False, True, True, False, False]
    keep, _ = S.kept_mask(x, 1.0, 0, 1e-9)       # vanishing top_p: the largest probability always survives
    assert keep[1] and keep[2] and not keep[0] and not keep[3]
    keep, e = S.kept_mask(x, 1.0, 0, 1.0)
    assert keep.all() and e[4] == 0.0            # -inf survives the (disabled) filters with zero mass: never drawn
    assert S.greedy(np.array([0.0, np.nan, 5.0, 5.0], dtype=np.float32)) == 2


def test_draws_follow_the_renormalised_distribution():
    rng = np.random.default_rng(3)
    x = (rng.standard_normal(64) * 2).astype(np.float32)
    T, k, p = 0.8, 20, 0.9
    keep, e = S.kept_mask(x, T, k, p)
    probs = e.astype(np.float64) / e.sum()
    n = 20000
    counts = np.zeros(64)
    for i in range(n):
        tok, info = S.sample_row(x, T, k, p, seed=1234, index=i, row=0)
        assert keep[tok] and info["lo"][tok] <= info["target"] < info["hi"][tok]
        counts[tok] += 1
    exp = probs * n
    sel = exp > 5
    chi2 = (((counts - exp) ** 2)[sel] / exp[sel]).sum()
    assert chi2 < 3 * sel.sum(), (chi2, sel.sum())
    assert counts[~keep].sum() == 0
