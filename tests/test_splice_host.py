"""CPU: host half of the splice (llava.model.llava_arch.build_source_index) against the golden outputs of the
reference's prepare_inputs_labels_for_multimodal and against the oracle; the device gather is emulated in numpy."""
import os

import numpy as np
import pytest
import torch

from llava.model.llava_arch import build_source_index, _PAD_ROW
from oracle import llava_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _emulate_gather(src, table, feats):
    B, S = src.shape
    out = np.zeros((B, S, table.shape[1]), dtype=np.float32)
    for b in range(B):
        for s in range(S):
            v = int(src[b, s])
            if v == _PAD_ROW:
                continue
            out[b, s] = table[v] if v >= 0 else feats[-v - 1]
    return out


@pytest.mark.parametrize("case,side,max_len", [("right", "right", None), ("left", "left", None), ("trunc", "right", 20)])
def test_source_index_reproduces_reference_splice(case, side, max_len):
    g = np.load(os.path.join(GOLD, "tiny_splice_edges.npz"))
    cfg = O.CONFIGS["tiny"]
    w = O.make_weights(cfg, seed=int(g["seed"]))
    feats = O.encode_images(w, torch.from_numpy(g["images"]), cfg).numpy()  # [4, P, h]
    P = feats.shape[1]
    src, labels, mask, pos, lens = build_source_index(
        g["input_ids"].astype(np.int64), g["attention_mask"].astype(bool), g["labels"].astype(np.int64),
        feats.shape[0] * P, [P] * feats.shape[0], max_len, side)
    emb = _emulate_gather(src, w["model.embed_tokens.weight"].numpy(), feats.reshape(-1, feats.shape[-1]))
    np.testing.assert_allclose(emb, g[f"{case}_embeds"], rtol=2e-4, atol=2e-4)
    assert (mask == g[f"{case}_mask"].astype(bool)).all()
    assert (pos == g[f"{case}_pos"]).all()
    assert (labels == g[f"{case}_labels"]).all()
    assert lens == [int(x) for x in g[f"{case}_mask"].astype(bool).sum(1)]


def test_grouped_images_flatten_into_one_slot():
    g = np.load(os.path.join(GOLD, "tiny_splice_edges.npz"))
    cfg = O.CONFIGS["tiny"]
    w = O.make_weights(cfg, seed=int(g["seed"]))
    im5 = torch.from_numpy(g["images5"])
    feats = O.encode_images(w, im5.flatten(0, 1), cfg).numpy()
    P = feats.shape[1]
    ids = g["ids5"].astype(np.int64)
    src, *_ = build_source_index(ids, np.ones_like(ids, bool), np.full_like(ids, -100), feats.shape[0] * P,
                                 [2 * P, 2 * P], None, "right")
    emb = _emulate_gather(src, w["model.embed_tokens.weight"].numpy(), feats.reshape(-1, feats.shape[-1]))
    np.testing.assert_allclose(emb, g["embeds5"], rtol=2e-4, atol=2e-4)


def test_missing_image_slot_raises_like_reference():
    ids = np.array([[1, 5, 6, 7], [1, O.IMAGE_TOKEN_INDEX, 5, 6]], dtype=np.int64)
    with pytest.raises(IndexError):  # row 0 consumes the only slot
        build_source_index(ids, np.ones_like(ids, bool), np.full_like(ids, -100), 16, [16], None, "right")


def test_empty_and_ragged_rows():
    ids = np.array([[1, 2, 3, 4, 5, 6], [1, O.IMAGE_TOKEN_INDEX, 9, 0, 0, 0]], dtype=np.int64)
    mask = np.array([[1, 1, 1, 1, 1, 1], [1, 1, 1, 0, 0, 0]], dtype=bool)
    src, labels, m, pos, lens = build_source_index(ids, mask, np.full_like(ids, -100), 8, [4, 4], None, "right")
    assert lens == [6, 6]  # 2 text + 4 feature rows
    assert (src[1] == np.array([1, -5, -6, -7, -8, 9])).all()  # second slot: rows 4..7 -> -(r)-1
    assert m.all() and (pos[1] == np.arange(6)).all()


def _random_case(rng, B, P, n_feat_rows_per_slot):
    """Random batch in the reference's input format: ragged rows (attention_mask), 0..3 <image> per row, labels."""
    Lt = int(rng.integers(3, 14))
    ids = rng.integers(3, 90, size=(B, Lt)).astype(np.int64)
    mask = np.zeros((B, Lt), dtype=bool)
    slots = 0
    for b in range(B):
        n = int(rng.integers(1, Lt + 1))
        mask[b, :n] = True
        k = int(rng.integers(0, 4))
        where = rng.choice(n, size=min(k, n), replace=False)
        ids[b, where] = O.IMAGE_TOKEN_INDEX
        slots += max(len(where), 1)  # a row without <image> still consumes one slot
    labels = rng.integers(0, 90, size=(B, Lt)).astype(np.int64)
    feats = [rng.standard_normal((n_feat_rows_per_slot, 8)).astype(np.float32) for _ in range(slots)]
    return ids, mask, labels, feats


@pytest.mark.parametrize("seed", range(40))
def test_source_index_fuzz_against_the_restatement(seed):
    """Randomised batches through build_source_index + emulated gather vs oracle.prepare_multimodal (itself pinned to the
    unmodified reference by tests/test_oracle_golden.py): embeddings, mask, position ids, labels and lengths must be
    identical for both padding sides, with and without truncation."""
    rng = np.random.default_rng(seed)
    B, P = int(rng.integers(1, 5)), int(rng.integers(1, 6))
    ids, mask, labels, feats = _random_case(rng, B, P, P)
    table = rng.standard_normal((100, 8)).astype(np.float32)
    w = {"model.embed_tokens.weight": torch.from_numpy(table)}
    side = "left" if seed % 2 else "right"
    max_len = None if seed % 3 else int(rng.integers(2, 12))
    want_e, want_m, want_p, want_l = O.prepare_multimodal(
        w, torch.from_numpy(ids), None, {}, attention_mask=torch.from_numpy(mask), labels=torch.from_numpy(labels),
        padding_side=side, max_length=max_len, image_features=[torch.from_numpy(f) for f in feats])
    src, new_labels, m, pos, lens = build_source_index(ids, mask, labels, len(feats) * P, [P] * len(feats), max_len, side)
    emb = _emulate_gather(src, table, np.concatenate(feats, 0))
    np.testing.assert_array_equal(emb, want_e.numpy())
    assert (m == want_m.numpy()).all() and (pos == want_p.numpy()).all() and (new_labels == want_l.numpy()).all()
    assert lens == [int(x) for x in want_m.sum(1)]


def test_source_index_equals_the_unmodified_reference_on_random_batches():
    """48 randomised batches (ragged rows, 0-3 <image> per row, labels, both padding sides, truncation) run through the
    UNMODIFIED reference prepare_inputs_labels_for_multimodal (tests/golden/make_splice_fuzz.py): the reference's output
    embeddings, stored as a row-matched source index, its mask, position ids and labels must be reproduced exactly."""
    g = np.load(os.path.join(GOLD, "tiny_splice_fuzz.npz"))
    n = int(g["n_cases"])
    assert n >= 40
    for c in range(n):
        ids, mask, labels = g[f"c{c}_ids"], g[f"c{c}_mask"].astype(bool), g[f"c{c}_labels"]
        side = "left" if int(g[f"c{c}_left"]) else "right"
        max_len = int(g[f"c{c}_maxlen"]) or None
        slots, P = int(g[f"c{c}_slots"]), int(g[f"c{c}_P"])
        src, new_labels, m, pos, lens = build_source_index(ids, mask, labels, slots * P, [P] * slots, max_len, side)
        want = g[f"c{c}_src"]
        assert src.shape == want.shape, (c, src.shape, want.shape)
        assert (src.astype(np.int64) == want).all(), f"case {c}: source index differs from the reference"
        assert (m == g[f"c{c}_omask"].astype(bool)).all() and (pos == g[f"c{c}_opos"]).all(), f"case {c}"
        assert (new_labels == g[f"c{c}_olabels"]).all(), f"case {c}"
        assert lens == [int(x) for x in g[f"c{c}_omask"].astype(bool).sum(1)]
