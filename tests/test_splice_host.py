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
