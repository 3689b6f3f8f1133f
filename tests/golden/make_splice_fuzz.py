"""Generate tests/golden/tiny_splice_fuzz.npz by running randomised batches through the UNMODIFIED reference
`prepare_inputs_labels_for_multimodal` (llava/model/llava_arch.py:99-240, via oracle/ref_shim.py) in the build container:

    python tests/golden/make_splice_fuzz.py

To keep the fixture small the reference's output embeddings are stored as a SOURCE INDEX: every output row is matched
bit-exactly against the rows of embed_tokens.weight (>= 0: token id) and of the reference's own encode_images output
(< 0: -(feature row) - 1); all-zero padding rows become INT32_MIN. tests/test_splice_host.py feeds the same inputs to
llava.model.llava_arch.build_source_index and requires the identical index, mask, position ids and labels."""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import llava_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
IMG, PAD = O.IMAGE_TOKEN_INDEX, -(2 ** 31)
N_CASES = 48


def main():
    torch.set_grad_enabled(False)
    cfg = O.CONFIGS["tiny"]
    w = O.make_weights(cfg, seed=0)
    model = ref_shim.build_reference_model(cfg, w, os.path.join(tempfile.mkdtemp(prefix="b2fuzz_"), "clip"))
    table = model.get_model().embed_tokens.weight.detach().float().numpy()
    row_of = {table[i].tobytes(): i for i in range(table.shape[0])}
    assert len(row_of) == table.shape[0], "embedding rows must be unique for the row matching"
    rng = np.random.default_rng(7)
    out = {"n_cases": np.int64(N_CASES)}
    for c in range(N_CASES):
        B, Lt = int(rng.integers(1, 5)), int(rng.integers(3, 14))
        ids = rng.integers(3, cfg["vocab"], size=(B, Lt)).astype(np.int64)
        mask = np.zeros((B, Lt), dtype=np.int64)
        slots = 0
        for b in range(B):
            n = int(rng.integers(1, Lt + 1))
            mask[b, :n] = 1
            k = int(rng.integers(0, 4))
            where = rng.choice(n, size=min(k, n), replace=False)
            ids[b, where] = IMG
            slots += max(len(where), 1)
        labels = rng.integers(0, cfg["vocab"], size=(B, Lt)).astype(np.int64)
        side = "left" if c % 2 else "right"
        max_len = 0 if c % 3 else int(rng.integers(4, 40))
        images = torch.from_numpy(rng.standard_normal((slots, 3, cfg["image_size"], cfg["image_size"])).astype(np.float32))
        model.config.tokenizer_padding_side = side
        model.config.tokenizer_model_max_length = max_len or None
        feats = model.encode_images(images).float().numpy()          # [slots, P, h]
        P = feats.shape[1]
        frow = {feats.reshape(-1, feats.shape[-1])[i].tobytes(): i for i in range(slots * P)}
        _, pos, am, _, emb, lab = model.prepare_inputs_labels_for_multimodal(
            torch.from_numpy(ids), torch.arange(Lt)[None].expand(B, Lt), torch.from_numpy(mask), None,
            torch.from_numpy(labels), images)
        emb = emb.float().numpy()
        src = np.empty(emb.shape[:2], dtype=np.int64)
        for b in range(emb.shape[0]):
            for s in range(emb.shape[1]):
                key = emb[b, s].tobytes()
                if not emb[b, s].any():
                    src[b, s] = PAD
                elif key in frow:
                    src[b, s] = -frow[key] - 1
                else:
                    src[b, s] = row_of[key]  # KeyError = a row that is neither a token nor a feature row
        out.update({f"c{c}_ids": ids, f"c{c}_mask": mask, f"c{c}_labels": labels, f"c{c}_left": np.int64(side == "left"),
                    f"c{c}_maxlen": np.int64(max_len), f"c{c}_slots": np.int64(slots), f"c{c}_P": np.int64(P),
                    f"c{c}_src": src.astype(np.int64), f"c{c}_omask": am.numpy().astype(np.int64),
                    f"c{c}_opos": pos.numpy().astype(np.int64), f"c{c}_olabels": lab.numpy().astype(np.int64)})
    np.savez_compressed(os.path.join(OUT, "tiny_splice_fuzz.npz"), **out)
    print("wrote", N_CASES, "cases")


if __name__ == "__main__":
    assert ref_shim.available(), "reference tree not found (this script only runs in the build container)"
    main()
