"""Generate the committed golden fixtures by running the UNMODIFIED reference files (via oracle/ref_shim.py)
in the build container:   python tests/golden/make_golden.py

The reference has no tests or golden vectors of its own, so these outputs of the reference itself
are what pins oracle/llava_oracle.py (tests/test_oracle_golden.py) and, through it, the CUDA path.
Weights are regenerated from a seed by oracle.make_weights (a checksum is stored to detect generator drift).
"""
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
IMG = O.IMAGE_TOKEN_INDEX


def weights_checksum(w):
    return float(sum(float(v.double().abs().sum()) for v in w.values()))


def ref_greedy(model, ids, images, n):
    """Manual greedy loop over the reference forward (HF generate breaks at llava_arch.py:105)."""
    out = model(input_ids=ids, images=images, use_cache=True)
    pkv = out.past_key_values
    toks, logs = [], []
    logits = out.logits
    for i in range(n):
        last = logits[:, -1].float()
        logs.append(last)
        nxt = last.argmax(-1, keepdim=True)
        toks.append(nxt)
        if i == n - 1:
            break
        o = model(input_ids=nxt, past_key_values=pkv, use_cache=True)
        pkv, logits = o.past_key_values, o.logits
    return torch.cat(toks, dim=1), torch.stack(logs)


def main():
    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    cfg = O.CONFIGS["tiny"]
    w = O.make_weights(cfg, seed=0)
    tmp = tempfile.mkdtemp(prefix="b2golden_")
    model = ref_shim.build_reference_model(cfg, w, os.path.join(tmp, "clip"))
    g = torch.Generator().manual_seed(1)
    P = (cfg["image_size"] // cfg["patch_size"]) ** 2

    # ---- case A: batched prefill + greedy decode, one image per row ----
    B, Lt, N = 2, 12, 6
    images = torch.randn(B, 3, cfg["image_size"], cfg["image_size"], generator=g)
    ids = torch.randint(3, cfg["vocab"], (B, Lt), generator=g)
    ids[:, 0] = 1
    ids[:, 5] = IMG
    feats = model.encode_images(images)
    tower = model.get_vision_tower()(images)
    _, _, _, _, embeds, _ = model.prepare_inputs_labels_for_multimodal(ids, None, None, None, None, images)
    out = model(input_ids=ids, images=images, use_cache=True)
    toks, step_logits = ref_greedy(model, ids, images, N)
    np.savez_compressed(
        os.path.join(OUT, "tiny_prefill_decode.npz"), images=images.numpy(), input_ids=ids.numpy(),
        tower_features=tower.numpy(), image_features=feats.numpy(), inputs_embeds=embeds.numpy(),
        logits=out.logits.float().numpy(), greedy_tokens=toks.numpy(), step_logits=step_logits.numpy(),
        weights_checksum=np.float64(weights_checksum(w)), seed=np.int64(0))

    # ---- case B: splice edge cases (multi-image row, zero-image row, padded row, labels) ----
    Lt = 10
    ids = torch.randint(3, cfg["vocab"], (3, Lt), generator=g)
    ids[0, 2] = IMG
    ids[0, 7] = IMG           # row 0: two images
    # row 1: no image (still consumes one image slot)
    ids[2, 4] = IMG           # row 2: one image, last 3 positions padded out by the mask
    mask = torch.ones(3, Lt, dtype=torch.long)
    mask[2, 7:] = 0
    labels = torch.randint(3, cfg["vocab"], (3, Lt), generator=g)
    images = torch.randn(4, 3, cfg["image_size"], cfg["image_size"], generator=g)  # 2 + 1 (dummy) + 1
    res = {}
    for side in ("right", "left"):
        model.config.tokenizer_padding_side = side
        _, pos, am, _, emb, lab = model.prepare_inputs_labels_for_multimodal(
            ids, torch.arange(Lt)[None].expand(3, Lt), mask, None, labels, images)
        res[side] = (emb.numpy(), am.numpy(), pos.numpy(), lab.numpy())
    model.config.tokenizer_padding_side = "right"
    model.config.tokenizer_model_max_length = 20
    _, pos, am, _, emb, lab = model.prepare_inputs_labels_for_multimodal(ids, torch.arange(Lt)[None].expand(3, Lt), mask, None, labels, images)
    res["trunc"] = (emb.numpy(), am.numpy(), pos.numpy(), lab.numpy())
    model.config.tokenizer_model_max_length = None
    # 5-D images: each row's 2 images are flattened into ONE <image> slot (App. C.3)
    ids5 = torch.randint(3, cfg["vocab"], (2, 6), generator=g)
    ids5[:, 1] = IMG
    images5 = torch.randn(2, 2, 3, cfg["image_size"], cfg["image_size"], generator=g)
    _, _, _, _, emb5, _ = model.prepare_inputs_labels_for_multimodal(ids5, None, None, None, None, images5)
    np.savez_compressed(
        os.path.join(OUT, "tiny_splice_edges.npz"), input_ids=ids.numpy(), attention_mask=mask.numpy(),
        labels=labels.numpy(), images=images.numpy(),
        **{f"{k}_{n}": v for k, vals in res.items() for n, v in zip(("embeds", "mask", "pos", "labels"), vals)},
        ids5=ids5.numpy(), images5=images5.numpy(), embeds5=emb5.numpy(),
        weights_checksum=np.float64(weights_checksum(w)), seed=np.int64(0))
    print("wrote fixtures to", OUT, "P =", P)


if __name__ == "__main__":
    assert ref_shim.available(), "reference tree not found (this script only runs in the build container)"
    main()
