"""Second reference-generated fixture, at a DIFFERENT configuration than tiny_prefill_decode.npz (so that the restatement is
not pinned at one shape only): hidden 512 / 4 heads / 3 layers, CLIP 4 layers / 8 heads / 64 patches, and NON-default
rms_norm_eps (3e-2) and rope_theta (50000) so that both parameters are exercised against the unmodified reference.

    python tests/golden/make_golden_small.py        (build container only: needs /root/reference)
"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import llava_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402
from make_golden import ref_greedy, weights_checksum  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CFG = O.make_config(hidden=512, inter=1024, layers=3, heads=4, vocab=2048, vit_hidden=512, vit_inter=1024, vit_layers=4,
                    vit_heads=8, image_size=112, patch_size=14, rms_eps=3e-2, rope_theta=50000.0)


def main():
    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    w = O.make_weights(CFG, seed=5)
    model = ref_shim.build_reference_model(CFG, w, os.path.join(tempfile.mkdtemp(prefix="b2golden_small_"), "clip"))
    g = torch.Generator().manual_seed(11)
    B, Lt, N = 2, 14, 6
    images = torch.randn(B, 3, CFG["image_size"], CFG["image_size"], generator=g).half().float()  # fp16-exact: stored as fp16
    ids = torch.randint(3, CFG["vocab"], (B, Lt), generator=g)
    ids[:, 0] = 1
    ids[0, 4] = O.IMAGE_TOKEN_INDEX
    ids[1, 9] = O.IMAGE_TOKEN_INDEX
    feats = model.encode_images(images)
    out = model(input_ids=ids, images=images, use_cache=True)
    toks, step_logits = ref_greedy(model, ids, images, N)
    np.savez_compressed(
        os.path.join(OUT, "small_prefill_decode.npz"), images=images.numpy().astype(np.float16), input_ids=ids.numpy(),
        image_features=feats.numpy(), last_logits=out.logits[:, -3:].float().numpy(), greedy_tokens=toks.numpy(),
        step_logits=step_logits.numpy(), weights_checksum=np.float64(weights_checksum(w)), seed=np.int64(5))
    print("wrote small_prefill_decode.npz")


if __name__ == "__main__":
    assert ref_shim.available(), "reference tree not found (this script only runs in the build container)"
    main()
