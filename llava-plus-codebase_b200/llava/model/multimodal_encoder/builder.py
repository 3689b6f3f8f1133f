"""`build_vision_tower(cfg, **kw)` — the factory the reference's model assembly calls
(llava/model/multimodal_encoder/builder.py:5-11, invoked from llava/model/llava_arch.py:32-34).

Contract kept: the tower name is read from `cfg.mm_vision_tower` (falling back to `cfg.vision_tower`); a name
that is an existing local path, or an `openai/...` / `laion/...` hub id, yields a `CLIPVisionTower`; anything
else is a `ValueError`. The returned object only holds the CLIP checkpoint tensors — its forward pass is the
`b2_vit_encode` entry point of libb2llava.so.
"""
import os

from .clip_encoder import CLIPVisionTower

_HUB_PREFIXES = ("openai", "laion")


def _tower_name(cfg):
    for attr in ("mm_vision_tower", "vision_tower"):
        name = getattr(cfg, attr, None)
        if name is not None:
            return name
    return None


def build_vision_tower(vision_tower_cfg, **kwargs):
    name = _tower_name(vision_tower_cfg)
    known = name is not None and (os.path.exists(name) or name.startswith(_HUB_PREFIXES))
    if not known:
        raise ValueError(f"Unknown vision tower: {name}")
    return CLIPVisionTower(name, args=vision_tower_cfg, **kwargs)
