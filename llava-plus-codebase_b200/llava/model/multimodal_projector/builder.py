"""build_vision_projector — parameter holders for the mm_projector (ref multimodal_projector/builder.py:33-51).

The modules built here only HOLD the checkpoint tensors under the reference's state-dict names
(`mm_projector.0.weight`, `mm_projector.2.weight`, ...). The arithmetic (Linear -> exact-erf GELU -> Linear)
runs in libb2llava.so as two tcgen05 GEMMs with fused bias/GELU epilogues (b2_project)."""
import re

import torch.nn as nn


class IdentityMap(nn.Module):
    def forward(self, x, *args, **kwargs):
        return x

    @property
    def config(self):
        return {"mm_projector_type": "identity"}


class _EngineProjector(nn.Sequential):
    """nn.Sequential(Linear, GELU, Linear) layout for state-dict compatibility; calls go to the engine."""

    def forward(self, x):
        owner = getattr(self, "_owner", None)
        model = owner() if owner is not None else None
        if model is None:
            raise RuntimeError("mm_projector is not attached to a LlavaLlamaForCausalLM; no PyTorch fallback exists")
        return model._ensure_engine().project(x)


def build_vision_projector(config, delay_load=False, **kwargs):
    projector_type = getattr(config, "mm_projector_type", "linear")
    m = re.match(r"^mlp(\d+)x_gelu$", projector_type)
    if m and int(m.group(1)) == 2:
        return _EngineProjector(
            nn.Linear(config.mm_hidden_size, config.hidden_size),
            nn.GELU(),
            nn.Linear(config.hidden_size, config.hidden_size),
        )
    if projector_type in ("linear", "identity") or m:
        raise NotImplementedError(
            f"mm_projector_type={projector_type!r}: only the LLaVA-1.5 'mlp2x_gelu' projector has a B200 kernel path")
    raise ValueError(f"Unknown projector type: {projector_type}")
