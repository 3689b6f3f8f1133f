"""CLIPVisionTower — same surface as the reference's llava/model/multimodal_encoder/clip_encoder.py:7-78
(is_loaded, load_model, feature_select semantics, forward, dtype/device/config/hidden_size/num_patches/
dummy_feature), but the module only HOLDS the CLIP ViT-L/14-336 checkpoint tensors under the HF
`vision_model.*` names; the forward pass (patch-embed GEMM, 23 live encoder layers, hidden_states[-2], CLS
dropped) runs in libb2llava.so (b2_vit_encode). The encoder layer above the selected hidden state and
post_layernorm — which the reference computes and discards — are never computed.
"""
import glob
import json
import os
import weakref

import torch
import torch.nn as nn
from transformers import CLIPImageProcessor, CLIPVisionConfig


class _Holder(nn.Module):
    """Parameter container (no forward)."""


def _p(*shape):
    return nn.Parameter(torch.empty(*shape), requires_grad=False)


def _ln(D):
    m = _Holder()
    m.weight, m.bias = _p(D), _p(D)
    return m


def _lin(o, i, bias=True):
    m = _Holder()
    m.weight = _p(o, i)
    if bias:
        m.bias = _p(o)
    return m


class _CLIPVisionWeights(nn.Module):
    """State-dict layout of transformers CLIPVisionModel (`vision_model.embeddings...`, `.encoder.layers.N...`)."""

    def __init__(self, cfg):
        super().__init__()
        self.config = cfg
        D, I = cfg.hidden_size, cfg.intermediate_size
        vm = _Holder()
        emb = _Holder()
        emb.class_embedding = _p(D)
        emb.patch_embedding = _Holder()
        emb.patch_embedding.weight = _p(D, cfg.num_channels, cfg.patch_size, cfg.patch_size)
        emb.position_embedding = _Holder()
        emb.position_embedding.weight = _p((cfg.image_size // cfg.patch_size) ** 2 + 1, D)
        vm.embeddings = emb
        vm.pre_layrnorm = _ln(D)
        enc = _Holder()
        layers = []
        for _ in range(cfg.num_hidden_layers):
            L = _Holder()
            L.self_attn = _Holder()
            for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
                setattr(L.self_attn, n, _lin(D, D))
            L.layer_norm1 = _ln(D)
            L.mlp = _Holder()
            L.mlp.fc1 = _lin(I, D)
            L.mlp.fc2 = _lin(D, I)
            L.layer_norm2 = _ln(D)
            layers.append(L)
        enc.layers = nn.ModuleList(layers)
        vm.encoder = enc
        vm.post_layernorm = _ln(D)
        self.vision_model = vm

    @property
    def dtype(self):
        return self.vision_model.embeddings.class_embedding.dtype

    @property
    def device(self):
        return self.vision_model.embeddings.class_embedding.device


def _read_checkpoint_dir(path):
    """Load a HF checkpoint directory (safetensors or torch .bin, sharded or not) into a flat dict."""
    sd = {}
    st_files = sorted(glob.glob(os.path.join(path, "*.safetensors")))
    if st_files:
        from safetensors.torch import load_file
        for f in st_files:
            sd.update(load_file(f))
        return sd
    for f in sorted(glob.glob(os.path.join(path, "pytorch_model*.bin"))):
        sd.update(torch.load(f, map_location="cpu", weights_only=True))
    return sd


def _resolve_checkpoint_dir(name):
    """Local directory of a checkpoint named like the reference names its tower (ref clip_encoder.py:22-27 hands the name
    to `from_pretrained`): an existing directory is used as is; a hub id such as `openai/clip-vit-large-patch14-336` (what a
    stock llava-v1.5 config.json carries in `mm_vision_tower`, ref builder.py:140-144) resolves through the local
    HuggingFace cache. Nothing is downloaded here."""
    if os.path.isdir(name):
        return name
    try:
        from huggingface_hub import snapshot_download
        return snapshot_download(name, local_files_only=True)
    except Exception as e:
        raise RuntimeError(
            f"vision tower {name!r} is neither a local directory nor present in the local HuggingFace cache "
            f"(HF_HOME / HF_HUB_CACHE); this path never downloads: populate the cache (`huggingface-cli download {name}`) "
            f"or point mm_vision_tower at a directory with config.json, preprocessor_config.json and the weights") from e


class CLIPVisionTower(nn.Module):
    def __init__(self, vision_tower, args, delay_load=False):
        super().__init__()
        self.is_loaded = False
        self.vision_tower_name = vision_tower
        self.select_layer = args.mm_vision_select_layer
        self.select_feature = getattr(args, "mm_vision_select_feature", "patch")
        if self.select_feature != "patch":
            # ref clip_encoder.py:31-36 also offers 'cls_patch'; LLaVA-1.5 uses 'patch' (train.py:71)
            raise NotImplementedError(f"mm_vision_select_feature={self.select_feature!r}: only 'patch' has a kernel path")
        self._owner = None
        if not delay_load:
            self.load_model()
        else:
            self.cfg_only = CLIPVisionConfig.from_pretrained(self.vision_tower_name)

    def load_model(self, random_init=False):
        """ref clip_encoder.py:22-27. Reads the CLIP checkpoint tensors (frozen) into the holder; the engine
        ingests them at the next call. `random_init=True` leaves the tensors for the caller to fill
        (benchmarks/tests: there are no checkpoints offline)."""
        self.image_processor = CLIPImageProcessor.from_pretrained(self.vision_tower_name)
        cfg = CLIPVisionConfig.from_pretrained(self.vision_tower_name)
        self.vision_tower = _CLIPVisionWeights(cfg)
        if not random_init:
            ckpt_dir = _resolve_checkpoint_dir(self.vision_tower_name)
            sd = {k: v for k, v in _read_checkpoint_dir(ckpt_dir).items() if k.startswith("vision_model.")}
            if not sd:
                raise RuntimeError(
                    f"no CLIP vision weights (vision_model.*) found under {ckpt_dir!r}: expected model.safetensors / "
                    f"pytorch_model.bin of a CLIPModel or CLIPVisionModel checkpoint")
            missing, unexpected = self.vision_tower.load_state_dict(sd, strict=False)
            missing = [k for k in missing if "position_ids" not in k]
            if missing:
                raise RuntimeError(f"CLIP checkpoint is missing tensors: {missing[:8]} ...")
        self.vision_tower.requires_grad_(False)
        self.is_loaded = True
        if self._owner is not None and self._owner() is not None:
            self._owner().invalidate_engine()

    @torch.no_grad()
    def forward(self, images):
        """ref clip_encoder.py:39-51: [B,3,H,W] (or list of [3,H,W]) -> hidden_states[select_layer][:, 1:]."""
        model = self._owner() if self._owner is not None else None
        if model is None:
            raise RuntimeError("CLIPVisionTower is not attached to a LlavaLlamaForCausalLM; no PyTorch fallback exists")
        engine = model._ensure_engine()
        if type(images) is list:
            return [engine.vit_encode(im.unsqueeze(0)).to(im.dtype) for im in images]
        return engine.vit_encode(images).to(images.dtype)

    @property
    def dummy_feature(self):
        return torch.zeros(1, self.hidden_size, device=self.device, dtype=self.dtype)

    @property
    def dtype(self):
        return self.vision_tower.dtype

    @property
    def device(self):
        return self.vision_tower.device

    @property
    def config(self):
        return self.vision_tower.config if self.is_loaded else self.cfg_only

    @property
    def hidden_size(self):
        return self.config.hidden_size

    @property
    def num_patches(self):
        return (self.config.image_size // self.config.patch_size) ** 2
