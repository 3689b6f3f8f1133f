"""ORACLE SUPPORT — TEST INFRASTRUCTURE ONLY. Imports the UNMODIFIED reference hot-path files from
/root/reference (build container only; the directory does not exist on the GPU box) on top of the installed
transformers 5.5.0, following the recipe probed in SURVEY.md §8c / Appendix E:

  (1) AutoConfig / AutoModelForCausalLM `register(..., exist_ok=True)` (llava_llama.py:110-111 collides with
      HF's built-in "llava" type),
  (2) namespace stubs for `llava`, `llava.model`, `llava.model.language_model` so their __init__.py (which
      import the broken MPT branch) never run,
  (3) a local random-init CLIP checkpoint directory so build_vision_tower's os.path.exists branch is taken,
  (4) eager attention to follow the pinned-version math.

Used by tests/golden/make_golden.py to generate the committed fixtures that pin oracle/llava_oracle.py.
Nothing here is copied from the reference: its files are executed where they lie.
"""
import json
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("LLAVA_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "llava", "model"))


def import_reference():
    """Returns (LlavaLlamaForCausalLM, LlavaConfig) classes defined by the reference's own files."""
    from transformers import AutoConfig, AutoModelForCausalLM

    if not getattr(AutoConfig, "_b2_patched", False):
        _r = AutoConfig.register
        AutoConfig.register = staticmethod(lambda mt, cfg, exist_ok=False: _r(mt, cfg, exist_ok=True))
        _m = AutoModelForCausalLM.register.__func__
        AutoModelForCausalLM.register = classmethod(lambda cls, c, m, exist_ok=False: _m(cls, c, m, exist_ok=True))
        AutoConfig._b2_patched = True
    R = os.path.join(REFERENCE_ROOT, "llava")
    for name, path in (("llava", R), ("llava.model", R + "/model"), ("llava.model.language_model", R + "/model/language_model")):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
    from llava.model.language_model.llava_llama import LlavaLlamaForCausalLM, LlavaConfig  # noqa: E402
    return LlavaLlamaForCausalLM, LlavaConfig


def unimport_reference():
    for k in [k for k in sys.modules if k == "llava" or k.startswith("llava.")]:
        del sys.modules[k]


def write_clip_dir(path, cfg):
    """Config-only CLIP directory (+ preprocessor) for build_vision_tower / CLIPVisionTower.load_model."""
    from transformers import CLIPVisionConfig, CLIPVisionModel

    os.makedirs(path, exist_ok=True)
    vc = CLIPVisionConfig(hidden_size=cfg["vit_hidden"], intermediate_size=cfg["vit_inter"],
                          num_hidden_layers=cfg["vit_layers"], num_attention_heads=cfg["vit_heads"],
                          image_size=cfg["image_size"], patch_size=cfg["patch_size"], projection_dim=64,
                          layer_norm_eps=cfg["vit_eps"], hidden_act="quick_gelu")
    CLIPVisionModel(vc).save_pretrained(path)
    with open(os.path.join(path, "preprocessor_config.json"), "w") as f:
        json.dump({"crop_size": cfg["image_size"], "do_center_crop": True, "do_normalize": True, "do_resize": True,
                   "image_mean": [0.48145466, 0.4578275, 0.40821073],
                   "image_std": [0.26862954, 0.26130258, 0.27577711], "resample": 3, "size": cfg["image_size"]}, f)
    return path


def build_reference_model(cfg, weights, clip_dir, dtype=torch.float32, **extra_cfg):
    """Reference LlavaLlamaForCausalLM at `cfg` dims carrying `weights` (oracle.make_weights layout)."""
    Model, Config = import_reference()
    write_clip_dir(clip_dir, cfg)
    hf_cfg = Config(vocab_size=cfg["vocab"], hidden_size=cfg["hidden"], intermediate_size=cfg["inter"],
                    num_hidden_layers=cfg["layers"], num_attention_heads=cfg["heads"],
                    num_key_value_heads=cfg["heads"], max_position_embeddings=4096, rms_norm_eps=cfg["rms_eps"],
                    rope_theta=cfg["rope_theta"], mm_vision_tower=clip_dir, mm_hidden_size=cfg["vit_hidden"],
                    mm_projector_type="mlp2x_gelu", mm_vision_select_layer=cfg["select_layer"],
                    mm_vision_select_feature="patch", mm_use_im_start_end=False, mm_use_im_patch_token=False,
                    use_mm_proj=True, attn_implementation="eager", **extra_cfg)
    model = Model(hf_cfg)
    model.get_vision_tower().load_model()
    vt = model.get_vision_tower().vision_tower
    try:
        vt.config._attn_implementation = "eager"
        vt.vision_model.config._attn_implementation = "eager"
    except Exception:
        pass
    sd = model.state_dict()
    missing = [k for k in sd if k not in weights and "position_ids" not in k and "inv_freq" not in k]
    assert not missing, f"oracle weight layout is missing reference keys: {missing[:5]}"
    model.load_state_dict({k: v for k, v in weights.items() if k in sd}, strict=False)
    return model.to(dtype).eval()
