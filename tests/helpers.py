"""Shared test helpers (test infrastructure): build engines / models from oracle weights, error metrics."""
import json
import os
import tempfile

import torch

from oracle import llava_oracle as O


def desc_from_cfg(cfg, max_batch=2, max_seq=128, max_images=4):
    return dict(image_size=cfg["image_size"], patch_size=cfg["patch_size"], vit_hidden=cfg["vit_hidden"],
                vit_inter=cfg["vit_inter"], vit_layers=cfg["vit_layers"], vit_heads=cfg["vit_heads"],
                vit_select_layer=cfg["select_layer"], vit_ln_eps=cfg["vit_eps"], hidden=cfg["hidden"],
                inter=cfg["inter"], layers=cfg["layers"], heads=cfg["heads"], vocab=cfg["vocab"],
                rms_eps=cfg["rms_eps"], rope_theta=cfg["rope_theta"], max_batch=max_batch, max_seq=max_seq,
                max_images=max_images)


def make_engine(cfg, weights, device="cuda", **limits):
    from llava._b2 import Engine

    eng = Engine(desc_from_cfg(cfg, **limits), device)
    for k, v in weights.items():
        eng.set_weight(k, v.to(device=device, dtype=torch.bfloat16))
    eng.finalize()
    return eng


def write_clip_config_dir(cfg, path=None):
    """Config-only CLIP directory for build_vision_tower (no weights: tests use load_model(random_init=True))."""
    from transformers import CLIPVisionConfig

    path = path or tempfile.mkdtemp(prefix="b2clip_")
    CLIPVisionConfig(hidden_size=cfg["vit_hidden"], intermediate_size=cfg["vit_inter"],
                     num_hidden_layers=cfg["vit_layers"], num_attention_heads=cfg["vit_heads"],
                     image_size=cfg["image_size"], patch_size=cfg["patch_size"], projection_dim=64,
                     layer_norm_eps=cfg["vit_eps"], hidden_act="quick_gelu").save_pretrained(path)
    with open(os.path.join(path, "preprocessor_config.json"), "w") as f:
        json.dump({"crop_size": cfg["image_size"], "do_center_crop": True, "do_normalize": True, "do_resize": True,
                   "image_mean": [0.48145466, 0.4578275, 0.40821073],
                   "image_std": [0.26862954, 0.26130258, 0.27577711], "resample": 3, "size": cfg["image_size"],
                   "image_processor_type": "CLIPImageProcessor"}, f)
    return path


def make_llava_config(cfg, clip_dir, **extra):
    from llava.model import LlavaConfig

    return LlavaConfig(vocab_size=cfg["vocab"], hidden_size=cfg["hidden"], intermediate_size=cfg["inter"],
                       num_hidden_layers=cfg["layers"], num_attention_heads=cfg["heads"],
                       num_key_value_heads=cfg["heads"], max_position_embeddings=4096, rms_norm_eps=cfg["rms_eps"],
                       rope_theta=cfg["rope_theta"], mm_vision_tower=clip_dir, mm_hidden_size=cfg["vit_hidden"],
                       mm_projector_type="mlp2x_gelu", mm_vision_select_layer=cfg["select_layer"],
                       mm_vision_select_feature="patch", mm_use_im_start_end=False, mm_use_im_patch_token=False,
                       use_mm_proj=True, **extra)


def make_model(cfg, weights, device="cuda", max_batch=4, max_seq=256, max_images=4, **extra_cfg):
    """LlavaLlamaForCausalLM (this repo's class) carrying oracle weights, through the public API."""
    from llava.model import LlavaLlamaForCausalLM

    clip_dir = write_clip_config_dir(cfg)
    model = LlavaLlamaForCausalLM(make_llava_config(cfg, clip_dir, **extra_cfg), device=device,
                                  max_batch=max_batch, max_seq=max_seq, max_images=max_images)
    model.get_vision_tower().load_model(random_init=True)
    model.to(device=device, dtype=torch.bfloat16)
    sd = model.state_dict()
    missing = [k for k in sd if k not in weights]
    assert not missing, missing[:5]
    model.load_state_dict({k: v for k, v in weights.items() if k in sd}, strict=True)
    return model.eval()


def rel_err(a, b):
    """max-abs and mean-abs error of a vs reference b, normalised by std(b)."""
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    s = float(b.std()) + 1e-12
    d = (a - b).abs()
    return float(d.max()) / s, float(d.mean()) / s


def synth_inputs(cfg, B, Lt, seed=1, image_pos=5):
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(B, 3, cfg["image_size"], cfg["image_size"], generator=g)
    ids = torch.randint(3, cfg["vocab"], (B, Lt), generator=g)
    ids[:, 0] = 1
    ids[:, image_pos] = O.IMAGE_TOKEN_INDEX
    return ids, images
