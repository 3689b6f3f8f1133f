"""Real-checkpoint ingestion (VERDICT r1 missing #5 / row f3): a synthetic HF-format `llava` checkpoint
directory — config.json whose `mm_vision_tower` is a HUB ID, SHARDED safetensors + index, a sentencepiece
tokenizer.model — goes through the reference's loader contract `load_pretrained_model(model_path, None, model_name)`
(/root/reference/llava/model/builder.py:26-151) and generates; the tower resolves through the local HuggingFace cache like
the reference's `from_pretrained(vision_tower_name)` (llava/model/multimodal_encoder/clip_encoder.py:22-27).

CPU part: directory layout, shard reading, hub-id resolution, failure modes. GPU part: load -> generate equals a model
built directly from the same tensors (run in a subprocess: the HF cache location is read at import time)."""
import json
import os
import subprocess
import sys
import textwrap

import pytest
import torch

from oracle import llava_oracle as O

HUB_ID = "openai/clip-b2test-patch14-56"  # accepted by build_vision_tower like openai/clip-vit-large-patch14-336


def write_fake_hf_cache(hf_home, cfg, weights):
    """$HF_HOME/hub/models--openai--.../snapshots/<rev>/{config.json, preprocessor_config.json, model.safetensors}"""
    from safetensors.torch import save_file
    from helpers import write_clip_config_dir

    rev = "0123456789abcdef0123456789abcdef01234567"
    root = os.path.join(hf_home, "hub", "models--" + HUB_ID.replace("/", "--"))
    snap = os.path.join(root, "snapshots", rev)
    os.makedirs(snap, exist_ok=True)
    os.makedirs(os.path.join(root, "refs"), exist_ok=True)
    with open(os.path.join(root, "refs", "main"), "w") as f:
        f.write(rev)
    write_clip_config_dir(cfg, snap)
    # a CLIPModel checkpoint carries text-side tensors too: the loader must pick vision_model.* only
    sd = {k[len(O.VT) - len("vision_model."):]: v.to(torch.float16).contiguous() for k, v in weights.items() if k.startswith(O.VT)}
    sd["text_model.embeddings.token_embedding.weight"] = torch.zeros(8, 4, dtype=torch.float16)
    sd["logit_scale"] = torch.ones((), dtype=torch.float16)
    save_file(sd, os.path.join(snap, "model.safetensors"))
    return snap


def write_tokenizer(path, vocab_size):
    import random

    import sentencepiece as spm

    words = ["the", "image", "shows", "a", "cat", "dog", "on", "mat", "what", "is", "in", "picture", "user", "assistant",
             "describe", "red", "blue", "green", "sitting", "running"]
    rnd = random.Random(0)
    lines = [" ".join(rnd.choice(words) for _ in range(12)) for _ in range(2000)]
    spm.SentencePieceTrainer.train(sentence_iterator=iter(lines), model_prefix=os.path.join(path, "tokenizer"),
                                   vocab_size=vocab_size, model_type="bpe", unk_id=0, bos_id=1, eos_id=2, pad_id=-1,
                                   byte_fallback=True, character_coverage=1.0, minloglevel=2)
    os.remove(os.path.join(path, "tokenizer.vocab"))
    with open(os.path.join(path, "tokenizer_config.json"), "w") as f:
        json.dump({"tokenizer_class": "LlamaTokenizer", "bos_token": "<s>", "eos_token": "</s>", "unk_token": "<unk>",
                   "model_max_length": 2048, "legacy": True, "add_bos_token": True, "add_eos_token": False}, f)


def write_llava_checkpoint(path, cfg, weights, shards=3):
    """llava-v1.5-style directory: LlavaConfig fields the reference persists, fp16 tensors in `shards` files."""
    from safetensors.torch import save_file

    os.makedirs(path, exist_ok=True)
    conf = dict(model_type="llava", architectures=["LlavaLlamaForCausalLM"], vocab_size=cfg["vocab"], hidden_size=cfg["hidden"],
                intermediate_size=cfg["inter"], num_hidden_layers=cfg["layers"], num_attention_heads=cfg["heads"],
                num_key_value_heads=cfg["heads"], max_position_embeddings=4096, rms_norm_eps=cfg["rms_eps"], rope_theta=cfg["rope_theta"],
                hidden_act="silu", bos_token_id=1, eos_token_id=2, pad_token_id=0, torch_dtype="float16", tie_word_embeddings=False,
                mm_vision_tower=HUB_ID, mm_hidden_size=cfg["vit_hidden"], mm_projector_type="mlp2x_gelu",
                mm_vision_select_layer=cfg["select_layer"], mm_vision_select_feature="patch", mm_use_im_start_end=False,
                mm_use_im_patch_token=False, use_mm_proj=True, image_aspect_ratio="pad", tune_mm_mlp_adapter=False,
                freeze_mm_mlp_adapter=False, max_sequence_length=2048)
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(conf, f)
    keys = [k for k in weights if not k.startswith(O.VT)]  # the tower lives in the separate CLIP checkpoint (ref llava_arch.py:33)
    index = {"metadata": {"total_size": 0}, "weight_map": {}}
    for s in range(shards):
        name = f"model-{s + 1:05d}-of-{shards:05d}.safetensors"
        part = {k: weights[k].to(torch.float16).contiguous() for k in keys[s::shards]}
        part.update({f"model.layers.{i}.self_attn.rotary_emb.inv_freq": torch.ones(4) for i in range(cfg["layers"]) if s == 0})
        save_file(part, os.path.join(path, name))
        index["weight_map"].update({k: name for k in part})
    with open(os.path.join(path, "model.safetensors.index.json"), "w") as f:
        json.dump(index, f)
    write_tokenizer(path, cfg["vocab"])


def test_checkpoint_directory_layout_and_shard_reading(tmp_path):
    from llava.model.multimodal_encoder.clip_encoder import _read_checkpoint_dir

    cfg = dict(O.CONFIGS["tiny"], vocab=320)
    w = O.make_weights(cfg, seed=4)
    ck = str(tmp_path / "llava-b2test-7b")
    write_llava_checkpoint(ck, cfg, w)
    sd = _read_checkpoint_dir(ck)
    want = [k for k in w if not k.startswith(O.VT)]
    assert all(k in sd for k in want) and len([f for f in os.listdir(ck) if f.endswith(".safetensors")]) == 3
    for k in want[::7]:
        assert torch.equal(sd[k], w[k].to(torch.float16))
    from transformers import AutoTokenizer

    tok = AutoTokenizer.from_pretrained(ck, use_fast=False)
    assert len(tok) == cfg["vocab"] and tok.bos_token_id == 1 and tok.eos_token_id == 2


_RESOLVE = textwrap.dedent("""
    import sys, torch
    sys.path[:0] = [sys.argv[1], sys.argv[2], sys.argv[2] + "/tests"]
    from types import SimpleNamespace
    from llava.model.multimodal_encoder.builder import build_vision_tower
    from llava.model.multimodal_encoder.clip_encoder import _resolve_checkpoint_dir
    hub_id = sys.argv[3]
    args = SimpleNamespace(mm_vision_tower=hub_id, mm_vision_select_layer=-2, mm_vision_select_feature="patch")
    tower = build_vision_tower(args, delay_load=True)               # config only, from the cache
    assert tower.config.image_size == 56 and not tower.is_loaded
    tower.load_model()                                              # weights from <cache>/snapshots/<rev>/model.safetensors
    sd = tower.vision_tower.state_dict()
    assert tower.is_loaded and tower.image_processor.crop_size["height"] == 56
    assert not any(k.startswith("text_model") for k in sd) and float(sd["vision_model.pre_layrnorm.weight"].abs().sum()) > 0
    try:
        _resolve_checkpoint_dir("openai/not-in-the-cache")
    except RuntimeError as e:
        assert "never downloads" in str(e)
    else:
        raise SystemExit("missing cache entry did not raise")
    print("resolve-ok")
""")


def test_hub_id_tower_resolves_through_the_local_hf_cache(tmp_path, repo_root):
    cfg = dict(O.CONFIGS["tiny"], vocab=320)
    w = O.make_weights(cfg, seed=4)
    hf_home = str(tmp_path / "hf_home")
    write_fake_hf_cache(hf_home, cfg, w)
    env = dict(os.environ, HF_HOME=hf_home, HF_HUB_OFFLINE="1", TRANSFORMERS_OFFLINE="1")
    env.pop("HF_HUB_CACHE", None)
    r = subprocess.run([sys.executable, "-c", _RESOLVE, os.path.join(repo_root, "llava-plus-codebase_b200"), repo_root, HUB_ID],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "resolve-ok" in r.stdout, r.stderr[-3000:]


_LOAD_AND_GENERATE = textwrap.dedent("""
    import sys, torch
    sys.path[:0] = [sys.argv[1], sys.argv[2], sys.argv[2] + "/tests"]
    from llava.model.builder import load_pretrained_model
    from llava.constants import IMAGE_TOKEN_INDEX
    from oracle import llava_oracle as O
    from helpers import make_model
    ck = sys.argv[3]
    tokenizer, model, image_processor, context_len = load_pretrained_model(ck, None, "llava-b2test-7b")
    assert context_len == 2048 and image_processor.crop_size["height"] == 56
    assert model.get_vision_tower().is_loaded and model.device.type == "cuda"
    a, b = tokenizer("what is in the").input_ids, tokenizer("picture").input_ids[1:]      # [bos ...] + <image> + ...
    ids = torch.tensor([a + [IMAGE_TOKEN_INDEX] + b])
    from PIL import Image
    img = Image.new("RGB", (80, 60), (200, 30, 90))
    pixels = image_processor.preprocess(img, return_tensors="pt")["pixel_values"].half().cuda()   # callers cast to fp16 (model_worker.py:139-141)
    out = model.generate(ids.cuda(), images=pixels, do_sample=False, max_new_tokens=12, eos_token_id=[])
    # the same tensors, loaded directly
    cfg = dict(O.CONFIGS["tiny"], vocab=320)
    w = {k: v.to(torch.float16).to(torch.bfloat16) for k, v in O.make_weights(cfg, seed=4).items()}
    direct = make_model(cfg, w, max_batch=1, max_seq=256)
    want = direct.generate(ids.cuda(), images=pixels, do_sample=False, max_new_tokens=12, eos_token_id=[])
    assert torch.equal(out.cpu(), want.cpu()), (out, want)
    text = tokenizer.decode(out[0, ids.shape[1]:], skip_special_tokens=True)
    print("load-generate-ok", repr(text))
""")


@pytest.mark.gpu
def test_load_pretrained_model_from_checkpoint_dir_and_generate(tmp_path, repo_root):
    cfg = dict(O.CONFIGS["tiny"], vocab=320)
    w = O.make_weights(cfg, seed=4)
    hf_home, ck = str(tmp_path / "hf_home"), str(tmp_path / "llava-b2test-7b")
    write_fake_hf_cache(hf_home, cfg, w)
    write_llava_checkpoint(ck, cfg, w)
    env = dict(os.environ, HF_HOME=hf_home, HF_HUB_OFFLINE="1", TRANSFORMERS_OFFLINE="1")
    env.pop("HF_HUB_CACHE", None)
    r = subprocess.run([sys.executable, "-c", _LOAD_AND_GENERATE, os.path.join(repo_root, "llava-plus-codebase_b200"), repo_root, ck],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "load-generate-ok" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])
