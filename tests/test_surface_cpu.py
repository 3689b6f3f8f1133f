"""CPU: the reference-facing class surface that can be checked without a GPU — factories, error behaviour,
state-dict key layout, tower properties — so that `llava/model/builder.py`-style callers find what they expect. Anything
that would compute must raise (there is no CPU path)."""
import types

import pytest
import torch

from helpers import make_llava_config, write_clip_config_dir
from oracle import llava_oracle as O


@pytest.fixture(scope="module")
def model():
    from llava.model import LlavaLlamaForCausalLM

    cfg = O.CONFIGS["tiny"]
    m = LlavaLlamaForCausalLM(make_llava_config(cfg, write_clip_config_dir(cfg)), device="cpu")
    return cfg, m


def test_build_vision_tower_contract():
    """ref multimodal_encoder/builder.py:5-11: existing local path or openai*/laion* hub id, else ValueError."""
    from llava.model.multimodal_encoder.builder import build_vision_tower

    cfg = O.CONFIGS["tiny"]
    d = write_clip_config_dir(cfg)
    ok = types.SimpleNamespace(mm_vision_tower=d, mm_vision_select_layer=-2, mm_vision_select_feature="patch")
    tower = build_vision_tower(ok, delay_load=True)
    assert not tower.is_loaded and tower.num_patches == (cfg["image_size"] // cfg["patch_size"]) ** 2
    legacy = types.SimpleNamespace(vision_tower=d, mm_vision_select_layer=-2)  # `vision_tower` fallback attribute
    assert build_vision_tower(legacy, delay_load=True).hidden_size == cfg["vit_hidden"]
    for bad in ("not/a/real/tower", None):
        with pytest.raises(ValueError, match="Unknown vision tower"):
            build_vision_tower(types.SimpleNamespace(mm_vision_tower=bad))


def test_build_vision_projector_contract():
    """ref multimodal_projector/builder.py:33-51: mlp2x_gelu -> Sequential(Linear, GELU, Linear) under keys 0.* / 2.*;
    the types the B200 path does not implement raise NotImplementedError, an unknown type ValueError (as the reference)."""
    from llava.model.multimodal_projector.builder import build_vision_projector

    c = types.SimpleNamespace(mm_projector_type="mlp2x_gelu", mm_hidden_size=64, hidden_size=128)
    p = build_vision_projector(c)
    assert sorted(p.state_dict()) == ["0.bias", "0.weight", "2.bias", "2.weight"]
    assert p.state_dict()["0.weight"].shape == (128, 64) and p.state_dict()["2.weight"].shape == (128, 128)
    with pytest.raises(RuntimeError):  # parameter holder only: no PyTorch fallback
        p(torch.zeros(1, 64))
    for t in ("linear", "identity", "mlp3x_gelu"):
        with pytest.raises(NotImplementedError):
            build_vision_projector(types.SimpleNamespace(mm_projector_type=t, mm_hidden_size=64, hidden_size=128))
    with pytest.raises(ValueError, match="Unknown projector type"):
        build_vision_projector(types.SimpleNamespace(mm_projector_type="conv", mm_hidden_size=64, hidden_size=128))


def test_state_dict_layout_and_tower_properties(model):
    """The checkpoint key layout of the reference (oracle.weight_shapes is asserted against the unmodified
    reference's own state_dict in oracle/ref_shim.py) and the CLIPVisionTower properties model_worker / builder.py read."""
    cfg, m = model
    vt = m.get_vision_tower()
    assert vt is m.get_model().get_vision_tower() and not vt.is_loaded
    assert not any("vision_tower" in k for k in m.state_dict()), "tower tensors appear only after load_model() (delay_load)"
    vt.load_model(random_init=True)
    assert vt.is_loaded
    want = {k: tuple(s) for k, s, _ in O.weight_shapes(cfg)}
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == want
    assert vt.num_patches == 16 and vt.hidden_size == cfg["vit_hidden"] and vt.dummy_feature.shape == (1, cfg["vit_hidden"])
    assert vt.config.image_size == cfg["image_size"] and vt.device.type == "cpu" and vt.dtype == torch.float32
    assert hasattr(vt, "image_processor") and hasattr(m, "lm_head") and m.get_model().embed_tokens.weight.shape[0] == cfg["vocab"]
    assert m.config.mm_projector_type == "mlp2x_gelu" and m.config.model_type == "llava"


def test_no_compute_without_a_gpu(model):
    _, m = model
    if torch.cuda.is_available():
        pytest.skip("checks the no-GPU failure mode")
    ids = torch.tensor([[1, 5, 6]])
    for call in (lambda: m(input_ids=ids), lambda: m.generate(ids, max_new_tokens=2),
                 lambda: m.encode_images(torch.zeros(1, 3, 56, 56))):
        with pytest.raises(RuntimeError):
            call()
    with pytest.raises(NotImplementedError):
        m.generate(ids, num_beams=4, max_new_tokens=2)
