"""CPU (build container only: needs the reference tree): INTEGRATION.md §1 — with LLAVA_REFERENCE_ROOT set, the reference's
OWN consumers of the hot path (llava/serve/cli.py, llava/serve/model_worker.py, llava/eval/model_vqa_loader.py, run unmodified)
import on top of this package's `llava.model`, and the entry points they call have the reference's signatures."""
import ast
import inspect
import os
import subprocess
import sys
import textwrap

import pytest

REF = os.environ.get("LLAVA_REFERENCE_ROOT", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "llava", "serve")),
                                reason="reference tree not present (it does not travel to the GPU box)")


def test_reference_consumers_import_on_top_of_this_package(repo_root, tmp_path):
    code = textwrap.dedent("""
        import sys, types
        sys.modules.setdefault("shortuuid", types.ModuleType("shortuuid"))   # the reference's own (absent) dependency
        import llava
        import llava.mm_utils as mu, llava.conversation as conv, llava.utils as ut
        import llava.model.builder as b, llava.model.language_model.llava_llama as ll, llava.model.llava_arch as arch
        import llava.serve.cli as cli, llava.serve.model_worker as mw, llava.eval.model_vqa_loader as vqa, llava.eval.run_llava as rl
        pkg, ref = sys.argv[1], sys.argv[2]
        for m in (mu, conv, ut, cli, mw, vqa, rl):
            assert m.__file__.startswith(ref), m.__file__          # the reference's files, unmodified
        for m in (b, ll, arch):
            assert m.__file__.startswith(pkg), m.__file__          # the hot path: this repo
        assert cli.load_pretrained_model is b.load_pretrained_model and mw.load_pretrained_model is b.load_pretrained_model
        assert vqa.load_pretrained_model is b.load_pretrained_model
        assert mu.IMAGE_TOKEN_INDEX == -200
        print("ok")
    """)
    pkg = os.path.join(repo_root, "llava-plus-codebase_b200")
    env = dict(os.environ, LLAVA_REFERENCE_ROOT=REF, PYTHONPATH=pkg, TRANSFORMERS_OFFLINE="1", HF_HUB_OFFLINE="1")
    r = subprocess.run([sys.executable, "-c", code, pkg, REF], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    # the reference's build_logger (llava/utils.py:17-57, run by model_worker at import) redirects stdout into its logger
    assert r.returncode == 0 and "ok" in (r.stdout + r.stderr).splitlines()[-1], r.stderr[-2000:]


def _ref_signature(path, func, cls=None):
    tree = ast.parse(open(path).read())
    scope = tree.body
    if cls is not None:
        scope = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls).body
    fn = next(n for n in scope if isinstance(n, ast.FunctionDef) and n.name == func)
    return [a.arg for a in fn.args.args]


def test_entry_point_signatures_match_the_reference():
    from llava.model.builder import load_pretrained_model
    from llava.model.language_model.llava_llama import LlavaLlamaForCausalLM
    from llava.model.multimodal_encoder.builder import build_vision_tower
    from llava.model.multimodal_projector.builder import build_vision_projector

    R = os.path.join(REF, "llava", "model")
    ours = list(inspect.signature(load_pretrained_model).parameters)
    assert ours[: len(_ref_signature(R + "/builder.py", "load_pretrained_model"))] == _ref_signature(R + "/builder.py", "load_pretrained_model")
    ref_fwd = _ref_signature(R + "/language_model/llava_llama.py", "forward", "LlavaLlamaForCausalLM")
    assert list(inspect.signature(LlavaLlamaForCausalLM.forward).parameters) == ref_fwd
    ref_prep = _ref_signature(R + "/llava_arch.py", "prepare_inputs_labels_for_multimodal", "LlavaMetaForCausalLM")
    assert list(inspect.signature(LlavaLlamaForCausalLM.prepare_inputs_labels_for_multimodal).parameters) == ref_prep
    assert list(inspect.signature(LlavaLlamaForCausalLM.encode_images).parameters) == \
        _ref_signature(R + "/llava_arch.py", "encode_images", "LlavaMetaForCausalLM")
    assert list(inspect.signature(build_vision_tower).parameters)[0] == _ref_signature(R + "/multimodal_encoder/builder.py", "build_vision_tower")[0]
    assert list(inspect.signature(build_vision_projector).parameters)[:2] == _ref_signature(R + "/multimodal_projector/builder.py", "build_vision_projector")[:2]


def test_reference_keywords_stopping_criteria_through_the_decode_loop():
    """The reference's OWN KeywordsStoppingCriteria (llava/mm_utils.py:79-114, loaded from the reference tree, unmodified)
    driving generate()'s host loop: it must see cat(prompt ids incl. IMAGE_TOKEN_INDEX, new tokens), one column more per
    step, and stop the loop at the step whose tail spells the keyword (GPU counterpart with a restated criterion:
    tests/test_generate_gpu.py)."""
    import importlib.util

    import torch

    from llava.model.language_model.llava_llama import _stream_decode
    from test_generate_host import FakeEngine, per_step_reference

    spec = importlib.util.spec_from_file_location("ref_mm_utils", os.path.join(REF, "llava", "mm_utils.py"))
    mu = importlib.util.module_from_spec(spec)
    sys.modules.setdefault("llava.constants", __import__("llava.constants", fromlist=["x"]))
    spec.loader.exec_module(mu)

    class Tok:
        bos_token_id = 1

        def __call__(self, text):
            return type("Enc", (), {"input_ids": [1] + [ord(c) - 97 for c in text]})()

        def batch_decode(self, ids, skip_special_tokens=True):
            return ["".join(chr(97 + int(i) % 26) if int(i) >= 0 else "?" for i in row) for row in ids]

    prompt = torch.tensor([[1, 5, -200, 7, 9]])
    free = per_step_reference([4], 40, set(), 0)[0].tolist()           # what the fake model generates unconstrained
    text = "".join(chr(97 + t % 26) for t in free)
    keyword = text[6:9]
    stop_at = text.find(keyword) + 3
    crit = mu.KeywordsStoppingCriteria([keyword], Tok(), prompt)
    assert crit.start_len == 5
    out = _stream_decode(FakeEngine([4]), None, None, None, 1, 40, set(), 0, prompt, None, [crit], run_ahead=8)
    assert out[0].tolist() == free[:stop_at]
