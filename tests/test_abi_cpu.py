"""CPU: the C-ABI library loads and exports every symbol include/b2llava.h declares; the binding covers them;
the product path fails loudly without a GPU (no CPU / oracle fallback)."""
import os
import re

import pytest
import torch


def _header_symbols(root):
    src = open(os.path.join(root, "include", "b2llava.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(repo_root):
    from llava import _b2

    syms = _header_symbols(repo_root)
    assert len(syms) >= 25
    lib = _b2.load_library()
    for s in syms:
        assert hasattr(lib, s), f"libb2llava.so does not export {s}"
        assert s in _b2.SIGNATURES, f"ctypes binding has no signature for {s}"
    assert set(_b2.SIGNATURES) == set(syms), "binding declares symbols the header does not"
    assert lib.b2_version() >= 1
    assert isinstance(_b2.last_error(), str)


def test_modeldesc_matches_header_struct(repo_root):
    from llava import _b2

    src = open(os.path.join(repo_root, "include", "b2llava.h")).read()
    body = src[src.index("typedef struct b2_model_desc {"):src.index("} b2_model_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\b(?:int32_t|float)\s+([a-z_]+)\s*;", body)
    assert fields == [f[0] for f in _b2.ModelDesc._fields_]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from llava import _b2

    with pytest.raises(RuntimeError):
        _b2.Engine(dict(image_size=56, patch_size=14, vit_hidden=256, vit_inter=512, vit_layers=3, vit_heads=4,
                        vit_select_layer=-2, vit_ln_eps=1e-5, hidden=256, inter=512, layers=2, heads=2, vocab=1024,
                        rms_eps=1e-5, rope_theta=10000.0, max_batch=1, max_seq=64, max_images=1), "cpu")
    with pytest.raises(RuntimeError):
        _b2.init(0)


def test_product_never_imports_oracle(repo_root):
    pkg = os.path.join(repo_root, "llava-plus-codebase_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("# oracle", ""), f"{f} references the oracle"


def test_stream_k_plan_helpers_are_host_only():
    """The stream-K GEMM's scratch sizing is pure host arithmetic (no GPU needed): the workspace holds, per 128-row
    weight tile, one [BN x 128] fp32 slot for every CTA that can share the tile; BN is the batch rounded up to 32/64/128."""
    from llava import _b2

    lib = _b2.load_library()
    slot = lambda bn: bn * 128 * 4
    for B, N, K in [(7, 4096, 4096), (32, 12288, 4096), (33, 4096, 11008), (128, 32000, 4096), (20, 200, 264)]:
        ws = lib.b2_op_gemm_skinny_workspace_bytes(B, N, K)
        bn = 32 if B <= 32 else (64 if B <= 64 else 128)
        tiles = (N + 127) // 128
        assert ws > 0 and ws % (tiles * slot(bn)) == 0, (B, N, K, ws)
        segs = ws // (tiles * slot(bn))
        assert 2 <= segs <= (K + 63) // 64 + 1  # at least own slot + one neighbour, never more than the k-blocks of a tile
        assert lib.b2_op_gemm_skinny_counter_bytes(N) == tiles * 4
    assert lib.b2_op_gemm_skinny_workspace_bytes(0, 128, 64) == -1
    assert lib.b2_op_gemm_skinny_workspace_bytes(129, 128, 64) == -1
    assert lib.b2_op_gemm_skinny_counter_bytes(0) == -1
