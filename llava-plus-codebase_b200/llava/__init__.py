"""B200-native drop-in for the reference's `llava` package, hot path only.

`llava.model` (LlavaLlamaForCausalLM, LlavaConfig, build_vision_tower, build_vision_projector,
load_pretrained_model) is implemented here on top of libb2llava.so. Everything the reference keeps outside
the hot path (llava.serve, llava.eval, llava.conversation, llava.mm_utils, ...) is NOT re-implemented: set
LLAVA_REFERENCE_ROOT=/path/to/LLaVA-Plus-Codebase and those submodules resolve to the reference's own files
while `llava.model` keeps resolving to this package.
"""
import os as _os

_ref = _os.environ.get("LLAVA_REFERENCE_ROOT")
if _ref:
    _p = _os.path.join(_ref, "llava")
    if _os.path.isdir(_p) and _p not in __path__:
        __path__.append(_p)  # our directory stays first: llava.model/* is ours, the rest falls through

from .model import LlavaLlamaForCausalLM, LlavaConfig  # noqa: E402,F401
