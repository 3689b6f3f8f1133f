"""`llava.model` on the B200 engine: the LLaMA/Vicuna LLaVA classes only (the reference also exports an MPT
variant, llava/model/__init__.py:2, which is outside the north-star path)."""
from .language_model.llava_llama import LlavaConfig, LlavaLlamaForCausalLM

__all__ = ["LlavaConfig", "LlavaLlamaForCausalLM"]
