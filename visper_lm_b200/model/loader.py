"""load_pretrained_model — ola_vlm/model/builder.py:26-191 on this package's classes."""
from __future__ import annotations

import json
import os


def load_pretrained_model(model_path, model_base=None, model_name=None, load_8bit=False, load_4bit=False,
                          device_map="auto", device="cuda", use_flash_attn=False, tokenizer=None, **kwargs):
    """→ (tokenizer, model, image_processor, context_len), as builder.py:26,184-191.

    model_path: a directory written by LLaVATrainer._save / save_pretrained (config.json + safetensors).  The
    class comes from config.json's `model_type` ("ola_llama" / "ola_phi3" / "llava_llama" / "llava_phi3",
    ola_llama.py:47, llava_llama.py:40), else from its `family` / aux-head keys, else from the name like the
    reference ('phi' in it → Phi-3).
    LoRA (`model_base`), 4/8-bit loading are not built (no shipped training script produces them)."""
    from ola_vlm.constants import DEFAULT_IM_END_TOKEN, DEFAULT_IM_START_TOKEN, DEFAULT_IMAGE_PATCH_TOKEN

    from . import (LlavaLlamaForCausalLM, LlavaPhi3ForCausalLM, OlaLlavaLlamaForCausalLM, OlaLlavaPhi3ForCausalLM)

    if model_base is not None or load_8bit or load_4bit:
        raise NotImplementedError("LoRA / quantised loading is outside the built path")
    model_name = model_name or os.path.basename(os.path.normpath(str(model_path)))
    with open(os.path.join(model_path, "config.json")) as fh:
        d = json.load(fh)
    by_type = {"ola_llama": OlaLlavaLlamaForCausalLM, "ola_phi3": OlaLlavaPhi3ForCausalLM,
               "llava_llama": LlavaLlamaForCausalLM, "llava_phi3": LlavaPhi3ForCausalLM}
    if d.get("model_type") in by_type:
        cls = by_type[d["model_type"]]
    else:   # a generic config: the decoder family and the presence of the aux-head keys decide, then the name
        fam = d.get("family")
        phi = fam == "phi3" if fam else "phi" in model_name.lower()
        ola = ("aux_mode" in d or "image_depth" in d) if fam else "ola" in model_name.lower()
        cls = (OlaLlavaPhi3ForCausalLM if phi else OlaLlavaLlamaForCausalLM) if ola else \
              (LlavaPhi3ForCausalLM if phi else LlavaLlamaForCausalLM)
    if tokenizer is None:
        from transformers import AutoTokenizer

        tokenizer = AutoTokenizer.from_pretrained(model_path, use_fast=False)
    dev = device if device_map == "auto" or not isinstance(device_map, str) else device_map
    model = cls.from_pretrained(str(model_path), device=dev, **kwargs)
    tower = model.get_vision_tower()
    if not tower.is_loaded:
        tower.load_model(device_map=device_map)
    if getattr(model.config, "mm_use_im_patch_token", True):                       # builder.py:166-171
        tokenizer.add_tokens([DEFAULT_IMAGE_PATCH_TOKEN], special_tokens=True)
    if getattr(model.config, "mm_use_im_start_end", False):
        tokenizer.add_tokens([DEFAULT_IM_START_TOKEN, DEFAULT_IM_END_TOKEN], special_tokens=True)
    if len(tokenizer) > model.get_input_embeddings().weight.shape[0]:
        model.resize_token_embeddings(len(tokenizer))
    context_len = getattr(model.config, "max_sequence_length", 4096)             # builder.py:186-189
    return tokenizer, model, tower.image_processor, context_len
