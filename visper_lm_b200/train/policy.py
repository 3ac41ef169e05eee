"""Which parameters train — the requires_grad bookkeeping that train() does inline in the reference
(ola_vlm/train/ola_vlm_train.py:1127-1136 adapter tuning / freezing, :1145 tower always frozen, :1239-1241
frozen teachers and DPT decoder, :1244-1258 task tokens), as one function over parameter NAMES so it does
not depend on the order in which the modules were created."""
from __future__ import annotations

FROZEN_ALWAYS = ("vision_tower.", "da_v2_head.", "gen_encoder.", "dav2_backbone.", "oneformer.")
ADDED_BY_DISTILLATION = ("image_gen_heads.", "image_depth_heads.", "image_seg_heads.", "_logit_scale")


def apply_freeze_policy(model, tune_mm_mlp_adapter=False, freeze_mm_mlp_adapter=False, freeze_task_token=False,
                        freeze_backbone=False):
    """Sets requires_grad on every parameter; returns the sorted list of trainable names.

    tune_mm_mlp_adapter (pretrain.sh): the LLM is frozen (`model.requires_grad_(False)`), the projector
    trains, and everything train() creates AFTERWARDS — heads, logit scales, task tokens — keeps its
    default requires_grad=True.  Otherwise (finetune.sh) everything trains except the always-frozen
    modules.  freeze_mm_mlp_adapter / freeze_task_token / freeze_backbone as in the reference."""
    names = []
    for n, p in model.named_parameters():
        if any(t in n for t in FROZEN_ALWAYS):
            on = False
        elif "mm_projector." in n:
            on = not freeze_mm_mlp_adapter
        elif "model.special_" in n:
            on = not freeze_task_token
        elif any(t in n for t in ADDED_BY_DISTILLATION):
            on = True
        else:  # the language model itself (embed_tokens, layers, norm, lm_head)
            # --freeze_backbone is `model.model.requires_grad_(False)` (ola_vlm_train.py:1043-1044): the decoder
            # under `model.`, not lm_head
            on = not (tune_mm_mlp_adapter or (freeze_backbone and n.startswith("model.")))
        p.requires_grad_(on)
        if on:
            names.append(n)
    return sorted(names)
