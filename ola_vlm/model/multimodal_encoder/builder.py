"""Drop-in path of ola_vlm/model/multimodal_encoder/builder.py:6-17 — the tower class is picked from the name."""
from visper_lm_b200.model.convnext import CLIPConvNextVisionTower
from visper_lm_b200.model.modules import CLIPVisionTower


def build_vision_tower(vision_tower_cfg, **kwargs):
    name = getattr(vision_tower_cfg, "mm_vision_tower", getattr(vision_tower_cfg, "vision_tower", None))
    if name is not None and "clip" in name and "convnext" not in name:
        vision = getattr(vision_tower_cfg, "vision", None) or dict(
            hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16, image_size=336,
            patch_size=14)
        return CLIPVisionTower(vision, getattr(vision_tower_cfg, "mm_vision_select_layer", -2),
                               getattr(vision_tower_cfg, "mm_vision_select_feature", "patch"), kwargs.get("device"))
    if name is not None and "convnext" in name.lower():
        return CLIPConvNextVisionTower(name, args=vision_tower_cfg, device=kwargs.get("device"))
    raise ValueError(f"Unknown vision tower: {name}")
