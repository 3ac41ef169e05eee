"""Drop-in path of ola_vlm/model/multimodal_encoder/clip_encoder.py."""
from visper_lm_b200.model.modules import CLIPVisionTower  # noqa: F401
