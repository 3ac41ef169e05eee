"""Drop-in path of ola_vlm/model/multimodal_encoder/clip_convnext_encoder.py."""
from visper_lm_b200.model.convnext import CLIPConvNextVisionTower, extract_res_interp  # noqa: F401
