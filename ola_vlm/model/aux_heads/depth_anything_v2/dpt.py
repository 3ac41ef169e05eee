"""Drop-in path of ola_vlm/model/aux_heads/depth_anything_v2/dpt.py (imported at base_ola_vlm.py:14)."""
from visper_lm_b200.model.dinov2 import DepthAnythingV2  # noqa: F401
from visper_lm_b200.model.dpt import DPTHead  # noqa: F401
