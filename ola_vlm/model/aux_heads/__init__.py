"""Drop-in path of ola_vlm/model/aux_heads/__init__.py: the heads and frozen teachers this repo builds."""
from visper_lm_b200.model.dpt import DAv2_Head  # noqa: F401
from visper_lm_b200.model.modules import (OneFormerTaskTokenSegHead, TaskTokenDepthHead,  # noqa: F401
                                          TaskTokenGenHead)
from visper_lm_b200.model.seg_teacher import OneFormerHead  # noqa: F401
