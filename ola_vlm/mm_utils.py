"""Drop-in path of ola_vlm/mm_utils.py:336-355."""
from visper_lm_b200.train.data import tokenizer_image_token  # noqa: F401
