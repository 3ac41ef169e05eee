"""Drop-in path of ola_vlm/train/llava_trainer.py: trainer, sampler helpers and adapter-state helper."""
from visper_lm_b200.train.checkpoint import get_mm_adapter_state as get_mm_adapter_state_maybe_zero_3  # noqa: F401
from visper_lm_b200.train.data import (LengthGroupedSampler, get_length_grouped_indices,  # noqa: F401
                                       get_modality_length_grouped_indices, split_to_even_chunks)
from visper_lm_b200.train.trainer import LLaVATrainer, TrainingArguments  # noqa: F401
