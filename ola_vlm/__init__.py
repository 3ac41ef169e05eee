"""Drop-in module path: `import ola_vlm` resolves to the B200-native hot path (visper_lm_b200).
Only the training-step surface of the reference package exists here."""
from .model import LlavaLlamaForCausalLM, LlavaPhi3ForCausalLM  # noqa: F401
