from visper_lm_b200.model import (LlavaConfig, LlavaLlamaForCausalLM, LlavaPhi3Config,  # noqa: F401
                                  LlavaPhi3ForCausalLM, OlaLlavaLlamaConfig, OlaLlavaLlamaForCausalLM,
                                  OlaLlavaPhi3Config, OlaLlavaPhi3ForCausalLM)
