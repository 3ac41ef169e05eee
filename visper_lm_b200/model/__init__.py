from .vlm import (LlavaConfig, LlavaLlamaForCausalLM, LlavaPhi3Config, LlavaPhi3ForCausalLM,  # noqa: F401
                  OlaCausalLLMOutputWithPast, OlaLlavaLlamaConfig, OlaLlavaLlamaForCausalLM,
                  OlaLlavaPhi3Config, OlaLlavaPhi3ForCausalLM, VisperConfig)
