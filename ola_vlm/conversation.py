"""Drop-in path of the chat templates the training scripts select (ola_vlm/conversation.py:225-251)."""
from visper_lm_b200.train.prompts import LLAMA3 as conv_llava_llama_3  # noqa: F401
from visper_lm_b200.train.prompts import PHI3 as conv_llava_phi_3  # noqa: F401
from visper_lm_b200.train.prompts import ChatTemplate as Conversation  # noqa: F401
from visper_lm_b200.train.prompts import conv_templates  # noqa: F401

default_conversation = conv_llava_phi_3
