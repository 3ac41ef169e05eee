"""Prompt templating and label masking of the supervised pipeline, host-side:

  * the two MPT-style chat templates the shipped scripts select (`--version llava_llama_3` /
    `llava_phi_3`: ola_vlm/conversation.py:225-243) and their rendering (conversation.py:32-107);
  * `preprocess_multimodal` (ola_vlm/train/ola_vlm_train.py:350-371): `<image>` moved to the front
    of the turn that carries it;
  * `preprocess_llama_3` / `preprocess_phi_3` (ola_vlm_train.py:374-548): tokenise the rendered
    conversation with `<image>` → IMAGE_TOKEN_INDEX and mask everything but the assistant turns —
    one generic routine here, the two differ only in a per-round token-count correction;
Same outputs as the reference's functions for the same tokenizer (tests/test_data_pipeline.py runs
both); the reference's consistency check (`tokenization mismatch` → the whole sample is ignored) is
kept, returning the mismatch count instead of printing.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import torch

from .data import IGNORE_INDEX, tokenizer_image_token

DEFAULT_IMAGE_TOKEN = "<image>"
DEFAULT_IM_START_TOKEN = "<im_start>"
DEFAULT_IM_END_TOKEN = "<im_end>"


@dataclass
class ChatTemplate:
    """MPT-style template: system, then every turn as role-prefix + text + separator."""
    system: str
    roles: Tuple[str, str]
    sep: str
    version: str
    round_correction: int = 0  # tokens the reference subtracts from every round after the first
    messages: List[List[str]] = field(default_factory=list)

    def copy(self):
        return ChatTemplate(self.system, self.roles, self.sep, self.version, self.round_correction,
                            [list(m) for m in self.messages])

    def append_message(self, role, message):
        self.messages.append([role, message])

    def get_prompt(self) -> str:
        out = self.system + self.sep
        for role, text in self.messages:
            out += role + text + self.sep if text else role
        return out


LLAMA3 = ChatTemplate(
    system="<|start_header_id|>system<|end_header_id|>\n\nA chat between a curious user and an artificial "
           "intelligence assistant. The assistant gives helpful, detailed, and polite answers to the user's questions.",
    roles=("<|start_header_id|>user<|end_header_id|>\n\n", "<|start_header_id|>assistant<|end_header_id|>\n\n"),
    sep="<|eot_id|>", version="llama3", round_correction=0)

PHI3 = ChatTemplate(
    system="<|system|>\nYou are a helpful AI assistant.",
    roles=("\n<|user|>\n", "\n<|assistant|>\n"),
    sep="<|end|>", version="phi3", round_correction=2)

conv_templates = {"llava_llama_3": LLAMA3, "llava_phi_3": PHI3}


def preprocess_multimodal(sources, is_multimodal=True, mm_use_im_start_end=False):
    """In place, like the reference: the image placeholder leads the turn that contains it."""
    if not is_multimodal:
        return sources
    for source in sources:
        for turn in source:
            text = turn["value"]
            if DEFAULT_IMAGE_TOKEN in text:
                text = (DEFAULT_IMAGE_TOKEN + "\n" + text.replace(DEFAULT_IMAGE_TOKEN, "").strip()).strip()
            if mm_use_im_start_end:
                text = text.replace(DEFAULT_IMAGE_TOKEN, DEFAULT_IM_START_TOKEN + DEFAULT_IMAGE_TOKEN + DEFAULT_IM_END_TOKEN)
            turn["value"] = text
    return sources


def render(sources, template: ChatTemplate) -> List[str]:
    """human/gpt turn lists → prompt strings (a leading non-human turn is dropped)."""
    name = {"human": template.roles[0], "gpt": template.roles[1]}
    prompts = []
    for i, source in enumerate(sources):
        if name[source[0]["from"]] != template.roles[0]:
            source = source[1:]
        conv = template.copy()
        conv.messages = []
        for j, turn in enumerate(source):
            role = name[turn["from"]]
            assert role == template.roles[j % 2], f"{i}"
            conv.append_message(role, turn["value"])
        prompts.append(conv.get_prompt())
    return prompts


def preprocess_mpt(sources, tokenizer, template: ChatTemplate, has_image=False) -> Dict:
    """input_ids / labels for MPT-style templates.  Labels keep only the assistant answers: the BOS,
    and in every (user, assistant) round the tokens up to and including the assistant role prefix
    (minus the reference's 2-token slack), are IGNORE_INDEX; so is everything after the last round."""
    prompts = render(sources, template)
    if has_image:
        input_ids = torch.stack([tokenizer_image_token(p, tokenizer, return_tensors="pt") for p in prompts], 0)
        count = lambda text: len(tokenizer_image_token(text, tokenizer))
    else:
        input_ids = tokenizer(prompts, return_tensors="pt", padding="longest",
                              max_length=tokenizer.model_max_length, truncation=True).input_ids
        count = lambda text: len(tokenizer(text).input_ids)
    labels = input_ids.clone()
    sep, answer_mark = template.sep, template.sep + template.roles[1]
    mismatches = 0
    for prompt, target in zip(prompts, labels):
        total = int(target.ne(tokenizer.pad_token_id).sum())
        pieces = prompt.split(sep)
        rounds = [sep.join(pieces[:3])] + [sep.join(pieces[i:i + 2]) for i in range(3, len(pieces), 2)]
        cur = 1
        target[:cur] = IGNORE_INDEX
        for r, text in enumerate(rounds):
            if not text:
                break
            parts = text.split(answer_mark)
            if len(parts) != 2:
                break
            n_round = count(text)
            n_instr = count(parts[0] + answer_mark) - 2
            if r > 0:
                n_round -= template.round_correction
                n_instr -= template.round_correction
            target[cur:cur + n_instr] = IGNORE_INDEX
            cur += n_round
        target[cur:] = IGNORE_INDEX
        if cur < tokenizer.model_max_length and cur != total:
            target[:] = IGNORE_INDEX  # "tokenization mismatch ... (ignored)"
            mismatches += 1
    return dict(input_ids=input_ids, labels=labels, mismatches=mismatches)


def preprocess_llama_3(sources, tokenizer, has_image=False):
    return preprocess_mpt(sources, tokenizer, LLAMA3, has_image)


def preprocess_phi_3(sources, tokenizer, has_image=False):
    return preprocess_mpt(sources, tokenizer, PHI3, has_image)
