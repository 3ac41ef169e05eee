"""TEST INFRASTRUCTURE ONLY — pulls single pure functions / classes out of the reference's source
files (no package import: the modules around them need deepspeed, accelerate, …) so the host-side
index logic can be compared against the reference itself where /root/reference is mounted."""
from __future__ import annotations

import ast
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch

from . import ref_shim


def extract(rel_path: str, names: Sequence[str], extra_globals: Optional[dict] = None) -> dict:
    path = os.path.join(ref_shim.REF_ROOT, rel_path)
    tree = ast.parse(open(path).read(), filename=path)
    def _name(n):
        if isinstance(n, (ast.FunctionDef, ast.ClassDef)):
            return n.name
        if isinstance(n, ast.Assign) and len(n.targets) == 1 and isinstance(n.targets[0], ast.Name):
            return n.targets[0].id
        return None

    keep = [n for n in tree.body if _name(n) in names]
    missing = set(names) - {_name(n) for n in keep}
    if missing:
        raise KeyError(f"{missing} not found in {rel_path}")
    ns = {"torch": torch, "Optional": Optional, "List": List, "Dict": Dict, "Sequence": Sequence,
          "dataclass": dataclass, "IGNORE_INDEX": -100, "IMAGE_TOKEN_INDEX": -200,
          "Sampler": torch.utils.data.Sampler, "dataclasses": __import__("dataclasses"),
          "auto": __import__("enum").auto, "Enum": __import__("enum").Enum, "Tuple": __import__("typing").Tuple,
          "copy": __import__("copy")}
    transformers = type("T", (), {"PreTrainedTokenizer": object})
    ns["transformers"] = transformers
    ns.update(extra_globals or {})
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    return {n: ns[n] for n in names}


def conversation_lib(default_version: str):
    """A stand-in for `ola_vlm.conversation` (imported as conversation_lib by the train script) holding
    the reference's own Conversation class and templates, with `default_conversation` selected."""
    import types

    ns = extract("ola_vlm/conversation.py", ["SeparatorStyle", "Conversation", "conv_llava_llama_3", "conv_llava_phi_3"])
    lib = types.SimpleNamespace(**ns)
    lib.default_conversation = {"llama3": ns["conv_llava_llama_3"], "phi3": ns["conv_llava_phi_3"]}[default_version]
    return lib


def preprocess_fns(default_version: str):
    """The reference's preprocess_multimodal / preprocess_llama_3 / preprocess_phi_3 bound to the template."""
    import types

    from visper_lm_b200.train.data import tokenizer_image_token as _unused  # noqa: F401

    lib = conversation_lib(default_version)
    tok = extract("ola_vlm/mm_utils.py", ["tokenizer_image_token"])["tokenizer_image_token"]
    g = {"conversation_lib": lib, "tokenizer_image_token": tok, "DEFAULT_IMAGE_TOKEN": "<image>",
         "DEFAULT_IM_START_TOKEN": "<im_start>", "DEFAULT_IM_END_TOKEN": "<im_end>",
         "DataArguments": object}
    return extract("ola_vlm/train/ola_vlm_train.py", ["preprocess_multimodal", "preprocess_llama_3", "preprocess_phi_3"], g)


def extract_method(rel_path: str, cls: str, method: str, extra_globals: Optional[dict] = None):
    """One method of a reference class as a plain function (first argument = self)."""
    path = os.path.join(ref_shim.REF_ROOT, rel_path)
    tree = ast.parse(open(path).read(), filename=path)
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for fn in node.body:
                if isinstance(fn, ast.FunctionDef) and fn.name == method:
                    ns = {"torch": torch}
                    ns.update(extra_globals or {})
                    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
                    return ns[method]
    raise KeyError(f"{cls}.{method} not found in {rel_path}")
