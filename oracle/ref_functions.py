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
    keep = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    missing = set(names) - {n.name for n in keep}
    if missing:
        raise KeyError(f"{missing} not found in {rel_path}")
    ns = {"torch": torch, "Optional": Optional, "List": List, "Dict": Dict, "Sequence": Sequence,
          "dataclass": dataclass, "IGNORE_INDEX": -100, "IMAGE_TOKEN_INDEX": -200,
          "Sampler": torch.utils.data.Sampler}
    transformers = type("T", (), {"PreTrainedTokenizer": object})
    ns["transformers"] = transformers
    ns.update(extra_globals or {})
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    return {n: ns[n] for n in names}
