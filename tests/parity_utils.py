"""Shared helpers for the parity tests: same seeded weights (by reference parameter name, rounded
to bf16 once) into the CUDA product model and the CPU fp32 oracle; same synthetic batch."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from oracle import configs, restate  # noqa: E402


def bf16_seeded(name, shape):
    return restate.seeded_param(name, shape).to(torch.bfloat16)


def product_classes():
    from visper_lm_b200 import model as pm

    return {("llama", True): pm.OlaLlavaLlamaForCausalLM, ("phi3", True): pm.OlaLlavaPhi3ForCausalLM,
            ("llama", False): pm.LlavaLlamaForCausalLM, ("phi3", False): pm.LlavaPhi3ForCausalLM}


def build_product(cfg_dict, distill, device):
    from visper_lm_b200.model import presets

    cls = product_classes()[(cfg_dict["family"], distill)]
    cfg = presets.from_dict(cfg_dict, distill=distill)
    model = cls(cfg, device=device)
    model.init_weights(seed_fn=bf16_seeded)
    return model


def oracle_state(model):
    """fp32 copies of the product's (bf16-rounded) weights, keyed by reference names."""
    return {n: p.detach().float().cpu() for n, p in model.named_parameters()}


def pt_freeze(model):
    """PT-stage policy (ola_vlm_train.py:1127-1131,1239-1266)."""
    for n, p in model.named_parameters():
        p.requires_grad_(("mm_projector" in n) or ("_heads." in n) or ("special_" in n)
                         or n.endswith("logit_scale"))


def round_batch(batch):
    """Inputs rounded to bf16 once so oracle and CUDA see identical values."""
    b = dict(batch)
    b["images"] = batch["images"].to(torch.bfloat16).float()
    if "targets" in batch:
        b["targets"] = {k: v.to(torch.bfloat16).float() for k, v in batch["targets"].items()}
    return b


def run_product(model, batch, distill, device):
    kw = dict(input_ids=batch["input_ids"], labels=batch["labels"], attention_mask=batch["attention_mask"],
              images=batch["images"].to(device))
    if distill:
        kw.update(distill_targets={k: v.to(device) for k, v in batch["targets"].items()},
                  depth_mask=batch["masks"]["depth"].clone().to(device),
                  seg_mask=batch["masks"]["seg"].clone().to(device),
                  gen_mask=batch["masks"]["gen"].clone().to(device))
    return model(**kw)


def rel_err(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def cos_sim(a, b):
    a, b = a.float().cpu().flatten(), b.float().cpu().flatten()
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-20)).item()
