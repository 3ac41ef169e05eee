"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.pt by running the UNMODIFIED reference
classes (through oracle/ref_shim.py) on seeded tiny configs.  Run here, where /root/reference
exists:  python -m oracle.make_golden

Each fixture stores the recipe (config name, batch seed/shape — weights come from
restate.seeded_param by parameter name) and the reference's outputs: every loss term, sub-sampled
logits / hidden states / head embeddings, and gradients of the PT-stage trainable parameters for
the live-mask objective  text_loss + Σ_task Σ_layer 0.5·_emb_loss(pred, ones, target, tau).
"""
from __future__ import annotations

import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import configs, ref_shim, restate  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"

CASES = [
    # name, cfg, family, distill, B, n_text, pad_rows
    ("tiny_llama_dsg", "TINY_LLAMA", "llama", True, 2, 40, 0),
    ("tiny_llama_dsg_padded", "TINY_LLAMA", "llama", True, 3, 48, 1),
    ("tiny_phi3_dsg", "TINY_PHI3", "phi3", True, 2, 40, 0),
    ("tiny_phi3_sw_dsg", "TINY_PHI3_SW", "phi3", True, 2, 40, 0),
    ("tiny_llama_ntp", "TINY_LLAMA", "llama", False, 2, 40, 0),
    # pad_rows = -1 selects configs.synthetic_batch_mixed: text-only row + two-image row + one-image row
    ("tiny_llama_ntp_mixed", "TINY_LLAMA", "llama", False, 3, 48, -1),
    ("wide_llama_dsg", "WIDE_LLAMA", "llama", True, 2, 40, 0),
    ("wide_phi3_dsg", "WIDE_PHI3", "phi3", True, 2, 40, 0),
]


def trainable_pt(name: str) -> bool:
    """PT-stage freeze policy (ola_vlm_train.py:1127-1131, 1239-1266): projector, heads, task tokens,
    logit scales train; LLM, tower and the frozen DPT head do not."""
    return (("mm_projector" in name) or ("_heads." in name) or ("special_" in name)
            or name.endswith("logit_scale"))


def sub(t, s0=1, s1=1):
    return t[..., ::s0, ::s1].contiguous().clone()


def run_case(name, cfg_name, family, distill, B, n_text, pad_rows):
    cfg = getattr(configs, cfg_name)
    model = ref_shim.build_reference_model(cfg, family, distill, seed_fn=restate.seeded_param)
    for n, p in model.named_parameters():
        p.requires_grad_(trainable_pt(n))
    batch = (configs.synthetic_batch_mixed(cfg, n_text, seed=1234) if pad_rows < 0 else
             configs.synthetic_batch(cfg, B, n_text, seed=1234, distill=distill, pad_rows=pad_rows))
    kwargs = dict(input_ids=batch["input_ids"], labels=batch["labels"],
                  attention_mask=batch["attention_mask"], images=batch["images"])
    fx = {"name": name, "cfg_name": cfg_name, "family": family, "distill": distill, "B": B,
          "n_text": n_text, "pad_rows": pad_rows, "seed": 1234}
    if distill:
        ref_shim.install_synthetic_teachers(model, batch["targets"])
        masks = {k: v.clone() for k, v in batch["masks"].items()}
        kwargs.update(pil_images=[None] * B, depth_mask=masks["depth"], seg_mask=masks["seg"],
                      gen_mask=masks["gen"])
    if not distill:
        kwargs["output_hidden_states"] = True  # LlavaLlama defers to HF forward (llava_llama.py:108-119)
    out = model(**kwargs)
    fx["loss_as_published"] = float(out.loss.detach())
    fx["logits_sub"] = sub(out.logits.detach(), 16, 8)
    fx["hidden_sub"] = [sub(h.detach(), 32, 8) for h in out.hidden_states]
    text_loss = torch.nn.functional.cross_entropy(
        out.logits[:, :-1].reshape(-1, out.logits.shape[-1]),
        _spliced_labels(model, batch)[:, 1:].reshape(-1), ignore_index=-100)
    fx["text_loss"] = float(text_loss)
    total = text_loss
    if distill:
        assert all(int(m.sum()) == 0 for m in masks.values()), "reference zeroes the masks in place"
        fx["masks_zeroed_in_place"] = True
        ones = torch.ones(B, dtype=torch.long)
        per = {}
        for task, embs, scale in (("depth", [e[0][0] for e in out.depth_embs], model.depth_logit_scale),
                                  ("seg", list(out.seg_embs), model.seg_logit_scale),
                                  ("gen", list(out.image_embs), model.gen_logit_scale)):
            per[task] = []
            for e in embs:
                l, s1, c = model._emb_loss(e, ones.clone(), batch["targets"][task], scale)
                per[task].append((float(l), float(s1), float(c)))
                total = total + 0.5 * l
            fx[f"{task}_emb_sub"] = [sub(e.detach().flatten(1), 1, 97) for e in embs]
        fx["emb_losses_live"] = per
    fx["loss_live"] = float(total)
    total.backward()
    grads = {}
    for n, p in model.named_parameters():
        if p.requires_grad and p.grad is not None:
            g = p.grad.detach().flatten()
            grads[n] = {"norm": float(g.norm()), "head": g[:16].clone(),
                        "stride": g[::max(1, g.numel() // 64)][:64].clone()}
    fx["grads"] = grads
    fx["state_spec"] = {n: tuple(p.shape) for n, p in model.named_parameters() if "da_v2_head" not in n}
    GOLDEN.mkdir(parents=True, exist_ok=True)
    torch.save(fx, GOLDEN / f"{name}.pt")
    print(name, "loss_as_published", fx["loss_as_published"], "loss_live", fx["loss_live"],
          "n_grads", len(grads), "size", (GOLDEN / f"{name}.pt").stat().st_size)


def _spliced_labels(model, batch):
    with torch.no_grad():
        return model.prepare_inputs_labels_for_multimodal(
            batch["input_ids"], None, batch["attention_mask"], None, batch["labels"], batch["images"])[5]


if __name__ == "__main__":
    only = sys.argv[1:] or None
    for case in CASES:
        if only is None or case[0] in only:
            run_case(*case)
