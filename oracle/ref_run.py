"""BENCH INFRASTRUCTURE ONLY — times the UNMODIFIED reference classes (oracle/ref_shim.py over
/root/reference or its byte-identical copy oracle/_ref/) on the workload bench.py measures.

Two legs, both the reference's own `model(**batch).loss → backward → AdamW` step (the stock HF
`Trainer.training_step` the reference relies on — SURVEY.md §0.6 — with torch AdamW standing in for
DeepSpeed's wrapper of the same optimizer, `scripts/zero2.json` has no optimizer key):

* CPU: fp32 on the host cores, BASELINE.json configs[0] (Phi-3-mini-4k + CLIP-ViT-L/14-336, one 336 px
  image, 128 text tokens, batch 1, PT freeze policy).  Whole steps, no extrapolation.
* GPU (only when CUDA is present): bf16, `attn_implementation="flash_attention_2"`
  (ola_vlm/train/ola_vlm_train_mem.py:5) + gradient checkpointing (scripts/train/pretrain.sh:52) on the
  bench's own config — the number `north_star` says to beat.

The frozen teachers are replaced by fixed synthetic targets (install_synthetic_teachers): the bench's
default workload feeds precomputed targets to both arms.  Only bench.py, tests/ and smoke() import this.
"""
from __future__ import annotations

import os
import time

import torch

from . import ref_shim


def n_sys(c):
    if c["family"] == "phi3":
        return 13
    return 26 if c["vocab"] < 128000 else 38


def make_batch(c, B, n_text, distill, seed, device="cpu", dtype=torch.float32):
    """Collator-schema batch (ola_vlm_train.py:887-925) + synthetic distillation targets."""
    g = torch.Generator().manual_seed(seed)
    S, V = n_sys(c), c["vocab"]
    ids = torch.randint(0, V - 1, (B, n_text), generator=g)
    ids[:, S] = -200
    labels = ids.clone()
    labels[:, :S + 8] = -100
    batch = dict(input_ids=ids.to(device), labels=labels.to(device),
                 attention_mask=torch.ones(B, n_text, dtype=torch.bool, device=device),
                 images=torch.randn(B, 3, c["image_size"], c["image_size"], generator=g).to(device=device, dtype=dtype))
    targets = None
    if distill:
        targets = dict(depth=torch.randn(B, 576, c["depth_dim"], generator=g).to(device=device, dtype=dtype),
                       seg=torch.randn(B, c["seg_dim"], 24, 24, generator=g).to(device=device, dtype=dtype),
                       gen=torch.randn(B, 1, c["gen_dim"], generator=g).to(device=device, dtype=dtype))
        batch["pil_images"] = [None] * B
    return batch, targets


def pt_freeze(model):
    """PT-stage policy (ola_vlm_train.py:1127-1131 then :1239-1266): everything frozen, then the projector,
    the heads, the task tokens and the logit scales train."""
    n_train = 0
    for n, p in model.named_parameters():
        on = ("mm_projector" in n) or ("_heads." in n) or ("special_" in n) or n.endswith("logit_scale")
        p.requires_grad_(on)
        n_train += p.numel() if on else 0
    return n_train


class ReferenceStep:
    """One optimizer step of the reference model, callable repeatedly."""

    def __init__(self, c, *, distill, B, n_text, device="cpu", dtype=torch.float32, attn="sdpa",
                 grad_ckpt=False, train="adapter", seed=1234, lr=1e-3):
        self.c, self.distill, self.B, self.device = c, distill, B, torch.device(device)
        t0 = time.perf_counter()
        self.model = ref_shim.build_reference_model(c, c["family"], distill, attn_implementation=attn,
                                                    device=device, dtype=dtype, fast_init=True, train_mode=True)
        if train == "adapter":
            self.n_trainable = pt_freeze(self.model)
        else:  # finetune.sh: everything but the tower / frozen DPT head
            self.n_trainable = 0
            for n, p in self.model.named_parameters():
                on = ("vision_tower" not in n) and ("da_v2_head" not in n)
                p.requires_grad_(on)
                self.n_trainable += p.numel() if on else 0
        if grad_ckpt:
            # ola_vlm_train.py:1081-1088 + HF Trainer(gradient_checkpointing=True)
            self.model.gradient_checkpointing_enable()
            self.model.enable_input_require_grads()
        self.batch, self.targets = make_batch(c, B, n_text, distill, seed, device, dtype)
        if distill:
            ref_shim.install_synthetic_teachers(self.model, self.targets)
        params = [p for p in self.model.parameters() if p.requires_grad]
        self.params = params
        kw = dict(fused=True) if self.device.type == "cuda" else {}
        self.opt = torch.optim.AdamW(params, lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, **kw)
        self.build_s = time.perf_counter() - t0
        self.n_params = sum(p.numel() for p in self.model.parameters())

    def __call__(self):
        b = dict(self.batch)
        if self.distill:  # the reference zeroes its masks in place (base_ola_vlm.py:472-473) → fresh ones each step
            for k in ("depth_mask", "seg_mask", "gen_mask"):
                b[k] = torch.ones(self.B, dtype=torch.long, device=self.device)
        self.opt.zero_grad(set_to_none=True)
        out = self.model(**b)
        loss = out.loss
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.params, 1.0)
        self.opt.step()
        return loss.detach()


def time_cpu(c, *, distill, n_text=128, B=1, steps=3, warmup=1, budget_s=240.0, threads=None):
    """CPU leg: whole reference steps at BASELINE configs[0] shapes.  Returns a dict."""
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = ReferenceStep(c, distill=distill, B=B, n_text=n_text, device="cpu", dtype=torch.float32, attn="sdpa")
    times, loss = [], None
    t_begin = time.perf_counter()
    for i in range(warmup + steps):
        if i > warmup and times and (time.perf_counter() - t_begin) + times[-1] > budget_s and len(times) >= 3:
            break
        t0 = time.perf_counter()
        loss = float(step())
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    T = n_text - 1 + 576 + (24 if distill else 0)
    mean = sum(times) / len(times)
    return {"samples_per_s": B / mean, "step_s": times, "step_s_mean": mean, "step_s_min": min(times),
            "step_s_max": max(times), "cores": cores, "build_s": step.build_s, "loss": loss, "T": T, "B": B,
            "n_params": step.n_params, "n_trainable": step.n_trainable, "warmup": warmup}


def time_gpu(c, *, distill, B, n_text, steps=5, warmup=3, attn="flash_attention_2", grad_ckpt=True,
             train="adapter", device="cuda:0"):
    """GPU leg: the reference classes, bf16, flash-attn + cuBLAS + gradient checkpointing."""
    torch.cuda.set_device(device)
    used = attn
    try:
        step = ReferenceStep(c, distill=distill, B=B, n_text=n_text, device=device, dtype=torch.bfloat16, attn=attn,
                             grad_ckpt=grad_ckpt, train=train)
        for _ in range(warmup):
            step()
    except Exception as ex:  # flash_attn not usable on this box → the library's next-best fused path
        if attn == "sdpa":
            raise
        used = f"sdpa (flash_attention_2 failed: {type(ex).__name__}: {str(ex)[:120]})"
        step = None
        torch.cuda.empty_cache()
        step = ReferenceStep(c, distill=distill, B=B, n_text=n_text, device=device, dtype=torch.bfloat16, attn="sdpa",
                             grad_ckpt=grad_ckpt, train=train)
        for _ in range(warmup):
            step()
    torch.cuda.synchronize()
    evs = []
    loss = None
    for _ in range(steps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        loss = step()
        e.record()
        evs.append((s, e))
    torch.cuda.synchronize()
    ms = [s.elapsed_time(e) for s, e in evs]
    mean = sum(ms) / len(ms)
    return {"samples_per_s": B / (mean / 1e3), "ms_per_step": mean, "ms_min": min(ms), "ms_max": max(ms),
            "steps": steps, "warmup": warmup, "attn_implementation": used, "gradient_checkpointing": grad_ckpt,
            "loss": float(loss), "B": B, "build_s": step.build_s, "n_trainable": step.n_trainable,
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}
