#!/usr/bin/env python
"""Headline benchmark: VisPer-LM data-parallel train-step samples/sec, Llama-3-8B + CLIP-ViT-L/14-336,
336 px images, embedded sequence 2048, bf16 (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # N=1; under torchrun for N>1
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one optimizer step on one synthetic batch per GPU: frozen CLIP tower forward →
mm_projector → splice → 32 decoder layers → lm_head+CE → six distillation heads + SL1/InfoNCE losses
(+ the frozen DPT decoder behind `depth_preds`) → backward → ZeRO-2 AdamW.

Default workload = the configuration `north_star`'s target sentence names ("with all dsg distill heads
active"): BASELINE.json configs[2]'s per-GPU slice — it fits one GPU, so it is also the N=1 line.  The
same run adds, as extra keys, `ntp` (configs[1]: NTP only, PT freeze policy) and `ift` (full fine-tune of the LLM as finetune.sh does: the 16 GB gradient reduce-scatter /
parameter all-gather of ZeRO-2).  `--workload ntp` makes configs[1] the main line instead.

Prints ONE JSON line (rank 0).  `value` = device-resident inputs; `e2e` = same step through the public
trainer API with HOST (pinned) inputs and a device→host read of the loss every step.  Under torchrun the
line also carries `dp_check`: a tiny-config step run data-parallel over NCCL (ZeRO-2 + cross-rank InfoNCE
negatives) against the same global batch run single-process on rank 0.

Reference arm (`--impl reference`): the UNMODIFIED reference classes (oracle/_ref, a byte-identical copy
of the reference package — oracle/build_ref.py) under oracle/ref_shim.py: CPU fp32 at BASELINE configs[0]
on the host cores (whole steps, no extrapolation) as the line's value, plus `reference_gpu`: the same
classes in bf16 with flash_attention_2 + gradient checkpointing on this arm's config and GPU — the number
north_star says to beat.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

FLOPS = {  # algorithmic FLOP per sample, Llama-3-8B, T=2048 (no recompute counted)
    # lm_head runs on the label rows only (69.7 % of the rows at this batch: positions past S+8)
    "ntp_adapter": 2 * 2.859e13 + 3.5 * 1.100e12 + 2 * 2.152e12 + 3.65e11 + 3 * 2.42e10,
    "dsg_adapter": 6.70e13,
    "ntp_full": 3 * 2.859e13 + 3.5 * 1.100e12 + 3 * 2.152e12 + 3.65e11 + 3 * 2.42e10,  # IFT, SURVEY §8d
}
ATTN_FLOP_FWD = lambda B, H, T, hd: 4.0 * B * H * T * T * hd * 0.5   # causal: half the square


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="dsg", choices=["ntp", "dsg"])
    ap.add_argument("--batch", type=int, default=8, help="per-GPU batch (samples per step per GPU)")
    ap.add_argument("--train", default="adapter", choices=["adapter", "full"],
                    help="adapter: PT-stage freeze policy (default); full: IFT-style full fine-tune of LLM + "
                         "projector (finetune.sh) — weight gradients for every layer")
    ap.add_argument("--seq", type=int, default=2048, help="embedded sequence length T")
    ap.add_argument("--model", default="llama3-8b", choices=["llama3-8b", "phi3-mini", "tiny"])
    ap.add_argument("--tower", default="clip-vit-l", choices=["clip-vit-l", "convnext-xxl"],
                    help="vision tower: CLIP-ViT-L/14-336 (default, the BASELINE metric's config) or the frozen "
                         "CLIP-ConvNeXt-XXL at 768 px of BASELINE configs[3] (576 image tokens of width 3072)")
    ap.add_argument("--layers", type=int, default=None, help="override decoder depth (debug only; reported)")
    ap.add_argument("--extras", default="auto", choices=["auto", "none", "ntp", "all"],
                    help="extra workloads measured after the main one and reported as keys of the same line "
                         "(auto = all = ntp + ift)")
    ap.add_argument("--ift-batch", type=int, default=4, help="per-GPU batch of the `ift` extra (full fine-tune)")
    ap.add_argument("--torch-profile", action="store_true",
                    help="diagnostic: torch.profiler over 2 device-leg steps, prints kernel totals and busy time")
    ap.add_argument("--teachers", action="store_true",
                    help="dsg only: compute the depth / seg / gen targets with the on-GPU frozen teachers "
                         "(DINOv2-L, Swin-L, unCLIP ViT-H) from synthetic images every step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-input leg")
    ap.add_argument("--no-dp-check", action="store_true")
    ap.add_argument("--no-reference-gpu", action="store_true", help="reference arm: skip the GPU leg")
    ap.add_argument("--profile", action="store_true",
                    help="bracket the timed device leg with cudaProfilerStart/Stop (ncu --profile-from-start off)")
    ap.add_argument("--cpu-config", default="configs0", choices=["configs0", "tiny"],
                    help="workload of the CPU reference legs: BASELINE configs[0] (default) or the tiny debug model "
                         "(contract tests)")
    ap.add_argument("--cpu-budget-s", type=float, default=240.0,
                    help="wall budget of the CPU reference legs (whole steps only; at least 3 are timed)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ helpers
CONVNEXT_XXL_FLOP = 3.563e12   # per 768 px image: stem + 3 downsamples + 40 blocks (fc1/fc2 GEMMs + 98 FLOP/elt depthwise)

# dp_check / debug model: the parity tests' tiny Llama (4 layers, hidden 128) — small enough that rank 0 can
# re-run the GLOBAL batch single-process in milliseconds
TINY = dict(family="llama", vocab=512, hidden=128, inter=256, layers=4, heads=4, kv_heads=2, max_pos=1024,
            rope_theta=500000.0, vis_hidden=64, vis_inter=128, vis_layers=3, vis_heads=2, image_size=336,
            patch_size=14, gen_dim=64, seg_dim=96, depth_dim=64, depth_layers="3-4", seg_layers="1-3",
            gen_layers="2-4", aux_mode="gen-depth-seg", num_task_tokens=8, tokenizer_model_max_length=1024)


def model_cfg(name, layers=None, tower="clip-vit-l"):
    c = _model_cfg(name, layers)
    if tower == "convnext-xxl":  # clip_convnext_encoder.py:61-174, timm convnext_xxlarge (norm_eps 1e-5)
        c.update(tower="convnext", cnx_depths=(3, 4, 30, 3), cnx_dims=(384, 768, 1536, 3072), cnx_eps=1e-5,
                 image_size=768)
    return c


def _model_cfg(name, layers=None):
    from visper_lm_b200.model import presets

    if name == "tiny":
        c = dict(TINY)
    else:
        c = dict(presets.LLAMA3_8B if name == "llama3-8b" else presets.PHI3_MINI)
    if layers is not None:
        c["layers"] = layers
        if c["layers"] < 20:
            c["depth_layers"], c["seg_layers"], c["gen_layers"] = "1-2", "1-2", "1-2"
    return c


def n_sys(c):
    if c["family"] == "phi3":
        return 13
    return 26 if c["vocab"] < 128000 else 38


def host_batch(c, B, T, distill, seed, teachers=False, n_text=None):
    """SURVEY.md §8(d) synthetic batch in the collator's schema (HOST tensors)."""
    g = torch.Generator().manual_seed(seed)
    S, V = n_sys(c), c["vocab"]
    if n_text is None:
        n_text = T - 575 - (24 if distill else 0)
    ids = torch.randint(0, V - 1, (B, n_text), generator=g)
    ids[:, S] = -200
    labels = ids.clone()
    labels[:, :S + 8] = -100
    batch = dict(input_ids=ids, labels=labels, attention_mask=torch.ones(B, n_text, dtype=torch.bool),
                 images=torch.randn(B, 3, c["image_size"], c["image_size"], generator=g))
    if distill:
        batch["distill_targets"] = dict(
            depth=torch.randn(B, 576, c["depth_dim"], generator=g).to(torch.bfloat16),
            seg=torch.randn(B, c["seg_dim"], 24, 24, generator=g).to(torch.bfloat16),
            gen=torch.randn(B, 1, c["gen_dim"], generator=g).to(torch.bfloat16))
        for k in ("depth_mask", "seg_mask", "gen_mask"):
            batch[k] = torch.ones(B, dtype=torch.long)
        if teachers:  # what the image pipeline hands the teachers: 336² uint8 RGB and CLIP-preprocessed 224²
            del batch["distill_targets"]
            batch["pil_images"] = dict(
                depth=torch.randint(0, 256, (B, 336, 336, 3), generator=g, dtype=torch.uint8),
                gen=torch.randn(B, 3, 224, 224, generator=g),
                seg=torch.randn(B, 3, 800, 800, generator=g).to(torch.bfloat16))  # OneFormerProcessor: 800²
    return batch


def cat_batches(bs):
    """Concatenate per-rank host batches into the global batch (rank-major, as DistributedSampler deals them)."""
    out = {}
    for k, v in bs[0].items():
        if isinstance(v, torch.Tensor):
            out[k] = torch.cat([b[k] for b in bs], 0)
        elif isinstance(v, dict):
            out[k] = {kk: torch.cat([b[k][kk] for b in bs], 0) for kk in v}
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.proc = None
        self.lines = []
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, pw, reasons = [], None, [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
                pw.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(pw) if pw else None}


# ------------------------------------------------------------------------------------------------ reference arm
CFG0_DESC = ("BASELINE configs[0]: Phi-3-mini-4k + CLIP-ViT-L/14-336, one 336 px image, 128 text tokens (T=727 with the "
             "24 task tokens), batch 1, fp32, PT freeze policy, all six dsg heads; one WHOLE optimizer step "
             "(forward, backward, clip, AdamW) of the unmodified reference classes, no extrapolation")


def reference_cpu(budget_s, steps, warmup=1, which="configs0"):
    """The unmodified reference (oracle/_ref under oracle/ref_shim) on the host cores at BASELINE configs[0]."""
    from oracle import ref_run
    from visper_lm_b200.model import presets

    if which == "tiny":
        c, n_text, desc = dict(TINY, num_sys_tokens=26), 60, "tiny debug model (4 layers, hidden 128), 60 text tokens, batch 1"
    else:
        c, n_text, desc = dict(presets.PHI3_MINI, num_sys_tokens=13), 128, CFG0_DESC
    r = ref_run.time_cpu(c, distill=True, n_text=n_text, B=1, steps=max(3, steps), warmup=warmup, budget_s=budget_s)
    sample = (f"{desc}; {len(r['step_s'])} timed steps after {r['warmup']} warm-up on {r['cores']} threads: "
              f"mean {r['step_s_mean']:.2f} s, min {r['step_s_min']:.2f}, max {r['step_s_max']:.2f} "
              f"(model construction {r['build_s']:.0f} s not timed)")
    return r, {"value": r["samples_per_s"], "unit": "samples/s", "cores": r["cores"], "kind": "reference",
               "source": "oracle/_ref (byte-identical copy of the reference package, oracle/ref_manifest.json) "
                         "under oracle/ref_shim.py", "sample": sample,
               "step_s": [round(t, 3) for t in r["step_s"]], "spread": (r["step_s_max"] - r["step_s_min"]) / r["step_s_mean"]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c = model_cfg(args.model, args.layers, args.tower)
    distill = args.workload == "dsg"
    import contextlib

    # the reference prints from its constructors ("Number of System Tokens: 38", ola_llama.py:69): keep stdout for
    # the ONE JSON line
    with contextlib.redirect_stdout(sys.stderr):
        r, cpu = reference_cpu(args.cpu_budget_s, args.steps, warmup=min(max(args.warmup, 1), 1), which=args.cpu_config)
    n_timed = len(r["step_s"])
    ms = 1000.0 * r["step_s_mean"]
    cfgd = workload_config(args, c, distill)   # identical to the B200 arm's config (the driver compares them)
    line = {
        "impl": "reference", "metric": "train-step samples/sec", "value": r["samples_per_s"], "unit": "samples/s",
        "n_gpus": args.gpus, "steps": n_timed, "steps_requested": args.steps, "warmup": r["warmup"],
        "warmup_requested": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": cfgd,
        "note": ("CPU leg = whole reference steps at BASELINE configs[0] (a Llama-3-8B fp32 step does not fit a few "
                 "minutes of host time); steps/warmup are what was actually timed inside the "
                 f"{args.cpu_budget_s:.0f} s budget. reference_gpu = the same classes on this arm's own config."),
        "cpu_baseline": cpu,
        "e2e": {"value": r["samples_per_s"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if torch.cuda.is_available() and not args.no_reference_gpu:
        from oracle import ref_run

        try:
            n_text = args.seq - 575 - (24 if distill else 0)
            cc = dict(c, num_sys_tokens=n_sys(c))
            with contextlib.redirect_stdout(sys.stderr):
                g = ref_run.time_gpu(cc, distill=distill, B=args.batch, n_text=n_text, steps=min(max(args.steps, 3), 8),
                                     warmup=3, train=args.train)
            line["reference_gpu"] = {
                "value": g["samples_per_s"], "unit": "samples/s", "n_gpus": 1, "ms_per_step": g["ms_per_step"],
                "ms_min": g["ms_min"], "ms_max": g["ms_max"], "steps": g["steps"], "warmup": g["warmup"],
                "dtype": "bf16", "attn_implementation": g["attn_implementation"],
                "gradient_checkpointing": g["gradient_checkpointing"], "per_gpu_batch": g["B"], "loss": g["loss"],
                "peak_mem_gb": g["peak_mem_gb"], "n_trainable": g["n_trainable"],
                "what": ("unmodified reference classes (oracle/_ref), transformers " + __import__("transformers").__version__
                         + " decoder/CLIP + flash_attn + cuBLAS, gradient_checkpointing as scripts/train/pretrain.sh:52, "
                           "fused torch AdamW on the bf16 trainables standing in for DeepSpeed ZeRO-2 (not installable), "
                           "teachers replaced by the same precomputed synthetic targets this repo's arm is fed, device-resident inputs; "
                           "as published the reference zeroes its distill masks (SURVEY §0.4), so its head backward carries zero gradients"),
                "config": cfgd["workload"]}
        except Exception as ex:
            line["reference_gpu"] = {"value": None, "error": f"{type(ex).__name__}: {str(ex)[:300]}"}
    print(json.dumps(line), flush=True)


def workload_config(args, c, distill, train=None, batch=None):
    train = train or getattr(args, "train", "adapter")
    batch = batch or args.batch
    label = "BASELINE configs[1]: " if not distill else "BASELINE configs[2] per-GPU slice: "
    if c.get("tower") == "convnext":
        label = ("BASELINE configs[3] per-GPU slice under ZeRO-2: " if distill
                 else "BASELINE configs[1] with configs[3]'s tower: ")
    return {"workload": label
            + f"{args.model} + " + ("CLIP-ConvNeXt-XXL, 768px" if c.get("tower") == "convnext" else "CLIP-ViT-L/14-336, 336px")
            + f", T={args.seq}, "
            + ("NTP only" if not distill else "NTP + dsg distill heads (d18-20_s10-18_g12-20) + frozen DPT decoder (depth_preds)")
            + (", depth (DINOv2-L) / seg (Swin-L @800) / gen (unCLIP ViT-H) targets from the on-GPU frozen teachers each step"
               if distill and getattr(args, "teachers", False) else "")
            + (", full fine-tune (LLM + mm_projector trainable: fwd + dgrad + wgrad)" if train == "full"
               else ", PT freeze policy (mm_projector" + ("+heads+task tokens" if distill else "") + " trainable; LLM fwd+dgrad)"),
            "per_gpu_batch": batch, "global_batch": batch * args.gpus, "seq_len": args.seq,
            "decoder_layers": c["layers"], "parallelism": f"dp{args.gpus} zero2",
            "l2_policy": "working set (16 GB weights + activations) exceeds the 126 MB L2; no explicit flush",
            "recompute": "none (activations kept; reference uses gradient checkpointing)"}


# ------------------------------------------------------------------------------------------------ B200 arm
class Env:
    def __init__(self):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1 and not dist.is_initialized():
            dist.init_process_group("nccl", device_id=self.dev)

    def sync_all(self):
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()


def build_trainer(env, c, distill, train, batch, teachers=False, distributed=True, lr=1e-3, warmup_ratio=0.03):
    from visper_lm_b200 import model as pm
    from visper_lm_b200.model import presets
    from visper_lm_b200.train.trainer import LLaVATrainer, TrainingArguments

    cfg = presets.from_dict(c, distill=distill)
    cls = {("llama", True): pm.OlaLlavaLlamaForCausalLM, ("phi3", True): pm.OlaLlavaPhi3ForCausalLM,
           ("llama", False): pm.LlavaLlamaForCausalLM, ("phi3", False): pm.LlavaPhi3ForCausalLM}[(c["family"], distill)]
    torch.manual_seed(0)
    model = cls(cfg, device=env.dev)
    model.init_weights(std=0.02, seed=0)
    if teachers:
        cfg.random_init_teachers = True  # no network: DINOv2-L / Swin-L / unCLIP ViT-H geometry, random weights
        model.init_target_models(cfg)
    for n, p in model.named_parameters():
        if train == "full":  # finetune.sh: everything but the frozen tower (and the frozen DPT head / teacher)
            p.requires_grad_(("vision_tower" not in n) and ("da_v2_head" not in n) and ("dav2_backbone" not in n)
                             and ("oneformer" not in n))
        else:  # PT freeze policy
            p.requires_grad_(("mm_projector" in n) or ("_heads." in n) or ("special_" in n) or n.endswith("logit_scale"))
    targs = TrainingArguments(per_device_train_batch_size=batch, learning_rate=lr, max_steps=10_000,
                              warmup_ratio=warmup_ratio)
    trainer = LLaVATrainer(model=model, args=targs, distributed=distributed)
    trainer.total_steps = 10_000
    trainer.create_optimizer()
    return model, trainer


def fresh(b):
    out = dict(b)
    for k in ("depth_mask", "seg_mask", "gen_mask"):
        if k in out:
            out[k] = out[k].clone()
    return out


def measure(env, args, c, *, distill, train, batch, steps, warmup, teachers=False, want_e2e=True, main=False):
    """Build the model, run warm-up, time `steps` device-resident steps (and the same through host inputs)."""
    from visper_lm_b200 import lib, ops

    dev, world = env.dev, env.world
    model, trainer = build_trainer(env, c, distill, train, batch, teachers)
    B, T = batch, args.seq
    hb = host_batch(c, B, T, distill, 1234 + env.rank, teachers)
    pinned = {k: (v.pin_memory() if isinstance(v, torch.Tensor) else
                  {kk: vv.pin_memory() for kk, vv in v.items()}) for k, v in hb.items()}
    # device-resident copies for the kernel-side leg (ids stay on the host: the splice is planned there)
    devb = {k: (v if k in ("input_ids", "labels", "attention_mask") else
                (v.to(dev) if isinstance(v, torch.Tensor) else {kk: vv.to(dev) for kk, vv in v.items()}))
            for k, v in hb.items()}

    def timed(b, n, read_loss):
        env.sync_all()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        last = None
        for _ in range(n):
            loss, _ = trainer.step(fresh(b))
            last = float(loss) if read_loss else loss  # float(): device→host read of the step's result
        e.record()
        env.sync_all()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), float(last)

    for _ in range(max(warmup, 3)):
        trainer.step(fresh(devb))
    env.sync_all()
    if main and args.torch_profile:
        from torch.profiler import ProfilerActivity, profile

        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            t0 = time.perf_counter()
            for _ in range(2):
                trainer.step(fresh(devb))
            t_cpu = time.perf_counter() - t0
            torch.cuda.synchronize()
            t_all = time.perf_counter() - t0
        evs = [e for e in prof.key_averages() if e.device_time_total > 0]
        busy = sum(e.device_time_total for e in evs) / 1e3
        if env.rank == 0:
            print(json.dumps({"profile_steps": 2, "cpu_enqueue_ms": t_cpu * 1e3, "wall_ms": t_all * 1e3,
                              "gpu_busy_ms": busy}), flush=True)
            for e in sorted(evs, key=lambda e: -e.device_time_total)[:30]:
                print(json.dumps({"kernel": e.key[:70], "calls": e.count, "total_ms": round(e.device_time_total / 1e3, 2)}),
                      flush=True)
        return None
    lib.reset_launch_count()
    ops.KERNEL_TIMER = ops.KernelTimer()
    if main and args.profile:
        torch.cuda.profiler.start()
    ms_dev, loss_dev = timed(devb, steps, read_loss=False)
    if main and args.profile:
        torch.cuda.profiler.stop()
    kstats = ops.KERNEL_TIMER.summary()
    ops.KERNEL_TIMER = None
    launches = lib.launch_count()
    if want_e2e:
        ms_e2e, loss_e2e = timed(pinned, steps, read_loss=True)
    else:
        ms_e2e, loss_e2e = ms_dev, loss_dev
    h2d = sum(v.numel() * v.element_size() for k, v in hb.items() if isinstance(v, torch.Tensor)
              and k not in ("input_ids", "attention_mask"))
    h2d += sum(vv.numel() * vv.element_size() for v in hb.values() if isinstance(v, dict) for vv in v.values())
    h2d += 3 * B * T * 4 + B * 576 * 4  # splice plan (kind/index/scatter int32) + inverse image map
    opt = trainer.optimizer
    res = dict(ms_dev=ms_dev, ms_e2e=ms_e2e, loss_dev=loss_dev, loss_e2e=loss_e2e, launches=launches, kstats=kstats,
               h2d=int(h2d), samples=B * world * steps, steps=steps, batch=B,
               trainable_params=int(sum(p.numel() for _, p in opt.named)),
               comm=opt.comm_summary() if hasattr(opt, "comm_summary") else None,
               mem_gb=torch.cuda.max_memory_allocated() / 2**30)
    del model, trainer, opt, devb, pinned
    gc.collect()
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
    return res


def brief(res, world):
    return {"value": res["samples"] / (res["ms_dev"] / 1e3), "unit": "samples/s",
            "ms_per_step": res["ms_dev"] / res["steps"],
            "e2e": {"value": res["samples"] / (res["ms_e2e"] / 1e3), "unit": "samples/s",
                    "h2d_bytes_per_step": res["h2d"], "d2h_bytes_per_step": 4,
                    "ms_per_step": res["ms_e2e"] / res["steps"]},
            "steps": res["steps"], "per_gpu_batch": res["batch"], "gpu_launches": int(res["launches"]),
            "loss": res["loss_dev"], "trainable_params": res["trainable_params"], "peak_mem_gb": round(res["mem_gb"], 1),
            "comm": res["comm"]}


def dp_check(env):
    """NCCL ZeRO-2 + cross-rank InfoNCE numerics on hardware: a tiny dsg model takes two optimizer steps
    data-parallel on per-rank batches; rank 0 re-runs the concatenated GLOBAL batches single-process (no
    collectives) from the same initial weights.  DP semantics to match: loss = mean of rank-local means, gradients
    averaged over ranks, InfoNCE negatives all-gathered with labels offset by rank·B (ola_utils.py:96-125), AdamW
    on the rank's shard then parameter all-gather (llava_trainer.py:890-995, scripts/zero2.json)."""
    c = dict(TINY)
    B, n_text, steps = 2, 40, 2
    T = n_text - 1 + 576 + 24

    def snapshot(model):
        return {n: p.detach().float().clone() for n, p in model.named_parameters() if p.requires_grad}

    model, trainer = build_trainer(env, c, True, "adapter", B, warmup_ratio=0.0)
    p0 = snapshot(model)
    losses, norms = [], []
    for s_ in range(steps):
        loss, _ = trainer.step(fresh(host_batch(c, B, T, True, 4321 + 10 * s_ + env.rank, n_text=n_text)))
        lt = loss.detach().float().reshape(1).clone()
        dist.all_reduce(lt)
        losses.append(lt.item() / env.world)
        norms.append(float(trainer.optimizer.last_grad_norm))
    trainer.optimizer.wait_params()
    env.sync_all()
    p_dp = snapshot(model)
    out = None
    if env.rank == 0:
        model1, trainer1 = build_trainer(env, c, True, "adapter", B * env.world, distributed=False, warmup_ratio=0.0)
        model1._gather_targets = lambda t: (t, 0, None)     # single process: the global batch holds every target
        q0 = snapshot(model1)
        losses1, norms1 = [], []
        for s_ in range(steps):
            gb = cat_batches([host_batch(c, B, T, True, 4321 + 10 * s_ + r, n_text=n_text) for r in range(env.world)])
            loss1, _ = trainer1.step(fresh(gb))
            losses1.append(float(loss1))
            norms1.append(float(trainer1.optimizer.last_grad_norm))
        torch.cuda.synchronize()
        p1 = snapshot(model1)
        up_dp = torch.cat([(p_dp[n] - p0[n]).flatten() for n in p0])
        up_1 = torch.cat([(p1[n] - q0[n]).flatten() for n in p0])
        out = {"world": env.world, "steps": steps,
               "config": "tiny Llama (4 layers, hidden 128) + dsg heads, B=2/rank, T=639, PT freeze policy, lr 1e-3, no warm-up",
               "loss_dp": losses, "loss_single": losses1,
               "loss_rel": max(abs(a - b) / abs(b) for a, b in zip(losses, losses1)),
               "grad_norm_dp": norms, "grad_norm_single": norms1,
               "grad_norm_rel": max(abs(a - b) / max(b, 1e-30) for a, b in zip(norms, norms1)),
               "init_params_equal": all(torch.equal(p0[n], q0[n]) for n in p0),
               "param_after_step_max_abs": max(float((p_dp[n] - p1[n]).abs().max()) for n in p0),
               "param_update_cos": float((up_dp @ up_1) / (up_dp.norm() * up_1.norm()).clamp_min(1e-30)),
               "param_update_norm_rel": float((up_dp.norm() - up_1.norm()).abs() / up_1.norm().clamp_min(1e-30)),
               "trainable_tensors": len(p0),
               "note": "AdamW's first steps move a weight by about lr*sign(g): max_abs is at most 2*lr per step where a "
                       "near-zero gradient changes sign under bf16 rounding; update_cos / update_norm_rel are the aggregates"}
        del model1, trainer1
    del model, trainer
    gc.collect()
    torch.cuda.empty_cache()
    env.sync_all()
    return out


def run_b200(args):
    env = Env()
    world, rank = env.world, env.rank
    distill = args.workload == "dsg"
    c = model_cfg(args.model, args.layers, args.tower)
    teachers = bool(args.teachers and distill)

    check = None
    if world > 1 and not args.no_dp_check:
        try:
            check = dp_check(env)
        except Exception as ex:  # reported, never fatal for the throughput line
            check = {"error": f"{type(ex).__name__}: {str(ex)[:300]}"}
            env.sync_all()

    clocks = ClockSampler(env.local)  # every rank samples its own GPU: synchronous DP runs at the slowest one's clock
    clocks.start()
    res = measure(env, args, c, distill=distill, train=args.train, batch=args.batch, steps=args.steps,
                  warmup=args.warmup, teachers=teachers, want_e2e=not args.no_e2e, main=True)
    clk = clocks.stop()
    if world > 1:
        every = [None] * world
        dist.all_gather_object(every, clk)
        clk = dict(every[0], sm_mhz_per_rank=[e.get("sm_mhz") for e in every],
                   reasons=sorted({r for e in every for r in e.get("reasons", [])}))
    if res is None:  # --torch-profile
        if world > 1:
            dist.destroy_process_group()
        return

    extras = {}
    want = {"auto": ["ntp", "ift"], "none": [], "ntp": ["ntp"], "all": ["ntp", "ift"]}[args.extras]
    full_model = args.model == "llama3-8b" and args.layers is None and args.seq == 2048 and args.tower == "clip-vit-l"
    if not (args.train == "adapter" and not teachers):
        want = []
    ksteps = max(3, min(args.steps, 8))
    for name in want:
        if name == "ntp" and not distill:
            continue
        try:
            if name == "ntp":
                r = measure(env, args, c, distill=False, train="adapter", batch=args.batch, steps=ksteps, warmup=3)
                d = brief(r, world)
                d["config"] = workload_config(args, c, False, "adapter")["workload"]
            else:
                r = measure(env, args, c, distill=False, train="full", batch=args.ift_batch, steps=ksteps, warmup=3)
                d = brief(r, world)
                d["config"] = workload_config(args, c, False, "full", args.ift_batch)["workload"]
                if full_model:
                    d["step_model_flops_frac"] = d["value"] / world * FLOPS["ntp_full"] / (peak_tflops()[0] * 1e12)
            extras[name] = d
        except Exception as ex:
            extras[name] = {"error": f"{type(ex).__name__}: {str(ex)[:300]}"}
            gc.collect()
            torch.cuda.empty_cache()
            env.sync_all()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak_tf, peak_src = peak_tflops()
    peaks = read_peaks()
    hbm = peaks.get("hbm_gbs", 6650.0)
    flop_key = "dsg_adapter" if distill else ("ntp_full" if args.train == "full" else "ntp_adapter")
    step_flop = FLOPS[flop_key]
    if c.get("tower") == "convnext":  # tower forward and the 3072-wide projector input replace the ViT-L terms
        step_flop += (CONVNEXT_XXL_FLOP - 3.65e11) + 3 * (2 * 576 * (3072 - 1024) * 4096)
    full_any_tower = args.model == "llama3-8b" and args.layers is None and args.seq == 2048
    b = brief(res, world)
    ks = res["kstats"]
    gemm = ks.get("gemm", {"tflops": None, "ms": 0.0, "launches": 0})
    line = {
        "metric": "train-step samples/sec", "value": b["value"], "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": b["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic", "config": workload_config(args, c, distill),
        "e2e": b["e2e"], "gpu_launches": b["gpu_launches"], "clocks": clk,
        "loss": {"device_leg": res["loss_dev"], "e2e_leg": res["loss_e2e"]},
        "trainable_params": b["trainable_params"], "peak_mem_gb": b["peak_mem_gb"], "comm": b["comm"],
        "roofline": {
            "bound": "tensor",
            "kernel": "gemm_tcgen05_pair_kernel / gemm_tcgen05_kernel (all GEMM launches of the timed steps)",
            "achieved": gemm["tflops"], "peak": peak_tf, "unit": "TFLOP/s",
            "frac": gemm["tflops"] / peak_tf if gemm["tflops"] else None,
            # ncu --set full inside the dsg step (profiles/r02_ncu_gemm_instep.csv), the launch that moves the most DRAM
            # bytes: gate|up dgrad, M=16384 N=4096 K=28672 — 7.41 GB read + 0.13 GB written per launch against 1.31 GB of
            # operands + output (A panels of 32 MB stay L2-resident, B streams once per panel; tensor pipe 99.4 % active,
            # 2.75 ms).  The other three decoder dgrad shapes of the same capture: 1.35 / 0.45 / 0.91 GB per launch.
            "traffic": 7.55e9 if full_model else None,
            "traffic_algorithmic": 1.31e9 if full_model else None,
            "traffic_source": "profiles/r02_ncu_gemm_instep.csv (one --set full capture, per launch)" if full_model else None,
            "peak_source": peak_src, "launches": gemm["launches"],
            "gemm_ms_per_step": gemm["ms"] / args.steps,
            "gemm_share_of_step": gemm["ms"] / res["ms_dev"] if res["ms_dev"] else None,
            "step_model_flops_frac": (b["value"] / world * step_flop / (peak_tf * 1e12)) if full_any_tower else None,
            # the other named kernels, timed live with CUDA events on the launching stream inside the same steps
            "kernels": kernel_rooflines(ks, args.steps, peak_tf, hbm),
        },
    }
    line.update(extras)
    if check is not None:
        line["dp_check"] = check
    if world == 1 and not args.no_cpu_baseline:
        try:
            import contextlib

            with contextlib.redirect_stdout(sys.stderr):
                _, line["cpu_baseline"] = reference_cpu(args.cpu_budget_s, 3, warmup=1, which=args.cpu_config)
        except Exception as ex:  # the baseline is reported, never required for the GPU number — but never silently lost
            import traceback

            traceback.print_exc(file=sys.stderr)
            line["cpu_baseline"] = {"value": None, "unit": "samples/s", "cores": os.cpu_count(), "kind": "reference",
                                    "sample": f"failed: {type(ex).__name__}: {ex}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def read_peaks():
    pf = ROOT / "MEASURED_PEAKS.json"
    return json.loads(pf.read_text()) if pf.exists() else {}


def peak_tflops():
    peaks = read_peaks()
    if peaks:
        return peaks.get("bf16_tflops_sustained", 1400.0), "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)"
    return 1400.0, "fallback 1.4 PFLOP/s sustained (of fallback)"


def kernel_rooflines(ks, steps, peak_tf, hbm_gbs):
    out = {}
    for name, st in ks.items():
        if name == "gemm" or not st["launches"]:
            continue
        d = {"launches": st["launches"], "ms_per_step": st["ms"] / steps}
        if st.get("flops"):
            d.update(bound="tensor", achieved=st["tflops"], unit="TFLOP/s", frac=st["tflops"] / peak_tf)
        elif st.get("bytes"):
            gbs = st["bytes"] / st["ms"] / 1e6
            d.update(bound="hbm", achieved=gbs, unit="GB/s", frac=gbs / hbm_gbs)
        if name in KERNEL_NOTES:
            d["note"] = KERNEL_NOTES[name]
        out[name] = d
    return out


# context for the small-problem entries: their fraction of a tensor peak says how small the launches are, not how the kernel runs
KERNEL_NOTES = {
    "attn_fwd_hd64": "CLIP ViT-L tower: 23 launches per step of B*16 heads x 5 query tiles at S=577 (tcgen05, one work item per CTA); "
                     "latency-bound at 4-5 key tiles per CTA",
    "attn_fwd_hd32": "Perceiver cross-attention of the six task-token heads (mma.sync): 576 keys + a few latents per sample, microseconds per launch",
    "attn_bwd_hd32": "backward of the same cross-attention (mma.sync), microseconds per launch",
}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
