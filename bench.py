#!/usr/bin/env python
"""Headline benchmark: VisPer-LM data-parallel train-step samples/sec, Llama-3-8B + CLIP-ViT-L/14-336,
336 px images, embedded sequence 2048, bf16 (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # N=1; under torchrun for N>1
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one optimizer step on one synthetic batch per GPU: frozen CLIP tower forward →
mm_projector → splice → 32 decoder layers → lm_head+CE (→ distillation heads with --workload dsg)
→ backward → ZeRO-2 AdamW.  Default workload = BASELINE.json configs[1] (NTP only, 1×B200) with the
PT-stage freeze policy (projector trainable, LLM forward + dgrad).

Prints ONE JSON line (rank 0).  `value` = device-resident inputs; `e2e` = same step through the
public trainer API with HOST (pinned) inputs and a device→host read of the loss every step.
The reference arm times the CPU oracle restatement of the same workload on the host cores (the
reference is pure Python on third-party libraries; its GPU build needs deepspeed/accelerate which are
not installable here — DESIGN.md §Baselines).
"""
from __future__ import annotations

import argparse
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
    "ntp_adapter": 2 * 2.859e13 + 3.5 * 1.100e12 + 2 * 2.152e12 + 3.65e11 + 3 * 2.42e10,
    "dsg_adapter": 6.70e13,
    "ntp_full": 3 * 2.859e13 + 3.5 * 1.100e12 + 3 * 2.152e12 + 3.65e11 + 3 * 2.42e10,  # IFT, SURVEY §8d
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ntp", choices=["ntp", "dsg"])
    ap.add_argument("--batch", type=int, default=8, help="per-GPU batch (samples per step per GPU)")
    ap.add_argument("--train", default="adapter", choices=["adapter", "full"],
                    help="adapter: PT-stage freeze policy (default, BASELINE configs[1]); full: IFT-style "
                         "full fine-tune of LLM + projector (finetune.sh) — weight gradients for every layer")
    ap.add_argument("--seq", type=int, default=2048, help="embedded sequence length T")
    ap.add_argument("--model", default="llama3-8b", choices=["llama3-8b", "phi3-mini", "tiny"])
    ap.add_argument("--tower", default="clip-vit-l", choices=["clip-vit-l", "convnext-xxl"],
                    help="vision tower: CLIP-ViT-L/14-336 (default, the BASELINE metric's config) or the frozen "
                         "CLIP-ConvNeXt-XXL at 768 px of BASELINE configs[3] (576 image tokens of width 3072)")
    ap.add_argument("--layers", type=int, default=None, help="override decoder depth (debug only; reported)")
    ap.add_argument("--torch-profile", action="store_true",
                    help="diagnostic: torch.profiler over 2 device-leg steps, prints kernel totals and busy time")
    ap.add_argument("--teachers", action="store_true",
                    help="dsg only: compute the depth / seg / gen targets with the on-GPU frozen teachers "
                         "(DINOv2-L, Swin-L, unCLIP ViT-H) from synthetic images every step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-input leg")
    ap.add_argument("--profile", action="store_true",
                    help="bracket the timed device leg with cudaProfilerStart/Stop (ncu --profile-from-start off)")
    ap.add_argument("--cpu-baseline-layers", type=int, default=1)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ helpers
CONVNEXT_XXL_FLOP = 3.563e12   # per 768 px image: stem + 3 downsamples + 40 blocks (fc1/fc2 GEMMs + 98 FLOP/elt depthwise)


def model_cfg(name, layers=None, tower="clip-vit-l"):
    c = _model_cfg(name, layers)
    if tower == "convnext-xxl":  # clip_convnext_encoder.py:61-174, timm convnext_xxlarge (norm_eps 1e-5)
        c.update(tower="convnext", cnx_depths=(3, 4, 30, 3), cnx_dims=(384, 768, 1536, 3072), cnx_eps=1e-5,
                 image_size=768)
    return c


def _model_cfg(name, layers=None):
    from visper_lm_b200.model import presets

    if name == "tiny":
        from oracle import configs as oc  # tiny debug config only (bench correctness smoke)

        c = dict(oc.TINY_LLAMA)
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


def host_batch(c, B, T, distill, seed, teachers=False):
    """SURVEY.md §8(d) synthetic batch in the collator's schema (HOST tensors)."""
    g = torch.Generator().manual_seed(seed)
    S, V = n_sys(c), c["vocab"]
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
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference(c, T, distill, layers_sampled, repeats=1, warmup=0, budget_s=None):
    """Times the oracle restatement (CPU fp32, all host threads) on a BOUNDED sample of the same
    workload: B=1, full CLIP tower + projector, `layers_sampled` of the decoder layers (fwd + bwd wrt
    activations, extrapolated linearly to all layers), final norm + full-vocab lm_head/CE fwd+bwd.
    Returns (samples_per_sec_estimate, description, per-repeat seconds)."""
    from oracle import restate

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    D, F, V, L = c["hidden"], c["inter"], c["vocab"], c["layers"]
    H, KVH = c["heads"], c["kv_heads"]
    hd = D // H
    g = torch.Generator().manual_seed(0)
    sd = {}

    def w(*shape, scale=0.02):
        return torch.empty(*shape).uniform_(-1.7 * scale, 1.7 * scale, generator=g)

    # one set of layer weights, aliased for every sampled layer (values do not affect timing)
    lw = {}
    if c["family"] == "phi3":
        lw = {"self_attn.qkv_proj.weight": w((H + 2 * KVH) * hd, D), "self_attn.o_proj.weight": w(D, D),
              "mlp.gate_up_proj.weight": w(2 * F, D), "mlp.down_proj.weight": w(D, F)}
    else:
        lw = {"self_attn.q_proj.weight": w(H * hd, D), "self_attn.k_proj.weight": w(KVH * hd, D),
              "self_attn.v_proj.weight": w(KVH * hd, D), "self_attn.o_proj.weight": w(D, D),
              "mlp.gate_proj.weight": w(F, D), "mlp.up_proj.weight": w(F, D), "mlp.down_proj.weight": w(D, F)}
    lw["input_layernorm.weight"] = torch.ones(D)
    lw["post_attention_layernorm.weight"] = torch.ones(D)
    for i in range(layers_sampled):
        for k, v in lw.items():
            sd[f"model.layers.{i}.{k}"] = v
    sd["model.norm.weight"] = torch.ones(D)
    sd["lm_head.weight"] = w(V, D)
    sd["model.embed_tokens.weight"] = sd["lm_head.weight"]
    convnext = c.get("tower") == "convnext"
    if convnext:  # one set of block weights per stage, aliased over its blocks
        cnx = dict(depths=c["cnx_depths"], dims=c["cnx_dims"], eps=c["cnx_eps"])
        stage_w = {}
        for k, shp in restate.convnext_state_spec(cnx).items():
            k0 = __import__("re").sub(r"blocks\.\d+\.", "blocks.0.", k)
            if k0 not in stage_w:
                stage_w[k0] = torch.ones(shp) if k.endswith(("norm.weight", "stem.1.weight", "downsample.0.weight")) \
                    else (torch.zeros(shp) if k.endswith("bias") else w(*shp))
            sd[k] = stage_w[k0]
    Dv, Fv = (c["cnx_dims"][-1], 0) if convnext else (c["vis_hidden"], c["vis_inter"])
    pv = "model.vision_tower.vision_tower.vision_model."
    sd[pv + "embeddings.patch_embedding.weight"] = w(Dv, 3, c["patch_size"], c["patch_size"])
    sd[pv + "embeddings.class_embedding"] = w(Dv)
    sd[pv + "embeddings.position_embedding.weight"] = w((c["image_size"] // c["patch_size"]) ** 2 + 1, Dv)
    vl = {}
    for nm, shp in () if convnext else (("self_attn.q_proj", (Dv, Dv)), ("self_attn.k_proj", (Dv, Dv)), ("self_attn.v_proj", (Dv, Dv)),
                    ("self_attn.out_proj", (Dv, Dv)), ("mlp.fc1", (Fv, Dv)), ("mlp.fc2", (Dv, Fv))):
        vl[nm + ".weight"] = w(*shp)
        vl[nm + ".bias"] = torch.zeros(shp[0])
    for nm in ("layer_norm1", "layer_norm2"):
        vl[nm + ".weight"], vl[nm + ".bias"] = torch.ones(Dv), torch.zeros(Dv)
    for i in range(0 if convnext else c["vis_layers"]):
        for k, v in vl.items():
            sd[f"{pv}encoder.layers.{i}.{k}"] = v
    for nm in ("pre_layrnorm",):
        sd[pv + nm + ".weight"], sd[pv + nm + ".bias"] = torch.ones(Dv), torch.zeros(Dv)
    sd["model.mm_projector.0.weight"] = w(D, Dv).requires_grad_(True)
    sd["model.mm_projector.0.bias"] = torch.zeros(D, requires_grad=True)
    sd["model.mm_projector.2.weight"] = w(D, D).requires_grad_(True)
    sd["model.mm_projector.2.bias"] = torch.zeros(D, requires_grad=True)
    cfg = dict(c, num_sys_tokens=n_sys(c), num_task_tokens=0)
    b = host_batch(c, 1, T, False, 1234)
    times = []
    t_begin = time.perf_counter()
    for r in range(warmup + repeats):
        if budget_s is not None and times and time.perf_counter() - t_begin > budget_s:
            break  # keep the whole run within a few minutes whatever --steps asks for
        t0 = time.perf_counter()
        with torch.no_grad():
            feats = restate.convnext_tower(sd, b["images"], cnx) if convnext else restate.clip_tower(sd, b["images"], cfg)
        img = restate.mm_projector(sd, feats)
        emb, labels, _ = restate.splice(sd, cfg, b["input_ids"], b["labels"], None, img)
        t1 = time.perf_counter()
        x = emb
        states = restate.decoder_stack(sd, cfg, x, None, n_layers=layers_sampled)
        h = states[-1]
        gl = torch.autograd.grad(h, emb, torch.ones_like(h), retain_graph=False)[0]
        t2 = time.perf_counter()
        hN = restate.rmsnorm(h.detach().requires_grad_(True), sd["model.norm.weight"])
        logits = torch.nn.functional.linear(hN, sd["lm_head.weight"]).float()
        loss = restate.ntp_loss(logits, labels)
        loss.backward()
        emb.backward(gl)  # projector wgrad/dgrad
        t3 = time.perf_counter()
        est = (t1 - t0) + (t2 - t1) * (L / layers_sampled) + (t3 - t2)
        if r >= warmup:
            times.append(est)
        for k in ("model.mm_projector.0.weight", "model.mm_projector.0.bias", "model.mm_projector.2.weight",
                  "model.mm_projector.2.bias"):
            sd[k].grad = None
    est = sum(times) / len(times)
    desc = (f"B=1, T={T}: full {'ConvNeXt-XXL @768' if convnext else 'CLIP'} tower fwd + mm_projector fwd/bwd + {layers_sampled}/{L} decoder layers "
            f"fwd+dgrad (x{L / layers_sampled:.0f} extrapolated) + final norm + full-vocab lm_head/CE fwd+bwd; "
            f"fp32 torch on {cores} threads")
    return 1.0 / est, desc, times, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c = model_cfg(args.model, args.layers, args.tower)
    distill = args.workload == "dsg"
    sps, desc, times, cores = cpu_reference(c, args.seq, distill, args.cpu_baseline_layers,
                                            repeats=max(1, args.steps), warmup=min(args.warmup, 1),
                                            budget_s=150.0)
    desc += f"; {len(times)} timed repeat(s) of the sample (150 s budget)"
    ms = 1000.0 / sps
    line = {
        "impl": "reference", "metric": "train-step samples/sec", "value": sps, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "steps_timed": len(times), "warmup": args.warmup,
        "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args, c, distill),
        "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, c, distill):
    label = "BASELINE configs[1]: " if not distill else "BASELINE configs[2] per-GPU slice: "
    if c.get("tower") == "convnext":
        label = ("BASELINE configs[3] per-GPU slice under ZeRO-2: " if distill
                 else "BASELINE configs[1] with configs[3]'s tower: ")
    return {"workload": label
            + f"{args.model} + " + ("CLIP-ConvNeXt-XXL, 768px" if c.get("tower") == "convnext" else "CLIP-ViT-L/14-336, 336px")
            + f", T={args.seq}, "
            + ("NTP only" if not distill else "NTP + dsg distill heads (d18-20_s10-18_g12-20)")
            + (", depth (DINOv2-L) / seg (Swin-L @800) / gen (unCLIP ViT-H) targets from the on-GPU frozen teachers each step"
               if distill and getattr(args, "teachers", False) else "")
            + (", full fine-tune (LLM + mm_projector trainable: fwd + dgrad + wgrad)" if getattr(args, "train", "adapter") == "full"
               else ", PT freeze policy (mm_projector" + ("+heads+task tokens" if distill else "") + " trainable; LLM fwd+dgrad)"),
            "per_gpu_batch": args.batch, "global_batch": args.batch * args.gpus, "seq_len": args.seq,
            "decoder_layers": c["layers"], "parallelism": f"dp{args.gpus} zero2",
            "l2_policy": "working set (16 GB weights + activations) exceeds the 126 MB L2; no explicit flush",
            "recompute": "none (activations kept; reference uses gradient checkpointing)"}


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    from visper_lm_b200 import lib, ops
    from visper_lm_b200.model import presets
    from visper_lm_b200 import model as pm
    from visper_lm_b200.train.trainer import LLaVATrainer, TrainingArguments

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    distill = args.workload == "dsg"
    c = model_cfg(args.model, args.layers, args.tower)
    cfg = presets.from_dict(c, distill=distill)
    fam = c["family"]
    cls = {("llama", True): pm.OlaLlavaLlamaForCausalLM, ("phi3", True): pm.OlaLlavaPhi3ForCausalLM,
           ("llama", False): pm.LlavaLlamaForCausalLM, ("phi3", False): pm.LlavaPhi3ForCausalLM}[(fam, distill)]
    torch.manual_seed(0)
    model = cls(cfg, device=dev)
    model.init_weights(std=0.02, seed=0)
    teachers = bool(args.teachers and distill)
    if teachers:
        cfg.random_init_teachers = True  # no network: DINOv2-L / Swin-L / unCLIP ViT-H geometry, random weights
        model.init_target_models(cfg)
    for n, p in model.named_parameters():
        if args.train == "full":  # finetune.sh: everything but the frozen tower (and the frozen DPT head / teacher)
            p.requires_grad_(("vision_tower" not in n) and ("da_v2_head" not in n) and ("dav2_backbone" not in n) and ("oneformer" not in n))
        else:  # PT freeze policy
            p.requires_grad_(("mm_projector" in n) or ("_heads." in n) or ("special_" in n) or n.endswith("logit_scale"))
    targs = TrainingArguments(per_device_train_batch_size=args.batch, learning_rate=1e-3, max_steps=10_000)
    trainer = LLaVATrainer(model=model, args=targs)
    trainer.total_steps = 10_000
    trainer.create_optimizer()

    B, T = args.batch, args.seq
    hb = host_batch(c, B, T, distill, 1234 + rank, teachers)
    # pinned host copies for the e2e leg
    pinned = {k: (v.pin_memory() if isinstance(v, torch.Tensor) else
                  {kk: vv.pin_memory() for kk, vv in v.items()}) for k, v in hb.items()}
    # device-resident copies for the kernel-side leg (ids stay on the host: the splice is planned there)
    devb = {k: (v if k in ("input_ids", "labels", "attention_mask") else
                (v.to(dev) if isinstance(v, torch.Tensor) else {kk: vv.to(dev) for kk, vv in v.items()}))
            for k, v in hb.items()}

    def fresh(b):
        out = dict(b)
        for k in ("depth_mask", "seg_mask", "gen_mask"):
            if k in out:
                out[k] = out[k].clone()
        return out

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(batch, steps, read_loss):
        sync_all()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        last = None
        for _ in range(steps):
            loss, _ = trainer.step(fresh(batch))
            if read_loss:
                last = float(loss)  # device→host read of the step's result
            else:
                last = loss
        e.record()
        sync_all()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), float(last)

    for _ in range(max(args.warmup, 3)):
        trainer.step(fresh(devb))
    sync_all()
    if args.torch_profile:
        from torch.profiler import ProfilerActivity, profile

        sync_all()
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            t0 = time.perf_counter()
            for _ in range(2):
                trainer.step(fresh(devb))
            t_cpu = time.perf_counter() - t0
            torch.cuda.synchronize()
            t_all = time.perf_counter() - t0
        evs = [e for e in prof.key_averages() if e.device_time_total > 0]
        busy = sum(e.device_time_total for e in evs) / 1e3
        print(json.dumps({"profile_steps": 2, "cpu_enqueue_ms": t_cpu * 1e3, "wall_ms": t_all * 1e3,
                          "gpu_busy_ms": busy}), flush=True)
        for e in sorted(evs, key=lambda e: -e.device_time_total)[:25]:
            print(json.dumps({"kernel": e.key[:70], "calls": e.count, "total_ms": round(e.device_time_total / 1e3, 2)}),
                  flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    lib.reset_launch_count()
    ops.GEMM_TIMER = ops.GemmTimer()
    if args.profile:
        torch.cuda.profiler.start()
    ms_dev, loss_dev = timed(devb, args.steps, read_loss=False)
    if args.profile:
        torch.cuda.profiler.stop()
    gemm_stats = ops.GEMM_TIMER.summary()
    ops.GEMM_TIMER = None
    launches = lib.launch_count()
    if args.no_e2e:
        ms_e2e, loss_e2e = ms_dev, loss_dev
    else:
        ms_e2e, loss_e2e = timed(pinned, args.steps, read_loss=True)
    clk = clocks.stop() if rank == 0 else None

    samples = B * world * args.steps
    value = samples / (ms_dev / 1e3)
    e2e = samples / (ms_e2e / 1e3)
    h2d = sum(v.numel() * v.element_size() for k, v in hb.items() if isinstance(v, torch.Tensor)
              and k not in ("input_ids", "attention_mask"))
    h2d += sum(vv.numel() * vv.element_size() for v in hb.values() if isinstance(v, dict) for vv in v.values())
    h2d += 3 * B * T * 4 + B * 576 * 4  # splice plan (kind/index/scatter int32) + inverse image map
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    pf = ROOT / "MEASURED_PEAKS.json"
    if pf.exists():
        peaks = json.loads(pf.read_text())
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s sustained (of fallback)"
    flop_key = "dsg_adapter" if distill else ("ntp_full" if args.train == "full" else "ntp_adapter")
    full_model = args.model == "llama3-8b" and args.layers is None and T == 2048
    step_flop = FLOPS[flop_key]
    if c.get("tower") == "convnext":  # tower forward and the 3072-wide projector input replace the ViT-L terms
        step_flop += (CONVNEXT_XXL_FLOP - 3.65e11) + 3 * (2 * 576 * (3072 - 1024) * 4096)
    line = {
        "metric": "train-step samples/sec", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic", "config": workload_config(args, c, distill),
        "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clk,
        "loss": {"device_leg": loss_dev, "e2e_leg": loss_e2e},
        "roofline": {
            "bound": "tensor",
            "kernel": "gemm_tcgen05_pair_kernel / gemm_tcgen05_kernel (all GEMM launches of the timed steps)",
            "achieved": gemm_stats["tflops"], "peak": peak_tf, "unit": "TFLOP/s",
            "frac": gemm_stats["tflops"] / peak_tf if gemm_stats["tflops"] else None,
            # dram read+write of the dominant launch shape (down_proj dgrad, M=16384 N=14336 K=4096) from
            # the ncu --set full capture profiles/r01_ncu_gemm_pair_shapes.csv (1.06 GB read + 0.45 GB
            # written); algorithmic: 0.25 GB of operands + 0.47 GB of output
            "traffic": 1.51e9 if full_model else None,
            "peak_source": peak_src, "launches": gemm_stats["launches"],
            "gemm_ms_per_step": gemm_stats["ms"] / args.steps,
            "gemm_share_of_step": gemm_stats["ms"] / ms_dev if ms_dev else None,
            "step_model_flops_frac": (value / world * step_flop / (peak_tf * 1e12)) if full_model else None,
        },
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            sps, desc, _, cores = cpu_reference(c, T, distill, args.cpu_baseline_layers, repeats=1, warmup=0)
            line["cpu_baseline"] = {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port", "sample": desc}
        except Exception as ex:  # the baseline is reported, never required for the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"failed: {type(ex).__name__}: {ex}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
