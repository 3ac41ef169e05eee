"""Data-parallel trainer for the hot path: the reference's `LLaVATrainer` constructor / `train()` /
`training_step` surface (ola_vlm/train/llava_trainer.py:217, ola_vlm_train.py:1299-1310) over a
B200-native step — one process per GPU, NCCL over NVLink for the three collectives of ZeRO-2
(scripts/zero2.json): gradient reduce-scatter, sharded fused AdamW on fp32 masters, bf16 parameter
all-gather.  HF Trainer / accelerate / DeepSpeed are not used (and not installed here).

Step semantics follow SURVEY.md §8a "Optimizer/step semantics": AdamW β=(0.9,0.999) eps 1e-8,
weight-decay / lr groups of create_optimizer (llava_trainer.py:890-995), cosine schedule with linear
warm-up, loss = rank-local batch mean, gradients averaged over ranks, optional global-norm clip.
"""
from __future__ import annotations

import json
import math
import os
import time
from dataclasses import dataclass, field
from typing import Optional

import torch
import torch.distributed as dist

from .. import ops
from ..ops import BF16


@dataclass
class TrainingArguments:
    """Subset of HF TrainingArguments + the reference's extras (ola_vlm_train.py:122-155) that the
    step itself consumes."""
    output_dir: str = "./checkpoints"
    per_device_train_batch_size: int = 8
    gradient_accumulation_steps: int = 1
    learning_rate: float = 1e-3
    weight_decay: float = 0.0
    adam_beta1: float = 0.9
    adam_beta2: float = 0.999
    adam_epsilon: float = 1e-8
    max_grad_norm: float = 1.0
    warmup_ratio: float = 0.03
    lr_scheduler_type: str = "cosine"
    num_train_epochs: float = 1.0
    max_steps: int = -1
    logging_steps: int = 1
    save_steps: int = 200
    save_total_limit: Optional[int] = None
    tune_mm_mlp_adapter: bool = False
    use_im_start_end: bool = False
    bf16: bool = True
    mm_projector_lr: Optional[float] = None
    mm_vision_lr: Optional[float] = None
    group_by_modality_length: bool = False
    model_max_length: int = 4096
    dataloader_num_workers: int = 0
    dataloader_prefetch_factor: int = 2
    zero_stage: int = 2
    seed: int = 42


def cosine_with_warmup(step, total, warmup):
    """HF get_cosine_schedule_with_warmup multiplier."""
    if step < warmup:
        return step / max(1, warmup)
    prog = (step - warmup) / max(1, total - warmup)
    return max(0.0, 0.5 * (1.0 + math.cos(math.pi * prog)))


class Zero2Optimizer:
    """ZeRO-2: every rank holds all bf16 parameters (one flat buffer) and the full bf16 gradient
    buffer; fp32 master weights and Adam moments exist only for the rank's 1/N shard.

    groups: list of (predicate(name) -> bool, lr_scale, weight_decay); first match wins."""

    ALIGN = 8

    def __init__(self, named_params, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                 max_grad_norm=1.0, groups=None, process_group=None):
        self.named = [(n, p) for n, p in named_params if p.requires_grad]
        assert self.named, "no trainable parameters"
        self.lr, self.betas, self.eps, self.max_grad_norm = lr, betas, eps, max_grad_norm
        self.pg = process_group
        self.world = dist.get_world_size(self.pg) if dist.is_initialized() else 1
        self.rank = dist.get_rank(self.pg) if dist.is_initialized() else 0
        dev = self.named[0][1].device
        groups = groups or [(lambda n: True, 1.0, weight_decay)]
        # layout: parameters sorted by group so each group is one contiguous range
        order = []
        for gi, (pred, _, _) in enumerate(groups):
            for n, p in self.named:
                if not any(n == o[0] for o in order) and pred(n) and \
                        not any(g[0](n) for g in groups[:gi]):
                    order.append((n, p, gi))
        self.named = [(n, p) for n, p, _ in order]
        offs, off = [], 0
        self.group_ranges = []
        cur_g, g_start = order[0][2], 0
        for n, p, gi in order:
            if gi != cur_g:
                self.group_ranges.append((g_start, off, cur_g))
                cur_g, g_start = gi, off
            offs.append(off)
            off += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.group_ranges.append((g_start, off, cur_g))
        self.groups = groups
        chunk = self.world * 1024
        self.total = (off + chunk - 1) // chunk * chunk
        self.offsets = offs
        self.flat_p = torch.zeros(self.total, dtype=BF16, device=dev)
        self.flat_g = torch.zeros(self.total, dtype=BF16, device=dev)
        with torch.no_grad():
            for (n, p), o in zip(self.named, offs):
                view = self.flat_p[o:o + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view  # parameters now live in the flat buffer (dtype becomes bf16)
        self.shard = self.total // self.world
        s0 = self.rank * self.shard
        self.s0, self.s1 = s0, s0 + self.shard
        self.master = self.flat_p[s0:s0 + self.shard].float()
        self.m = torch.zeros(self.shard, dtype=torch.float32, device=dev)
        self.v = torch.zeros(self.shard, dtype=torch.float32, device=dev)
        self.g_shard = torch.zeros(self.shard, dtype=BF16, device=dev) if self.world > 1 else None
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self.step_count = 0
        self.last_grad_norm = None

    def zero_grad(self):
        for _, p in self.named:
            p.grad = None

    def _pack_grads(self):
        for (n, p), o in zip(self.named, self.offsets):
            dst = self.flat_g[o:o + p.numel()]
            if p.grad is None:
                dst.zero_()
                continue
            g = p.grad
            if g.dtype == BF16 and g.is_contiguous() and g.numel() % 8 == 0:
                ops.axpby(g.reshape(-1), None, 1.0, 0.0, out=dst)
            else:
                dst.copy_(g.reshape(-1))

    def step(self, lr_mult=1.0):
        self.step_count += 1
        self._pack_grads()
        if self.world > 1:
            dist.reduce_scatter_tensor(self.g_shard, self.flat_g, op=dist.ReduceOp.SUM, group=self.pg)
            g = self.g_shard
        else:
            g = self.flat_g
        ops.grad_sumsq(g, out=self.sumsq)
        if self.world > 1:
            dist.all_reduce(self.sumsq, group=self.pg)
        # gradients are SUMS over ranks here → average with 1/world inside the clip coefficient
        coef, norm = ops.clip_coef(self.sumsq, self.max_grad_norm or 0.0, 1.0 / self.world)
        self.last_grad_norm = norm
        for (a, b, gi) in self.group_ranges:
            lo, hi = max(a, self.s0), min(b, self.s1)
            if lo >= hi:
                continue
            _, lr_scale, wd = self.groups[gi]
            sl = slice(lo - self.s0, hi - self.s0)
            ops.adamw_step_(self.master[sl], self.m[sl], self.v[sl], g[sl], self.flat_p[lo:hi],
                            self.lr * lr_scale * lr_mult, self.betas[0], self.betas[1], self.eps, wd,
                            self.step_count, grad_scale=coef)
        if self.world > 1:
            dist.all_gather_into_tensor(self.flat_p, self.flat_p[self.s0:self.s1], group=self.pg)

    def state_dict(self):
        return {"step": self.step_count, "master": self.master, "m": self.m, "v": self.v,
                "rank": self.rank, "world": self.world}


def _no_decay(name):
    return name.endswith(".bias") or "norm" in name or "layernorm" in name


class LLaVATrainer:
    """Constructor and entry points of ola_vlm/train/llava_trainer.py:217 (HF Trainer subclass)."""

    def __init__(self, model=None, tokenizer=None, args: TrainingArguments = None, train_dataset=None,
                 eval_dataset=None, data_collator=None, **kwargs):
        self.model = model
        self.tokenizer = tokenizer
        self.args = args or TrainingArguments()
        self.train_dataset = train_dataset
        self.eval_dataset = eval_dataset
        self.data_collator = data_collator
        self.deepspeed = None
        self.optimizer = None
        self.state = {"global_step": 0, "log_history": []}
        self.is_dist = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size() if self.is_dist else 1
        self.rank = dist.get_rank() if self.is_dist else 0
        self.total_steps = None
        self._pinned = {}

    # -- optimizer groups: decay / no-decay × projector-lr (llava_trainer.py:903-976) ---------------
    def create_optimizer(self):
        if self.optimizer is not None:
            return self.optimizer
        a = self.args
        groups = []
        if a.mm_projector_lr is not None:
            scale = a.mm_projector_lr / a.learning_rate
            groups.append((lambda n: "mm_projector" in n and not _no_decay(n), scale, a.weight_decay))
            groups.append((lambda n: "mm_projector" in n and _no_decay(n), scale, 0.0))
        groups.append((lambda n: not _no_decay(n), 1.0, a.weight_decay))
        groups.append((lambda n: True, 1.0, 0.0))
        self.optimizer = Zero2Optimizer(self.model.named_parameters(), a.learning_rate,
                                        (a.adam_beta1, a.adam_beta2), a.adam_epsilon, a.weight_decay,
                                        a.max_grad_norm, groups)
        return self.optimizer

    # -- one optimisation step on a HOST batch (collator output) ------------------------------------
    def _to_device(self, batch):
        dev = self.model.device
        out = {}
        for k, v in batch.items():
            if isinstance(v, torch.Tensor):
                if k in ("input_ids", "labels", "attention_mask"):
                    out[k] = v  # consumed on the host by the splice planner
                elif v.is_cuda:
                    out[k] = v
                elif dev.type != "cuda":
                    out[k] = v  # host-only unit tests of the trainer logic
                else:
                    if not v.is_pinned():
                        buf = self._pinned.get(k)
                        if buf is None or buf.shape != v.shape or buf.dtype != v.dtype:
                            buf = torch.empty(v.shape, dtype=v.dtype).pin_memory()
                            self._pinned[k] = buf
                        buf.copy_(v)
                        v = buf
                    out[k] = v.to(dev, non_blocking=True)
            elif isinstance(v, dict):
                out[k] = {kk: (vv if vv.is_cuda else vv.to(dev, non_blocking=True)) for kk, vv in v.items()}
            else:
                out[k] = v
        return out

    def compute_loss(self, model, inputs, return_outputs=False, **_):
        """HF Trainer.compute_loss as the reference relies on it: `model(**inputs)` and the
        `loss` field (or element 0 of the tuple form)."""
        out = model(**self._to_device(inputs))
        loss = out[0] if isinstance(out, tuple) else out.loss
        return (loss, out) if return_outputs else loss

    def _get_train_sampler(self):
        """llava_trainer.py:219-232: the modality-length-grouped sampler under --group_by_modality_length
        (batch = per-device batch, world = world_size x grad-accum), else a seeded random sampler.  train()
        consumes the same order through _index_order."""
        ds = self.train_dataset
        if ds is None or not hasattr(ds, "__len__"):
            return None
        a = self.args
        if a.group_by_modality_length:
            from .data import LengthGroupedSampler

            return LengthGroupedSampler(a.per_device_train_batch_size, self.world * a.gradient_accumulation_steps,
                                        lengths=ds.modality_lengths, group_by_modality=True)
        return torch.utils.data.RandomSampler(ds, generator=torch.Generator().manual_seed(a.seed))

    def training_step(self, model, inputs):
        """HF Trainer.training_step semantics (llava_trainer.py:357-381 is a dead verbatim copy):
        forward → loss → backward; returns the detached loss tensor (no host sync)."""
        inputs = self._to_device(inputs)
        out = model(**inputs)
        loss = out.loss
        loss.backward()
        return loss.detach(), out

    def step(self, inputs):
        opt = self.create_optimizer()
        opt.zero_grad()
        loss, out = self.training_step(self.model, inputs)
        a = self.args
        total = self.total_steps or max(1, a.max_steps)
        warm = math.ceil(total * a.warmup_ratio)
        mult = cosine_with_warmup(self.state["global_step"], total, warm) if a.lr_scheduler_type == "cosine" else 1.0
        if self.total_steps is None and a.max_steps <= 0:
            mult = 1.0
        opt.step(lr_mult=mult)
        self.state["global_step"] += 1
        return loss, out

    def accumulated_step(self, micro_batches):
        """gradient_accumulation_steps > 1 (HF Trainer semantics): each micro-batch's loss is divided by
        the number of micro-batches before backward, gradients add up in p.grad, one optimizer step."""
        opt = self.create_optimizer()
        opt.zero_grad()
        k = len(micro_batches)
        total = None
        for mb in micro_batches:
            out = self.model(**self._to_device(mb))
            (out.loss / k).backward()
            total = out.loss.detach() / k if total is None else total + out.loss.detach() / k
        a = self.args
        steps = self.total_steps or max(1, a.max_steps)
        warm = math.ceil(steps * a.warmup_ratio)
        mult = cosine_with_warmup(self.state["global_step"], steps, warm) if a.lr_scheduler_type == "cosine" else 1.0
        if self.total_steps is None and a.max_steps <= 0:
            mult = 1.0
        opt.step(lr_mult=mult)
        self.state["global_step"] += 1
        return total, out

    def _index_order(self, epoch):
        """Global sample order of one epoch.  --group_by_modality_length → the reference's
        LengthGroupedSampler (llava_trainer.py:219-232) with batch = per-device batch and world =
        world_size × grad-accum; otherwise HF's seeded RandomSampler (randperm of seed + epoch)."""
        a = self.args
        n = len(self.train_dataset)
        g = torch.Generator().manual_seed(a.seed + epoch)
        if a.group_by_modality_length and hasattr(self.train_dataset, "modality_lengths"):
            from .data import LengthGroupedSampler
            # the per-modality grouping draws from the GLOBAL torch RNG (as in the reference, where
            # set_seed(args.seed) precedes it): pin it to (seed, epoch) so a resumed run sees the
            # same order, and leave the caller's RNG stream untouched
            keep = torch.get_rng_state()
            torch.manual_seed(a.seed + epoch)
            try:
                return list(LengthGroupedSampler(a.per_device_train_batch_size,
                                                 self.world * a.gradient_accumulation_steps,
                                                 lengths=self.train_dataset.modality_lengths, generator=g,
                                                 group_by_modality=True))
            finally:
                torch.set_rng_state(keep)
        return torch.randperm(n, generator=g).tolist()

    def _batches(self, epoch=0, skip=0):
        """Per-rank batches: consecutive per-device batches of the global order are dealt to the
        ranks round-robin (what accelerate's BatchSamplerShard does under HF Trainer), so one
        length-grouped mega-batch spreads its balanced chunks over all ranks."""
        a = self.args
        B = a.per_device_train_batch_size
        order = self._index_order(epoch)
        nb = len(order) // (B * self.world)  # drop the ragged tail (dataloader_drop_last semantics)
        ga = max(1, int(a.gradient_accumulation_steps))
        nb -= nb % ga                        # whole optimizer steps only
        skip *= ga                           # `skip` counts optimizer steps
        def build(k):
            s = (k * self.world + self.rank) * B
            return self.data_collator([self.train_dataset[j] for j in order[s:s + B]])

        workers = int(getattr(a, "dataloader_num_workers", 0) or 0)
        if workers <= 0:
            for k in range(skip, nb):
                yield build(k)
            return
        # dataloader_num_workers (pretrain.sh: 4): image decode / preprocessing / collation of the next
        # batches runs on host threads while the GPU executes the current step; batches come back in order
        from collections import deque
        from concurrent.futures import ThreadPoolExecutor

        depth = workers * int(getattr(a, "dataloader_prefetch_factor", 2) or 2)
        with ThreadPoolExecutor(max_workers=workers, thread_name_prefix="vpb-data") as pool:
            pending = deque()
            nxt = skip
            try:
                while nxt < nb or pending:
                    while nxt < nb and len(pending) < depth:
                        pending.append(pool.submit(build, nxt))
                        nxt += 1
                    yield pending.popleft().result()
            finally:
                for f in pending:
                    f.cancel()

    def steps_per_epoch(self):
        return len(self.train_dataset) // (self.args.per_device_train_batch_size * self.world
                                           * max(1, int(self.args.gradient_accumulation_steps)))

    def train(self, resume_from_checkpoint=None):
        from . import checkpoint as ckpt

        a = self.args
        per_epoch = max(1, self.steps_per_epoch())
        self.total_steps = a.max_steps if a.max_steps > 0 else int(per_epoch * a.num_train_epochs)
        self.create_optimizer()
        if resume_from_checkpoint:
            path = (ckpt.get_last_checkpoint(a.output_dir) if resume_from_checkpoint is True
                    else resume_from_checkpoint)
            if path is None:
                raise ValueError(f"no checkpoint-* directory under {a.output_dir}")
            ckpt.load_checkpoint(self, path)
        t0 = time.time()
        last = None
        done = self.state["global_step"]
        while done < self.total_steps:
            epoch, skip = divmod(done, per_epoch)
            ga = max(1, int(a.gradient_accumulation_steps))
            micro = []
            for batch in self._batches(epoch, skip):  # a resumed run skips the batches already consumed
                if ga == 1:
                    loss, _ = self.step(batch)
                else:
                    micro.append(batch)
                    if len(micro) < ga:
                        continue
                    loss, _ = self.accumulated_step(micro)
                    micro = []
                done += 1
                if done % a.logging_steps == 0:
                    last = float(loss)  # the only host sync, outside forward (cf. ola_llama.py:146-168)
                    self.state["log_history"].append({"step": done, "loss": last})
                if a.save_steps and a.save_steps > 0 and done % a.save_steps == 0:
                    self._save_checkpoint()
                if done >= self.total_steps:
                    break
        return {"global_step": done, "training_loss": last, "train_runtime": time.time() - t0}

    # -- checkpoint surface (llava_trainer.py:997-1021, ola_vlm_train.py:228-263) -----
    def _save_checkpoint(self, model=None, trial=None, metrics=None):
        from . import checkpoint as ckpt

        out = os.path.join(self.args.output_dir, f"{ckpt.PREFIX_CHECKPOINT_DIR}-{self.state['global_step']}")
        ckpt.save_checkpoint(self, out)
        if self.is_dist:
            dist.barrier()
        if self.rank == 0:
            ckpt.rotate_checkpoints(self.args.output_dir, self.args.save_total_limit)
        return out

    def save_state(self):
        os.makedirs(self.args.output_dir, exist_ok=True)
        if self.rank == 0:
            with open(os.path.join(self.args.output_dir, "trainer_state.json"), "w") as f:
                json.dump(self.state, f)

    def save_model(self, output_dir=None):
        self._save(output_dir)

    def _save(self, output_dir=None, state_dict=None):
        """Full-model save used by safe_save_model_for_hf_trainer's non-adapter branch."""
        output_dir = output_dir or self.args.output_dir
        os.makedirs(output_dir, exist_ok=True)
        if self.rank == 0:
            from . import checkpoint as ckpt

            sd = state_dict if state_dict is not None else {k: v.detach().cpu() for k, v in self.model.state_dict().items()}
            ckpt.save_config(self.model.config, output_dir)
            # HF save_pretrained layout (safetensors, 5 GB shards + index) so builder.py / from_pretrained load it
            ckpt.save_pretrained_weights(sd, output_dir, getattr(self.args, "max_shard_size", "5GB"))
