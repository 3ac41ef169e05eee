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
    zero_bucket_elems: int = 64 * 1024 * 1024   # ZeRO-2 reduce / gather bucket (scripts/zero2.json reduce_bucket_size)
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


class _Bucket:
    """A contiguous range [lo, hi) of the flat parameter space that is reduced / gathered as one collective.
    hi - lo is a multiple of world·ALIGN; rank r owns [lo + r·slice, lo + (r+1)·slice)."""

    __slots__ = ("idx", "lo", "hi", "slice", "shard_off", "group", "params", "pads", "pending", "buf",
                 "ev_reduced", "ev_gathered", "launched", "touched")

    def __init__(self, idx, lo, group):
        self.idx, self.lo, self.hi, self.group = idx, lo, lo, group
        self.slice = self.shard_off = 0
        self.params, self.pads = [], []
        self.pending = 0
        self.buf = self.ev_reduced = self.ev_gathered = None
        self.launched = self.touched = False


class LayerGradSink:
    """Lets a hand-written backward write a weight gradient straight into the optimizer's gradient
    buffer (no `p.grad` twin, no per-parameter copy): `dst(key)` → ([rows, K] view, accumulate?) or None
    when that key is not (entirely) trainable; `done(key)` after the kernel that wrote it was launched."""

    def __init__(self, opt, keys):
        self.opt, self.keys = opt, keys          # key -> (first param index, [param indices], (rows, K))

    def dst(self, key):
        ent = self.keys.get(key)
        if ent is None or not self.opt.sinks_enabled:
            return None
        first, idxs, shape = ent
        buf, acc = self.opt._grad_dst(first, sum(self.opt.numels[i] for i in idxs), idxs)
        return buf.view(shape), acc

    def done(self, key):
        for i in self.keys[key][1]:
            self.opt._mark_written(i)


class Zero2Optimizer:
    """ZeRO-2 (scripts/zero2.json: contiguous gradients, bucketed reduce, overlap): every rank holds all
    bf16 parameters in one flat buffer; gradients, fp32 master weights and Adam moments exist only for the
    rank's 1/N share.

    Layout.  Trainable parameters are laid out by optimizer group (decay / no-decay × lr scale), in
    model order inside a group, and cut into BUCKETS of ~bucket_elems elements (a fused q|k|v or gate|up
    group is never split and stays adjacent, so modules.FusedRows adopts its region of the flat buffer
    instead of owning a copy).  Rank r owns slice r of every bucket; its shard (master, m, v, gradient)
    is the concatenation of those slices in bucket order.

    Gradients.  Producers write into the gradient space directly: the decoder backward through a
    LayerGradSink (the wgrad GEMM's output pointer), everything else through a post-accumulate hook
    that moves `p.grad` there and drops it.  With world > 1 a bucket's gradient space is a staging
    buffer allocated at its first gradient: when the bucket's last gradient is written — backward runs the
    buckets in reverse order — the comm stream waits for that point and reduce-scatters the bucket into
    the rank's gradient shard, so the transfer hides under the rest of backward and full-size gradient
    memory never exists (a staging buffer is freed to the caching allocator as soon as its reduce is launched).
    With world == 1 the gradient shard is the gradient space.

    step(): wait for the reduces, global grad-norm (fp32 scalar all-reduce), clip coefficient on device,
    fused AdamW on the shard (fp32 master → bf16 parameter slice), then the per-bucket parameter
    all-gathers on the comm stream; the next forward waits for them only where it first touches a
    trainable parameter (`wait_params`, after the frozen tower).

    groups: list of (predicate(name) -> bool, lr_scale, weight_decay); first match wins."""

    ALIGN = 8

    def __init__(self, named_params, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                 max_grad_norm=1.0, groups=None, process_group=None, distributed=True,
                 bucket_elems=64 * 1024 * 1024, keep_together=(), overlap=True, use_order=None):
        self.named = [(n, p) for n, p in named_params if p.requires_grad]
        assert self.named, "no trainable parameters"
        self.lr, self.betas, self.eps, self.max_grad_norm = lr, betas, eps, max_grad_norm
        self.pg = process_group
        use_dist = distributed and dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(self.pg) if use_dist else 1
        self.rank = dist.get_rank(self.pg) if use_dist else 0
        dev = self.named[0][1].device
        self.dev = dev
        groups = groups or [(lambda n: True, 1.0, weight_decay)]
        self.groups = groups
        # ---- order: by group, model order inside a group ------------------------------------------------
        order, seen = [], set()
        for gi, (pred, _, _) in enumerate(groups):
            for n, p in self.named:
                if n not in seen and pred(n):
                    seen.add(n)
                    order.append((n, p, gi))
        self.named = [(n, p) for n, p, _ in order]
        self.index = {id(p): i for i, (_, p) in enumerate(self.named)}
        self.numels = [p.numel() for _, p in self.named]
        together = {}
        for grp in keep_together:  # parameters that must stay adjacent and in one bucket (FusedRows groups)
            ids = [self.index[id(p)] for p in grp if id(p) in self.index]
            if len(ids) == len(list(grp)) and ids == list(range(ids[0], ids[0] + len(ids))) \
                    and all(self.numels[i] % self.ALIGN == 0 for i in ids):
                for k in ids[1:]:
                    together[k] = ids[0]
        # ---- buckets ----------------------------------------------------------------------------------
        unit = self.world * self.ALIGN
        offs, off, buckets, cur = [], 0, [], None

        def close(b, off):
            end = (off + unit - 1) // unit * unit
            if end > off:
                b.pads.append((off, end))
            b.hi = end
            return end

        for i, (n, p, gi) in enumerate(order):
            if cur is None or gi != cur.group or (cur.hi - cur.lo >= bucket_elems and i not in together):
                if cur is not None:
                    off = close(cur, off)
                cur = _Bucket(len(buckets), off, gi)
                buckets.append(cur)
            offs.append(off)
            cur.params.append(i)
            nxt = off + (self.numels[i] + self.ALIGN - 1) // self.ALIGN * self.ALIGN
            if nxt > off + self.numels[i]:
                cur.pads.append((off + self.numels[i], nxt))
            off = nxt
            cur.hi = off
        off = close(cur, off)
        self.total = off
        self.offsets = offs
        self.buckets = buckets
        self.bucket_of = [None] * len(order)
        shard_off = 0
        for b in buckets:
            b.slice = (b.hi - b.lo) // self.world
            b.shard_off = shard_off
            shard_off += b.slice
            for i in b.params:
                self.bucket_of[i] = b
        self.shard = shard_off
        assert self.shard * self.world == self.total
        # ---- storage ----------------------------------------------------------------------------------
        self.flat_p = torch.zeros(self.total, dtype=BF16, device=dev)
        with torch.no_grad():
            for (n, p), o in zip(self.named, offs):
                view = self.flat_p[o:o + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view  # parameters now live in the flat buffer (dtype becomes bf16)
                p._vpb_flat_owner = True
        self.master = torch.cat([self.flat_p[b.lo + self.rank * b.slice: b.lo + (self.rank + 1) * b.slice]
                                 for b in buckets]).float()
        self.m = torch.zeros(self.shard, dtype=torch.float32, device=dev)
        self.v = torch.zeros(self.shard, dtype=torch.float32, device=dev)
        self.g_shard = torch.zeros(self.shard, dtype=BF16, device=dev)
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self.step_count = 0
        self.last_grad_norm = None
        # AdamW runs: contiguous in shard space AND in parameter space, one optimizer group each
        runs = []
        for b in buckets:
            plo = b.lo + self.rank * b.slice
            if runs and runs[-1][3] == b.group and runs[-1][1] == b.shard_off and runs[-1][2] + (runs[-1][1] - runs[-1][0]) == plo:
                runs[-1][1] = b.shard_off + b.slice
            else:
                runs.append([b.shard_off, b.shard_off + b.slice, plo, b.group])
        self.update_runs = [tuple(r) for r in runs]
        # ---- gradient plumbing --------------------------------------------------------------------------
        self.cuda = dev.type == "cuda"
        # overlap="force": exercise the staging pool / comm-stream path on ONE GPU (tests): the "reduce-scatter"
        # of world 1 is a copy of the bucket into the gradient shard on the comm stream
        self.force_staging = overlap == "force" and self.cuda and self.world == 1
        self.overlap = bool(overlap and (self.world > 1 or self.force_staging) and self.cuda)
        self.sinks_enabled = self.cuda
        self.accumulating = False
        self.flat_g = None                     # full-size gradient space: only world > 1 without overlap
        self.comm = torch.cuda.Stream(device=dev) if (self.cuda and (self.world > 1 or self.force_staging)) else None
        self.staging_peak = 0                  # most gradient-staging bytes alive at one time (overlap mode)
        self.written = [False] * len(order)
        # parameters that received no gradient in the previous step (e.g. linear_2/3 of the depth heads, which the
        # reference's loss never reaches): they are not waited for, so their buckets can be reduced during backward
        self.expected_missing = set()
        self._pending_gather = []
        # use_order(name) -> rank of first use in the forward pass: the parameter all-gathers are issued in that
        # order and a consumer waits only for the buckets up to its own (wait_params(upto=...))
        key = use_order or (lambda n: 0)
        self.gather_order = sorted(self.buckets, key=lambda b: (min(key(self.named[i][0]) for i in b.params), b.idx))
        self.gather_pos = {b.idx: k for k, b in enumerate(self.gather_order)}
        self.use_rank = sorted({key(self.named[i][0]) for i in range(len(self.named))})
        self._rank_last_pos = {}
        for r in self.use_rank:   # position (in issue order) of the last bucket holding a parameter used at rank <= r
            self._rank_last_pos[r] = max(self.gather_pos[self.bucket_of[i].idx] for i in range(len(self.named))
                                         if key(self.named[i][0]) <= r)
        self._events = {}
        self._gather_waited = 0
        self._gather_wait_pairs = []
        self._hooks = []
        if hasattr(torch.Tensor, "register_post_accumulate_grad_hook"):
            for _, p in self.named:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self._reset_step_state()

    # ---- layout helpers -----------------------------------------------------------------------------------
    def flat_params(self):
        """All bf16 parameters as one flat tensor (waits for a pending parameter all-gather)."""
        self.wait_params()
        return self.flat_p

    def layer_sink(self, keyed_params):
        """keyed_params: {key: [Parameter, ...]} (adjacent rows of one fused weight).  Keys whose parameters
        are not all trainable and adjacent are left out (their gradients take the hook path)."""
        keys = {}
        for key, ps in keyed_params.items():
            idxs = [self.index.get(id(p)) for p in ps]
            if any(i is None for i in idxs) or idxs != list(range(idxs[0], idxs[0] + len(idxs))):
                continue
            if any(self.offsets[i + 1] != self.offsets[i] + self.numels[i] for i in idxs[:-1]):
                continue
            if len({self.bucket_of[i].idx for i in idxs}) != 1:
                continue
            rows = sum(self.named[i][1].shape[0] for i in idxs)
            keys[key] = (idxs[0], idxs, (rows, self.named[idxs[0]][1].shape[1]))
        return LayerGradSink(self, keys) if keys else None

    # ---- per-step gradient collection ---------------------------------------------------------------------
    def _reset_step_state(self):
        self.written = [False] * len(self.named)
        miss = self.expected_missing
        for b in self.buckets:
            b.pending = sum(1 for i in b.params if i not in miss)
            b.launched = b.touched = False
            b.buf = None

    def zero_grad(self):
        for _, p in self.named:
            p.grad = None
        self._reset_step_state()
        self.accumulating = False

    def set_accumulating(self, flag=True):
        """gradient_accumulation_steps > 1: gradients add up over several backward passes before step();
        the bucketed overlap (one reduce per bucket per backward) is switched off for such steps."""
        self.accumulating = bool(flag)

    def _use_staging(self):
        return self.overlap and not self.accumulating

    def _grad_space(self, b):
        """The tensor that holds bucket b's gradient space [b.lo, b.hi) for this step."""
        if self.world == 1 and not (self.force_staging and self._use_staging()):
            return self.g_shard[b.shard_off: b.shard_off + b.slice]
        if not self._use_staging():
            if self.flat_g is None:
                self.flat_g = torch.zeros(self.total, dtype=BF16, device=self.dev)
            return self.flat_g[b.lo:b.hi]
        if b.buf is None:
            # a staging buffer lives from the bucket's first gradient to the launch of its reduce-scatter; the
            # caching allocator hands the block to a later bucket only after the comm stream is done with it
            # (record_stream in _reduce_bucket), so gradient memory is bounded by the buckets open at one time
            b.buf = torch.empty(b.hi - b.lo, dtype=BF16, device=self.dev)
            self.staging_peak = max(self.staging_peak, self._staging_live() )
        return b.buf

    def _staging_live(self):
        return sum((x.hi - x.lo) * 2 for x in self.buckets if x.buf is not None)

    def _touch(self, b):
        space = self._grad_space(b)
        if not b.touched:
            b.touched = True
            if self._use_staging():
                for lo, hi in b.pads:          # padding never receives a gradient: keep it zero for the norm
                    space[lo - b.lo: hi - b.lo].zero_()
            for i in b.params:                 # predicted to get no gradient: zero now, the bucket does not wait for it
                if i in self.expected_missing:
                    o = self.offsets[i] - b.lo
                    space[o:o + self.numels[i]].zero_()
        return space

    def _grad_dst(self, first, n, idxs):
        b = self.bucket_of[first]
        space = self._touch(b)
        o = self.offsets[first] - b.lo
        acc = all(self.written[i] for i in idxs)
        if not acc:
            for i in idxs:
                if self.written[i]:           # partially written group: cannot happen with whole-group sinks
                    raise RuntimeError("gradient sink: group written piecewise")
        return space[o:o + n], acc

    def _mark_written(self, i):
        if self.written[i]:
            return
        self.written[i] = True
        b = self.bucket_of[i]
        if i in self.expected_missing:
            if b.launched:
                raise RuntimeError(f"gradient of {self.named[i][0]} arrived after its bucket was reduced: the set of "
                                   "parameters that receive gradients changed between steps (call "
                                   "optimizer.expected_missing.clear() when switching workloads)")
            return                             # pre-zeroed at first touch, now overwritten; it was never counted
        b.pending -= 1
        if b.pending == 0 and self._use_staging() and not b.launched:
            self._reduce_bucket(b)

    def _collect(self, i, g):
        """Move one parameter's gradient into the gradient space (generic path)."""
        n = self.numels[i]
        dst, acc = self._grad_dst(i, n, [i])
        g = g.reshape(-1)
        if g.dtype == BF16 and g.is_contiguous() and n % 8 == 0:
            ops.axpby(g, dst if acc else None, 1.0, 1.0 if acc else 0.0, out=dst)
        elif acc:
            dst.add_(g.to(dst.dtype))
        else:
            dst.copy_(g)
        self._mark_written(i)

    def _on_grad(self, p):
        i = self.index.get(id(p))
        if i is None or p.grad is None:
            return
        self._collect(i, p.grad)
        p.grad = None

    def _reduce_bucket(self, b):
        b.launched = True
        if self.world == 1 and not (self.force_staging and self._use_staging()):
            return
        src = self._grad_space(b)
        out = self.g_shard[b.shard_off: b.shard_off + b.slice]
        if self.comm is None:                 # CPU (gloo) path of the tests: synchronous
            self._reduce_scatter_cpu(out, src)
            return
        ready = torch.cuda.Event()
        ready.record()
        self.comm.wait_event(ready)
        with torch.cuda.stream(self.comm):
            if self.world == 1:
                out.copy_(src)
            else:
                dist.reduce_scatter_tensor(out, src, op=dist.ReduceOp.SUM, group=self.pg)
            b.ev_reduced = torch.cuda.Event()
            b.ev_reduced.record()
        if b.buf is not None:
            b.buf.record_stream(self.comm)
            b.buf = None                      # the block returns to the allocator once the comm stream has read it

    def _reduce_scatter_cpu(self, out, src):
        # gloo has no reduce_scatter: all-reduce the bucket (fp32 sum rounded to bf16, as NCCL's bf16 sum) and slice
        t = src.clone()
        dist.all_reduce(t, group=self.pg)
        out.copy_(t[self.rank * out.numel():(self.rank + 1) * out.numel()])

    def _finish_gradients(self):
        """Collect stragglers (p.grad set without the hook, parameters that got no gradient) and launch
        the reduces that are still outstanding."""
        for i, (_, p) in enumerate(self.named):
            if p.grad is not None:
                self._collect(i, p.grad)
                p.grad = None
        missing = set()
        for b in self.buckets:
            if not b.launched:
                space = self._touch(b)
                for i in b.params:
                    if not self.written[i]:   # no gradient this step (e.g. linear_2/3 of the depth heads)
                        missing.add(i)
                        if i not in self.expected_missing:   # (predicted ones were zeroed at first touch)
                            o = self.offsets[i] - b.lo
                            space[o:o + self.numels[i]].zero_()
                b.pending = 0
                self._reduce_bucket(b)
            else:
                missing.update(i for i in b.params if not self.written[i])
        if not self.accumulating:
            self.expected_missing = missing
        if self.comm is not None:
            cur = torch.cuda.current_stream()
            for b in self.buckets:
                if b.ev_reduced is not None:
                    cur.wait_event(b.ev_reduced)

    def step(self, lr_mult=1.0):
        self.step_count += 1
        timing = self.cuda and self.world > 1
        if timing:
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record()
        self._finish_gradients()
        if timing:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            self._events["reduce_wait"] = (e0, e1)
        g = self.g_shard
        ops.grad_sumsq(g, out=self.sumsq)
        if self.world > 1:
            dist.all_reduce(self.sumsq, group=self.pg)
        # gradients are SUMS over ranks here → average with 1/world inside the clip coefficient
        coef, norm = ops.clip_coef(self.sumsq, self.max_grad_norm or 0.0, 1.0 / self.world)
        self.last_grad_norm = norm
        for (slo, shi, plo, gi) in self.update_runs:
            _, lr_scale, wd = self.groups[gi]
            sl = slice(slo, shi)
            ops.adamw_step_(self.master[sl], self.m[sl], self.v[sl], g[sl], self.flat_p[plo:plo + (shi - slo)],
                            self.lr * lr_scale * lr_mult, self.betas[0], self.betas[1], self.eps, wd,
                            self.step_count, grad_scale=coef)
        if self.world > 1:
            self._gather_params()
        self._reset_step_state()

    def _gather_params(self):
        if self.comm is None:                 # gloo tests
            for b in self.buckets:
                full = self.flat_p[b.lo:b.hi]
                mine = full[self.rank * b.slice:(self.rank + 1) * b.slice].clone()
                parts = [torch.empty_like(mine) for _ in range(self.world)]
                dist.all_gather(parts, mine, group=self.pg)
                full.copy_(torch.cat(parts))
            return
        upd = torch.cuda.Event()
        upd.record()
        self.comm.wait_event(upd)
        evs = []
        with torch.cuda.stream(self.comm):
            for b in self.gather_order:       # order of first use in the forward pass
                full = self.flat_p[b.lo:b.hi]
                dist.all_gather_into_tensor(full, full[self.rank * b.slice:(self.rank + 1) * b.slice], group=self.pg)
                ev = torch.cuda.Event()
                ev.record()
                evs.append(ev)
        self._pending_gather = evs
        self._gather_waited = 0
        self._events.pop("gather_wait", None)
        self._gather_wait_pairs = []

    def wait_params(self, upto=None):
        """The compute stream waits for the updated parameters: all of them (upto=None), or only the buckets that
        hold parameters whose use_order rank is <= upto.  No-op when nothing (more) is pending."""
        if not self._pending_gather:
            return
        n = len(self._pending_gather)
        if upto is not None:
            ranks = [r for r in self.use_rank if r <= upto]
            n = (self._rank_last_pos[ranks[-1]] + 1) if ranks else 0
        if n <= self._gather_waited:
            return
        cur = torch.cuda.current_stream()
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
        cur.wait_event(self._pending_gather[n - 1])   # issued in order on one stream: the last one covers the rest
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        self._gather_wait_pairs.append((e0, e1))
        self._gather_waited = n
        if n == len(self._pending_gather):
            self._pending_gather = []

    def comm_summary(self):
        """Exposed communication of the LAST step: how long the compute stream sat in the two waits
        (after backward for the outstanding reduce-scatters; before the first trainable weight of the
        next forward for the parameter all-gathers).  Synchronises."""
        if not (self.cuda and self.world > 1):
            return None
        torch.cuda.synchronize()
        out = {"buckets": len(self.buckets), "bucket_mb": round(max(b.hi - b.lo for b in self.buckets) * 2 / 2**20, 1),
               "grad_bytes_per_step": self.total * 2, "overlap": self.overlap,
               "staging_peak_mb": round(self.staging_peak / 2**20, 1)}
        for k, (a, b) in self._events.items():
            out[f"exposed_{k}_ms"] = round(a.elapsed_time(b), 3)
        out["exposed_gather_wait_ms"] = round(sum(a.elapsed_time(b) for a, b in getattr(self, "_gather_wait_pairs", [])), 3)
        # a rank also waits in a collective for the OTHER ranks to arrive (clock / power-cap skew between GPUs):
        # the minimum over ranks is what the communication itself costs, the maximum includes the skew
        t = torch.tensor([out.get("exposed_reduce_wait_ms", 0.0), out["exposed_gather_wait_ms"]], device=self.dev)
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.pg)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=self.pg)
        out["exposed_reduce_wait_ms_min_max_over_ranks"] = [round(lo[0].item(), 3), round(hi[0].item(), 3)]
        out["exposed_gather_wait_ms_min_max_over_ranks"] = [round(lo[1].item(), 3), round(hi[1].item(), 3)]
        return out

    def state_dict(self):
        return {"step": self.step_count, "master": self.master, "m": self.m, "v": self.v,
                "rank": self.rank, "world": self.world}


def forward_use_rank(name):
    """Rank of a parameter's first use in the forward pass (ola_llama.py:79-188): projector → embedding / task
    tokens (splice) → decoder layer i → final norm, lm_head, heads.  The model passes the same ranks to
    `_pre_trainable_hook`, so layer i waits only for the all-gathers up to its own weights."""
    import re

    if "mm_projector" in name:
        return 0
    if "embed_tokens" in name or "special_" in name:
        return 1
    m = re.search(r"model\.layers\.(\d+)\.", name)
    if m:
        return 2 + int(m.group(1))
    return 1000


def _no_decay(name):
    return name.endswith(".bias") or "norm" in name or "layernorm" in name


class LLaVATrainer:
    """Constructor and entry points of ola_vlm/train/llava_trainer.py:217 (HF Trainer subclass)."""

    def __init__(self, model=None, tokenizer=None, args: TrainingArguments = None, train_dataset=None,
                 eval_dataset=None, data_collator=None, distributed=True, **kwargs):
        self.model = model
        self.tokenizer = tokenizer
        self.args = args or TrainingArguments()
        self.train_dataset = train_dataset
        self.eval_dataset = eval_dataset
        self.data_collator = data_collator
        self.deepspeed = None
        self.optimizer = None
        self.state = {"global_step": 0, "log_history": []}
        self.is_dist = bool(distributed) and dist.is_available() and dist.is_initialized()
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
        from ..model.modules import DecoderLayer, FusedRows

        fused = [fr.params for m in self.model.modules() for fr in vars(m).values() if isinstance(fr, FusedRows)]
        self.optimizer = Zero2Optimizer(self.model.named_parameters(), a.learning_rate,
                                        (a.adam_beta1, a.adam_beta2), a.adam_epsilon, a.weight_decay,
                                        a.max_grad_norm, groups, distributed=self.is_dist, keep_together=fused,
                                        bucket_elems=int(getattr(a, "zero_bucket_elems", 64 * 1024 * 1024)),
                                        use_order=forward_use_rank)
        for m in self.model.modules():   # decoder weight gradients are written straight into the optimizer's buffer
            if isinstance(m, DecoderLayer):
                m._grad_sink = self.optimizer.layer_sink(m.sink_params())
        if getattr(self.model, "supports_param_sync", False):
            # the parameter all-gather of step k overlaps the frozen tower of step k+1: the model calls this
            # where it first touches a trainable weight
            self.model._pre_trainable_hook = self.optimizer.wait_params
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
        if self.optimizer is not None and not getattr(model, "supports_param_sync", False):
            self.optimizer.wait_params()
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
        opt.set_accumulating(True)
        opt.wait_params()
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
        if self.train_dataset is not None and self.steps_per_epoch() == 0:
            raise ValueError(f"dataset of {len(self.train_dataset)} samples is smaller than one global batch "
                             f"({a.per_device_train_batch_size} x {self.world} ranks x {a.gradient_accumulation_steps} accumulation)")
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

            if self.optimizer is not None:
                self.optimizer.wait_params()
            sd = state_dict if state_dict is not None else {k: v.detach().cpu() for k, v in self.model.state_dict().items()}
            ckpt.save_config(self.model.config, output_dir)
            # HF save_pretrained layout (safetensors, 5 GB shards + index) so builder.py / from_pretrained load it
            ckpt.save_pretrained_weights(sd, output_dir, getattr(self.args, "max_shard_size", "5GB"))
