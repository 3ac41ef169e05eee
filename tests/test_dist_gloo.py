"""world_size-2 gloo tests (CPU) for the multi-GPU host logic: ZeRO-2 partition math of
Zero2Optimizer and the cross-rank InfoNCE target gather.  The CUDA kernels cannot run here, so the
four optimizer kernels are replaced by torch TEST DOUBLES inside this file only (the product has no
CPU fallback); what is under test is the sharding / collective plumbing around them."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from parity_utils import restate


def _install_test_doubles():
    from visper_lm_b200 import ops

    def axpby(a, b=None, alpha=1.0, beta=1.0, out=None):
        r = alpha * a.float() + (beta * b.float() if b is not None else 0)
        out.copy_(r.to(out.dtype))
        return out

    def grad_sumsq(grad, out=None, accumulate=False):
        s = grad.float().pow(2).sum().reshape(1)
        out.copy_(out + s if accumulate else s)
        return out

    def clip_coef(sumsq, max_norm, extra_scale=1.0):
        norm = sumsq.sqrt() * extra_scale
        c = torch.ones(1) if max_norm <= 0 else torch.clamp(max_norm / (norm + 1e-6), max=1.0)
        return c * extra_scale, norm

    def adamw_step_(master, m, v, grad, param, lr, b1, b2, eps, wd, step, grad_scale=None):
        g = grad.float() * (grad_scale if grad_scale is not None else 1.0)
        master.mul_(1 - lr * wd)
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = v.sqrt() / (1 - b2 ** step) ** 0.5 + eps
        master.addcdiv_(m, denom, value=-lr / (1 - b1 ** step))
        param.copy_(master.to(param.dtype))

    ops.axpby, ops.grad_sumsq, ops.clip_coef, ops.adamw_step_ = axpby, grad_sumsq, clip_coef, adamw_step_


class _Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(7)
        mk = lambda *s: torch.nn.Parameter(torch.randn(*s, generator=g).to(torch.bfloat16))
        self.proj = torch.nn.ParameterDict({"weight": mk(37, 24), "bias": mk(37)})
        self.norm = torch.nn.ParameterDict({"weight": mk(24)})
        self.logit_scale = torch.nn.Parameter(torch.tensor(2.0))
        self.frozen = torch.nn.Parameter(torch.randn(5, generator=g), requires_grad=False)


def _grads(step, rank, shape):
    g = torch.Generator().manual_seed(1000 * step + 10 * rank + len(shape))
    return (torch.randn(tuple(shape), generator=g) * 0.3).to(torch.bfloat16)


def _zero2_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _install_test_doubles()
    from visper_lm_b200.train.trainer import Zero2Optimizer, _no_decay

    model = _Toy()
    groups = [(lambda n: not _no_decay(n), 1.0, 0.01), (lambda n: True, 1.0, 0.0)]
    opt = Zero2Optimizer(model.named_parameters(), 1e-2, (0.9, 0.999), 1e-8, 0.01, 1.0, groups)
    # every bucket splits evenly (and 16-byte aligned) over the ranks; the shard is the sum of the rank's slices
    assert opt.shard * world == opt.total and all((b.hi - b.lo) % (world * 8) == 0 for b in opt.buckets)
    assert len({b.group for b in opt.buckets}) == 2   # decay / no-decay groups never share a bucket
    for step in range(3):
        opt.zero_grad()
        for n, p in opt.named:
            p.grad = _grads(step, rank, tuple(p.shape))
        opt.step()
    out = {n: p.detach().float().clone() for n, p in model.named_parameters()}
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        ret["params"] = gathered
        ret["norm"] = float(opt.last_grad_norm)
    dist.destroy_process_group()


def test_zero2_matches_single_process_adamw():
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_zero2_worker, args=(world, 29541, ret), nprocs=world, join=True)
    p0, p1 = ret["params"]
    for n in p0:
        assert torch.equal(p0[n], p1[n]), f"rank divergence in {n}"
    # single-process reference: fp32 AdamW on the rank-averaged gradient, global-norm clip 1.0
    ref = _Toy()
    params = {n: torch.nn.Parameter(p.detach().float()) for n, p in ref.named_parameters() if p.requires_grad}
    decay = [p for n, p in params.items() if not (n.endswith(".bias") or "norm" in n)]
    nodecay = [p for n, p in params.items() if (n.endswith(".bias") or "norm" in n)]
    opt = torch.optim.AdamW([{"params": decay, "weight_decay": 0.01}, {"params": nodecay, "weight_decay": 0.0}],
                            lr=1e-2, betas=(0.9, 0.999), eps=1e-8)
    for step in range(3):
        for n, p in params.items():
            # the collective sums bf16 gradients (as NCCL does), then averages
            s = sum(_grads(step, r, tuple(p.shape)).float() for r in range(world)).to(torch.bfloat16).float()
            p.grad = s / world
        torch.nn.utils.clip_grad_norm_(list(params.values()), 1.0)
        opt.step()
    for n, p in params.items():
        assert torch.allclose(p0[n], p.detach(), rtol=2 ** -7, atol=1e-3), n
    assert torch.equal(p0["frozen"], ref.frozen.float())


def _infonce_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from visper_lm_b200.model.vlm import gather_targets

    B, n = 3, 64
    g = torch.Generator().manual_seed(5)
    preds = torch.randn(world * B, n, generator=g)
    tgts = torch.randn(world * B, n, generator=g)
    mine_p = preds[rank * B:(rank + 1) * B].clone().requires_grad_(True)
    tall, off, ev = gather_targets(tgts[rank * B:(rank + 1) * B].clone())
    assert ev is None   # CPU tensors: synchronous collective
    assert off == rank * B and torch.equal(tall, tgts)
    tau = torch.tensor(2.0)
    ce = restate.calculate_contrastive_loss(mine_p, tall, tau, rank=rank)
    ce.mean().backward()
    res = [None] * world
    dist.all_gather_object(res, (ce.detach(), mine_p.grad))
    if rank == 0:
        ret["res"] = res
    dist.destroy_process_group()


def test_cross_rank_infonce_equals_global_batch():
    world, B = 2, 3
    ret = mp.Manager().dict()
    mp.spawn(_infonce_worker, args=(world, 29543, ret), nprocs=world, join=True)
    g = torch.Generator().manual_seed(5)
    preds = torch.randn(world * B, 64, generator=g).requires_grad_(True)
    tgts = torch.randn(world * B, 64, generator=g)
    ce = restate.calculate_contrastive_loss(preds, tgts, torch.tensor(2.0), rank=0)
    # global objective = mean over the GLOBAL batch; DP averages rank-local means → same gradient
    ce.mean().backward()
    for r, (ce_r, grad_r) in enumerate(ret["res"]):
        assert torch.allclose(ce_r, ce[r * B:(r + 1) * B].detach(), atol=1e-6)
        assert torch.allclose(grad_r / world, preds.grad[r * B:(r + 1) * B], atol=1e-6)


def test_splice_plan_matches_oracle_splice():
    """Host-side plan (product) vs the oracle's restatement of ola_arch.py:337-444 on ragged input."""
    from visper_lm_b200.model.vlm import SplicePlan

    g = torch.Generator().manual_seed(3)
    B, N, D, nimg, ntask = 4, 30, 16, 576, 24
    ids = torch.randint(0, 50, (B, N), generator=g)
    ids[0, 5] = -200
    ids[1, 0] = -200
    ids[2, N - 1] = -200          # image token last
    am = torch.ones(B, N, dtype=torch.bool)   # row 3: no image token at all
    am[1, 20:] = False
    labels = torch.where(am, ids, torch.full_like(ids, -100))
    plan = SplicePlan(ids, labels, am, nimg, ntask, 620, "right")
    embed = torch.randn(50, D, generator=g)
    feats = torch.randn(B * nimg, D, generator=g)
    sd = {"model.embed_tokens.weight": embed,
          "model.special_gen_tokens": torch.randn(8, D, generator=g),
          "model.special_depth_tokens": torch.randn(576, D, generator=g),
          "model.special_seg_tokens": torch.randn(576, D, generator=g)}
    cfg = dict(num_task_tokens=8, aux_mode="gen-depth-seg", tokenizer_model_max_length=620)
    emb, lab, mask = restate.splice(sd, cfg, ids, labels, am, feats.view(B, nimg, D))
    task_rows = torch.cat(restate.pooled_task_tokens(sd, cfg), 0)
    srcs = [embed, feats, task_rows]
    kind, index = plan.np["kind"], plan.np["index"]
    out = torch.zeros(B * plan.T, D)
    for r in range(B * plan.T):
        if kind[r] >= 0:
            out[r] = srcs[kind[r]][index[r]]
    assert plan.T == emb.shape[1]
    assert torch.equal(out.view(B, plan.T, D), emb)
    assert torch.equal(torch.from_numpy(plan.np["labels"]), lab)
    assert torch.equal(torch.from_numpy(plan.np["mask"]), mask)
    # inverse maps are consistent
    inv = plan.np["inv_img"]
    for j in (0, 575, 576, 4 * 576 - 1):
        if inv[j] >= 0:
            assert kind[inv[j]] == 1 and index[inv[j]] == j
