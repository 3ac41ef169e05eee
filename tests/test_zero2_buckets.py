"""Host logic of the bucketed ZeRO-2 optimizer on CPU (torch TEST DOUBLES stand in for the four optimizer
kernels, as in test_dist_gloo.py): bucket layout, the autograd-hook gradient path, parameters that get no
gradient, gradient accumulation, and single ownership of fused q|k|v storage (FusedRows used to
re-copy the weights out of the optimizer's flat buffer, so AdamW updated memory no GEMM read)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from test_dist_gloo import _install_test_doubles


class _Layer(torch.nn.Module):
    def __init__(self, seed):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        mk = lambda *s: torch.nn.Parameter((0.2 * torch.randn(*s, generator=g)).to(torch.bfloat16))
        self.q, self.k, self.v = mk(16, 8), mk(8, 8), mk(8, 8)
        self.o = mk(8, 16)
        self.unused = mk(24)          # never reached by the loss: its gradient must read as zero
        self.bias = mk(8)

    def forward(self, x):
        w = torch.cat([self.q, self.k, self.v], 0).float()
        h = x @ w.t()
        return (h[:, :16] @ self.o.float().t()) + h[:, 16:24] * h[:, 24:] + self.bias.float()


class _Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.layers = torch.nn.ModuleList([_Layer(s) for s in (1, 2, 3)])
        self.scale = torch.nn.Parameter(torch.tensor(2.0))

    def forward(self, x):
        for l in self.layers:
            x = x + l(x)
        return (x.pow(2).mean()) * self.scale.float()


def _batch(step, rank):
    g = torch.Generator().manual_seed(100 * step + rank)
    return torch.randn(4, 8, generator=g)


def _make(world_groups=False, **kw):
    from visper_lm_b200.train.trainer import Zero2Optimizer, _no_decay

    net = _Net()
    groups = [(lambda n: not _no_decay(n), 1.0, 0.01), (lambda n: True, 1.0, 0.0)]
    together = [[l.q, l.k, l.v] for l in net.layers]
    opt = Zero2Optimizer(net.named_parameters(), 1e-2, (0.9, 0.999), 1e-8, 0.01, 1.0, groups, keep_together=together,
                         bucket_elems=200, **kw)
    return net, opt


def test_layout_buckets_and_adjacency():
    _install_test_doubles()
    net, opt = _make()
    assert len(opt.buckets) >= 3
    for b in opt.buckets:
        assert (b.hi - b.lo) % (opt.world * 8) == 0
    for l in net.layers:      # fused groups adjacent, in one bucket, and still views of the flat buffer
        iq = opt.index[id(l.q)]
        assert opt.offsets[iq + 1] == opt.offsets[iq] + l.q.numel()
        assert opt.bucket_of[iq] is opt.bucket_of[iq + 2]
        assert l.k.data_ptr() == l.q.data_ptr() + l.q.numel() * 2
    lo, hi = opt.flat_p.data_ptr(), opt.flat_p.data_ptr() + opt.flat_p.numel() * 2
    assert all(lo <= p.data_ptr() < hi for _, p in opt.named)


def test_fused_rows_adopts_optimizer_storage_and_updates_reach_it():
    _install_test_doubles()
    from visper_lm_b200.model.modules import FusedRows

    net, opt = _make()
    l = net.layers[1]
    fr = FusedRows([l.q, l.k, l.v])
    fused = fr.get()
    assert fused.shape == (32, 8) and fused.data_ptr() == l.q.data_ptr()          # adopted, not copied
    lo, hi = opt.flat_p.data_ptr(), opt.flat_p.data_ptr() + opt.flat_p.numel() * 2
    before = fused.clone()
    net(_batch(0, 0)).backward()
    opt.step()
    assert fr.get() is fused and lo <= l.q.data_ptr() < hi and lo <= l.v.data_ptr() < hi
    assert not torch.equal(fused, before), "AdamW's update did not reach the fused weight the GEMM reads"
    assert torch.equal(fused[:16], l.q.data) and torch.equal(fused[24:], l.v.data)
    # without an optimizer the object owns a concatenated copy, as before
    m = _Layer(9)
    f2 = FusedRows([m.q, m.k, m.v]).get()
    assert m.q.data_ptr() == f2.data_ptr() and m.v.data_ptr() == f2.data_ptr() + 24 * 8 * 2


def _reference_run(world, steps, accumulate=1):
    ref = _Net()
    params = {n: torch.nn.Parameter(p.detach().float()) for n, p in ref.named_parameters()}
    nd = lambda n: n.endswith(".bias") or "norm" in n
    opt = torch.optim.AdamW([{"params": [p for n, p in params.items() if not nd(n)], "weight_decay": 0.01},
                             {"params": [p for n, p in params.items() if nd(n)], "weight_decay": 0.0}],
                            lr=1e-2, betas=(0.9, 0.999), eps=1e-8)
    for step in range(steps):
        grads = {n: torch.zeros_like(p) for n, p in params.items()}
        for r in range(world):
            acc = {n: torch.zeros(p.shape, dtype=torch.bfloat16) for n, p in params.items()}
            for k in range(accumulate):
                net = _Net()
                with torch.no_grad():
                    for n, p in net.named_parameters():
                        p.copy_(params[n].to(torch.bfloat16))
                (net(_batch(step * accumulate + k, r)) / accumulate).backward()
                for n, p in net.named_parameters():
                    if p.grad is not None:   # bf16 accumulation across micro-batches, as p.grad / the sink does
                        acc[n] = (acc[n].float() + p.grad.float()).to(torch.bfloat16)
            for n in grads:
                grads[n] += acc[n].float()
        for n, p in params.items():
            p.grad = grads[n].to(torch.bfloat16).float() / world
        torch.nn.utils.clip_grad_norm_(list(params.values()), 1.0)
        opt.step()
    return {n: p.detach() for n, p in params.items()}


def _worker(rank, world, port, steps, accumulate, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _install_test_doubles()
    net, opt = _make()
    for step in range(steps):
        opt.zero_grad()
        if accumulate > 1:
            opt.set_accumulating(True)
        for k in range(accumulate):
            (net(_batch(step * accumulate + k, rank)) / accumulate).backward()   # hooks move p.grad into the buckets
        assert all(p.grad is None for _, p in opt.named)
        opt.step()
    out = {n: p.detach().float().clone() for n, p in net.named_parameters()}
    res = [None] * world
    dist.all_gather_object(res, out)
    if rank == 0:
        ret["params"] = res
    dist.destroy_process_group()


def _check(world, steps, accumulate, port):
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, steps, accumulate, ret), nprocs=world, join=True)
    ps = ret["params"]
    for n in ps[0]:
        for other in ps[1:]:
            assert torch.equal(ps[0][n], other[n]), f"rank divergence in {n}"
    ref = _reference_run(world, steps, accumulate)
    for n, p in ref.items():
        assert torch.allclose(ps[0][n], p, rtol=2 ** -6, atol=2e-3), n
    unused = ps[0]["layers.0.unused"]
    assert torch.allclose(unused, _Net().layers[0].unused.float() * (1 - 1e-2 * 0.01) ** steps, rtol=2 ** -7), \
        "a parameter without gradient must only see weight decay"


def test_bucketed_hooks_world2_matches_single_process():
    _check(2, 3, 1, 29551)


def test_gradient_accumulation_world2():
    _check(2, 2, 2, 29553)


def test_single_process_hooks_and_stale_gradients():
    """world 1: the gradient shard IS the gradient space; a parameter that gets no gradient in a later step
    must not keep the previous step's values."""
    _install_test_doubles()
    net, opt = _make()
    opt.zero_grad()
    net(_batch(0, 0)).backward()
    i = opt.index[id(net.layers[0].o)]
    opt._finish_gradients()
    o = opt.offsets[i]
    assert opt.g_shard[o:o + 8].float().abs().sum() > 0
    opt._reset_step_state()
    opt.zero_grad()
    opt._finish_gradients()               # no backward at all this step
    assert opt.g_shard.float().abs().sum() == 0
