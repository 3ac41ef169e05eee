"""GPU unit tests: every C-ABI kernel against a plain torch fp32 restatement of the same op.

Tolerances are bf16-output tolerances: |err| <= atol + rtol*|ref| with rtol ~ 2^-7 unless noted.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

BF = torch.bfloat16


def dev():
    return torch.device("cuda:0")


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(BF).to(dev())


def close(got, ref, rtol=1.6e-2, atol=None, name=""):
    got = got.float()
    ref = ref.float()
    if atol is None:
        atol = 1e-2 * max(ref.abs().max().item(), 1e-6)
    err = (got - ref).abs()
    bad = err > (atol + rtol * ref.abs())
    assert not torch.isnan(got).any(), f"{name}: NaN in output"
    assert not bad.any(), (
        f"{name}: {int(bad.sum())}/{bad.numel()} mismatches, max err {err.max().item():.4g}, "
        f"ref max {ref.abs().max().item():.4g}")


# ------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("al,bl", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 512, 256), (304, 520, 200), (1000, 264, 72),
                                   (2048, 6144, 1024), (4096, 256, 4096)])
def test_gemm_layouts(M, N, K, al, bl):
    from visper_lm_b200 import ops
    a = rnd(M, K, seed=1)
    b = rnd(N, K, seed=2)
    ref = a.float() @ b.float().t()
    a_in = a if al == 0 else a.t().contiguous()
    b_in = b if bl == 0 else b.t().contiguous()
    if (al == 1 and M % 8) or (bl == 1 and N % 8) or (al == 0 and K % 8) or (bl == 0 and K % 8):
        pytest.skip("stride not 16-byte aligned for this layout")
    out = ops.gemm(a_in, b_in, a_layout=al, b_layout=bl)
    torch.cuda.synchronize()
    close(out, ref, name=f"gemm {M}x{N}x{K} a{al} b{bl}")


@pytest.mark.parametrize("al,bl", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(4096, 4096, 512), (2304, 5000, 328), (2000, 8192, 1024), (16384, 6144, 256)])
def test_gemm_cta_pair_matches_single_cta(M, N, K, al, bl):
    """cta_group::2 kernel (two CTAs per 256x256 tile) against the 1-CTA kernel: same K order, so
    bit-identical; both against torch.  Shapes include ragged M / N edges."""
    from visper_lm_b200 import ops
    if (al == 1 and M % 8) or (bl == 1 and N % 8):
        pytest.skip("stride not 16-byte aligned for this layout")
    a = rnd(M, K, seed=41)
    b = rnd(N, K, seed=42)
    bias, res = rnd(N, seed=43), rnd(M, N, seed=44)
    a_in = a if al == 0 else a.t().contiguous()
    b_in = b if bl == 0 else b.t().contiguous()
    ops.set_option(ops.OPT_GEMM_1CTA, 1)
    try:
        ref1 = ops.gemm(a_in, b_in, a_layout=al, b_layout=bl, bias=bias, act=1, residual=res)
        torch.cuda.synchronize()
    finally:
        ops.set_option(ops.OPT_GEMM_1CTA, 0)
    out = ops.gemm(a_in, b_in, a_layout=al, b_layout=bl, bias=bias, act=1, residual=res)
    torch.cuda.synchronize()
    assert torch.equal(out, ref1), f"pair kernel differs from 1-CTA kernel: {(out.float() - ref1.float()).abs().max().item()}"
    ref = torch.nn.functional.gelu(a.float() @ b.float().t() + bias.float()) + res.float()
    close(out, ref, name=f"pair gemm {M}x{N}x{K} a{al} b{bl}")


@pytest.mark.parametrize("act", [0, 1, 2, 3])
def test_gemm_epilogues(act):
    from visper_lm_b200 import ops
    M, N, K = 520, 776, 320
    a, b = rnd(M, K, seed=3, scale=0.5), rnd(N, K, seed=4, scale=0.2)
    bias, res = rnd(N, seed=5), rnd(M, N, seed=6)
    pre_ref = a.float() @ b.float().t() + bias.float()
    if act == 1:
        y = torch.nn.functional.gelu(pre_ref)
    elif act == 2:
        y = pre_ref * torch.sigmoid(1.702 * pre_ref)
    elif act == 3:
        y = torch.relu(pre_ref)
    else:
        y = pre_ref
    ref = y + res.float()
    out, pre = ops.gemm(a, b, bias=bias, act=act, residual=res, want_pre=True)
    torch.cuda.synchronize()
    close(pre, pre_ref, name="pre-activation")
    close(out, ref, name=f"epilogue act={act}")
    # in-place accumulate: residual aliases the output
    acc = res.clone()
    ops.gemm(a, b, residual=acc, out=acc)
    torch.cuda.synchronize()
    close(acc, a.float() @ b.float().t() + res.float(), name="accumulate")


def test_gemm_strided_views():
    from visper_lm_b200 import ops
    big = rnd(300, 1024, seed=7)
    a = big[:, 256:512]            # lda = 1024
    w = rnd(384, 256, seed=8)
    outbuf = torch.zeros(300, 1024, dtype=BF, device=dev())
    ops.gemm(a, w, out=outbuf[:, 128:512])
    torch.cuda.synchronize()
    close(outbuf[:, 128:512], a.float() @ w.float().t(), name="strided")
    assert outbuf[:, :128].abs().max().item() == 0 and outbuf[:, 512:].abs().max().item() == 0


@pytest.mark.parametrize("M,F,K", [(200, 128, 64), (1000, 384, 200), (2048, 1792, 512), (4096, 14336, 256),
                                   (2100, 3584, 320)])
def test_gemm_swiglu_fused(M, F, K):
    """SwiGLU in the GEMM epilogues: bit-identical to the separate GEMM + swiglu kernels."""
    from visper_lm_b200 import ops
    x = rnd(M, K, seed=31, scale=0.5)
    wgu = rnd(2 * F, K, seed=32, scale=0.2)
    wd = rnd(K, F, seed=33, scale=0.2)       # down_proj [D, F]
    dy = rnd(M, K, seed=34)
    gu_ref = ops.gemm(x, wgu)
    h_ref = ops.swiglu_fwd(gu_ref)
    h, gu = ops.gemm_swiglu_fwd(x, wgu)
    torch.cuda.synchronize()
    assert torch.equal(gu, gu_ref), "fused gate|up differs from the plain GEMM"
    assert torch.equal(h, h_ref), "fused silu(g)*u differs from swiglu_fwd"
    g, u = gu.float()[:, :F], gu.float()[:, F:]
    close(h, torch.nn.functional.silu(g) * u, name="swiglu fwd vs torch")
    h2, none = ops.gemm_swiglu_fwd(x, wgu, want_gu=False)
    assert none is None and torch.equal(h2, h_ref)
    dgu_ref = ops.swiglu_bwd(gu_ref, ops.gemm(dy, wd, b_layout=1))
    dgu = ops.gemm_swiglu_bwd(dy, wd, gu, b_layout=1)
    dgu_t = ops.gemm_swiglu_bwd(dy, wd.t().contiguous(), gu, b_layout=0)
    torch.cuda.synchronize()
    assert torch.equal(dgu, dgu_ref), "fused swiglu backward (MN-major W) differs"
    assert torch.equal(dgu_t, dgu_ref), "fused swiglu backward (K-major Wt) differs"
    # tile-major g|u (what the frozen-LLM step saves): same h, same gradients
    h3, gut = ops.gemm_swiglu_fwd(x, wgu, tiled=True)
    dgu3 = ops.gemm_swiglu_bwd(dy, wd, gut, b_layout=1, tiled=True, F=F)
    torch.cuda.synchronize()
    assert torch.equal(h3, h_ref) and torch.equal(dgu3, dgu_ref), "tile-major g|u path differs"


@pytest.mark.parametrize("M,H,KVH,K,T", [(300, 2, 2, 128, 100), (2048, 32, 8, 512, 512), (1000, 4, 2, 264, 1000)])
def test_gemm_rope_fused(M, H, KVH, K, T):
    """RoPE in the QKV GEMM epilogue: bit-identical to GEMM + in-place rope kernel."""
    from visper_lm_b200 import ops
    hd = 128
    N = (H + 2 * KVH) * hd
    if N % 256:
        pytest.skip("needs N % 256 == 0")
    x = rnd(M, K, seed=35, scale=0.5)
    w = rnd(N, K, seed=36, scale=0.2)
    cos, sin = ops.rope_tables(max(T, 64), hd, 500000.0, dev())
    ref = ops.gemm(x, w)
    ops.rope_(ref, T, cos, sin, H + KVH, hd)
    out = ops.gemm_rope(x, w, T, cos, sin, H + KVH)
    torch.cuda.synchronize()
    assert torch.equal(out, ref), f"fused rope differs: {(out.float() - ref.float()).abs().max().item()}"
    pos = torch.randint(0, max(T, 64), (M,), device=dev(), dtype=torch.int32)
    ref2 = ops.gemm(x, w)
    ops.rope_(ref2, T, cos, sin, H + KVH, hd, pos_ids=pos)
    out2 = ops.gemm_rope(x, w, T, cos, sin, H + KVH, pos_ids=pos)
    torch.cuda.synchronize()
    assert torch.equal(out2, ref2), "fused rope with explicit position ids differs"


# ------------------------------------------------------------------------------------------- norms
@pytest.mark.parametrize("M,D", [(37, 128), (512, 4096), (100, 3072), (64, 1024), (1003, 2048), (16384, 4096), (5001, 3072),
                                 (2048, 1024), (1500, 4096)])
def test_rmsnorm(M, D):
    from visper_lm_b200 import ops
    x = rnd(M, D, seed=11).requires_grad_(False)
    w = (1 + 0.1 * rnd(D, seed=12).float()).to(BF)
    dy = rnd(M, D, seed=13)
    dres = rnd(M, D, seed=14)
    y, rstd = ops.rmsnorm_fwd(x, w, 1e-5)
    xf = x.float().requires_grad_(True)
    wf = w.float().requires_grad_(True)
    ref = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5) * wf
    ref.backward(dy.float())
    close(y, ref, name="rmsnorm fwd")
    dx = ops.rmsnorm_bwd(dy, x, w, rstd, dres)
    close(dx, xf.grad + dres.float(), name="rmsnorm bwd")
    dw = ops.colsum(dy, x, None, rstd, out_dtype=torch.float32)
    close(dw, wf.grad, rtol=2e-2, name="rmsnorm dw")


def test_rmsnorm_strided_row_views():
    """RMSNorm forward / backward on row views of a wider buffer (ld > D), output into a strided view."""
    from visper_lm_b200 import ops
    M, D = 777, 4096
    big = rnd(M, 3 * D, seed=31)
    x, dy, dres = big[:, :D], big[:, D:2 * D], big[:, 2 * D:]
    w = (1 + 0.1 * rnd(D, seed=32).float()).to(BF)
    out = torch.zeros(M, 2 * D, dtype=BF, device=dev())
    y, rstd = ops.rmsnorm_fwd(x, w, 1e-5, out=out[:, D:])
    dx = ops.rmsnorm_bwd(dy, x, w, rstd, dres)
    y0, rstd0 = ops.rmsnorm_fwd(x.contiguous(), w, 1e-5)
    dx0 = ops.rmsnorm_bwd(dy.contiguous(), x.contiguous(), w, rstd0, dres.contiguous())
    torch.cuda.synchronize()
    assert torch.equal(out[:, :D], torch.zeros_like(out[:, :D])), "wrote outside its row view"
    assert torch.equal(y, y0) and torch.equal(rstd, rstd0) and torch.equal(dx, dx0)
    # M >= 1024 takes the cp.async-pipelined kernels: same values as the CTA-per-row kernels up to the order of the
    # fp32 row sums, also on strided views and without the residual-gradient input
    M = 3000
    big = rnd(M, 3 * D, seed=33)
    x, dy, dres = big[:, :D], big[:, D:2 * D], big[:, 2 * D:]
    y, rstd = ops.rmsnorm_fwd(x, w, 1e-5)
    dx = ops.rmsnorm_bwd(dy, x, w, rstd, dres)
    dxn = ops.rmsnorm_bwd(dy, x, w, rstd, None)
    ops.set_option(ops.OPT_NORM_R1, 1)
    try:
        y0, rstd0 = ops.rmsnorm_fwd(x, w, 1e-5)
        dx0 = ops.rmsnorm_bwd(dy, x, w, rstd0, dres)
        dxn0 = ops.rmsnorm_bwd(dy, x, w, rstd0, None)
        torch.cuda.synchronize()
    finally:
        ops.set_option(ops.OPT_NORM_R1, 0)
    assert torch.allclose(rstd, rstd0, rtol=1e-6)
    assert (y.float() - y0.float()).abs().max().item() <= 2 ** -6 * y0.float().abs().max().item()
    assert ((dx.float() - dx0.float()).norm() / dx0.float().norm()).item() < 2e-3
    assert ((dxn.float() - dxn0.float()).norm() / dxn0.float().norm()).item() < 2e-3


@pytest.mark.parametrize("M,D", [(50, 64), (300, 1024), (77, 1536)])
def test_layernorm(M, D):
    from visper_lm_b200 import ops
    x = rnd(M, D, seed=21)
    w = (1 + 0.1 * rnd(D, seed=22).float()).to(BF)
    b = rnd(D, seed=23, scale=0.1)
    dy = rnd(M, D, seed=24)
    y, mean, rstd = ops.layernorm_fwd(x, w, b, 1e-5)
    xf = x.float().requires_grad_(True)
    wf = w.float().requires_grad_(True)
    bf = b.float().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xf, (D,), wf, bf, 1e-5)
    ref.backward(dy.float())
    close(y, ref, name="ln fwd")
    dx = ops.layernorm_bwd(dy, x, w, mean, rstd)
    close(dx, xf.grad, name="ln bwd")
    close(ops.colsum(dy, x, mean, rstd, out_dtype=torch.float32), wf.grad, rtol=2e-2, name="ln dw")
    close(ops.colsum(dy, out_dtype=torch.float32), bf.grad, rtol=2e-2, name="ln db")


# ------------------------------------------------------------------------------------------- rope / swiglu / act
def _rope_ref(x, T, hd, theta):
    # x [M, nh*hd] fp32; HF rotate_half convention
    M = x.shape[0]
    nh = x.shape[1] // hd
    pos = (torch.arange(M, device=x.device) % T).float()
    inv = 1.0 / (theta ** (torch.arange(0, hd, 2, device=x.device).float() / hd))
    ang = pos[:, None] * inv[None, :]
    cos = torch.cat([ang.cos(), ang.cos()], -1)[:, None, :]
    sin = torch.cat([ang.sin(), ang.sin()], -1)[:, None, :]
    xv = x.view(M, nh, hd)
    rot = torch.cat([-xv[..., hd // 2:], xv[..., :hd // 2]], -1)
    return (xv * cos + rot * sin).reshape(M, nh * hd)


@pytest.mark.parametrize("hd,theta", [(128, 500000.0), (96, 10000.0), (32, 10000.0)])
def test_rope(hd, theta):
    from visper_lm_b200 import ops
    T, B, nh_q, nh_kv = 50, 3, 4, 2
    width = (nh_q + 2 * nh_kv) * hd
    qkv = rnd(B * T, width, seed=31)
    orig = qkv.clone()
    cos, sin = ops.rope_tables(64, hd, theta, dev())
    ops.rope_(qkv, T, cos, sin, nh_q + nh_kv, hd)
    ref = _rope_ref(orig[:, :(nh_q + nh_kv) * hd].float(), T, hd, theta)
    close(qkv[:, :(nh_q + nh_kv) * hd], ref, name="rope fwd")
    assert torch.equal(qkv[:, (nh_q + nh_kv) * hd:], orig[:, (nh_q + nh_kv) * hd:]), "V must be untouched"
    ops.rope_(qkv, T, cos, sin, nh_q + nh_kv, hd, inverse=True)
    close(qkv, orig.float(), name="rope inverse round trip")


def test_swiglu_and_act_bwd():
    from visper_lm_b200 import ops
    M, F = 130, 512
    gu = rnd(M, 2 * F, seed=41)
    dh = rnd(M, F, seed=42)
    guf = gu.float().requires_grad_(True)
    ref = torch.nn.functional.silu(guf[:, :F]) * guf[:, F:]
    ref.backward(dh.float())
    close(ops.swiglu_fwd(gu), ref, name="swiglu fwd")
    close(ops.swiglu_bwd(gu, dh), guf.grad, name="swiglu bwd")
    for act, fn in [(1, torch.nn.functional.gelu), (2, lambda t: t * torch.sigmoid(1.702 * t)),
                    (3, torch.relu)]:
        pre = rnd(M, F, seed=43)
        pf = pre.float().requires_grad_(True)
        fn(pf).backward(dh.float())
        close(ops.act_bwd(pre, dh, act), pf.grad, name=f"act_bwd {act}")


# ------------------------------------------------------------------------------------------- attention
def _attn_ref(q, k, v, scale, causal):
    # q [B,H,sq,hd], k/v [B,H,sk,hd] fp32
    s = (q @ k.transpose(-1, -2)) * scale
    if causal:
        sq, sk = s.shape[-2:]
        m = torch.ones(sq, sk, dtype=torch.bool, device=s.device).tril(sk - sq)
        s = s.masked_fill(~m, float("-inf"))
    return torch.softmax(s, -1) @ v


@pytest.mark.parametrize("B,H,KVH,sq,sk,sk2,hd,causal", [
    (2, 4, 2, 300, 300, 0, 128, True),
    (1, 2, 2, 577, 577, 0, 64, False),
    (2, 4, 4, 200, 200, 0, 96, True),
    (2, 4, 4, 70, 203, 70, 32, False),
    (2, 4, 4, 1, 150, 1, 32, False),
    (2, 4, 2, 130, 130, 0, 32, True),
    (1, 8, 2, 1024, 1024, 0, 128, True),
])
def test_attention_fwd_bwd(B, H, KVH, sq, sk, sk2, hd, causal):
    from visper_lm_b200 import ops
    qw, kw = H * hd, KVH * hd
    qb = rnd(B * sq, qw, seed=51)
    kvb = rnd(B * sk, 2 * kw, seed=52)
    kv2 = rnd(B * sk2, 2 * kw, seed=53) if sk2 else None
    do = rnd(B * sq, qw, seed=54)
    scale = hd ** -0.5
    k, v = kvb[:, :kw], kvb[:, kw:]
    k2 = kv2[:, :kw] if sk2 else None
    v2 = kv2[:, kw:] if sk2 else None
    o, lse = ops.attn_fwd(qb, k, v, B, H, KVH, sq, sk, hd, scale, causal, k2=k2, v2=v2, sk2=sk2)
    torch.cuda.synchronize()

    qf = qb.float().view(B, sq, H, hd).transpose(1, 2).detach().requires_grad_(True)
    kcat = kvb.float().view(B, sk, 2, KVH, hd)
    if sk2:
        kcat = torch.cat([kcat, kv2.float().view(B, sk2, 2, KVH, hd)], 1)
    kcat = kcat.detach().requires_grad_(True)
    kf = kcat[:, :, 0].transpose(1, 2).repeat_interleave(H // KVH, 1)
    vf = kcat[:, :, 1].transpose(1, 2).repeat_interleave(H // KVH, 1)
    ref = _attn_ref(qf, kf, vf, scale, causal)
    ref_rows = ref.transpose(1, 2).reshape(B * sq, qw)
    close(o, ref_rows, name="attn fwd")
    ref_rows.backward(do.float())

    dq = torch.empty_like(qb)
    dkv = torch.empty_like(kvb)
    dkv2 = torch.empty_like(kv2) if sk2 else None
    ops.attn_bwd(qb, k, v, o, do, lse, dq, dkv[:, :kw], dkv[:, kw:], B, H, KVH, sq, sk, hd, scale,
                 causal, k2=k2, v2=v2, sk2=sk2, dk2=dkv2[:, :kw] if sk2 else None,
                 dv2=dkv2[:, kw:] if sk2 else None)
    torch.cuda.synchronize()
    close(dq, qf.grad.transpose(1, 2).reshape(B * sq, qw), rtol=3e-2, name="attn dq")
    g = kcat.grad  # [B, sk+sk2, 2, KVH, hd]
    close(dkv, g[:, :sk].reshape(B * sk, 2 * kw), rtol=3e-2, name="attn dkv")
    if sk2:
        close(dkv2, g[:, sk:].reshape(B * sk2, 2 * kw), rtol=3e-2, name="attn dkv2")


@pytest.mark.parametrize("B,H,KVH,sq,sk,causal", [(2, 8, 8, 1024, 1024, True), (1, 4, 4, 333, 333, True),
                                                   (2, 4, 2, 200, 520, False), (1, 32, 32, 2048, 2048, True),
                                                   (1, 3, 3, 577, 577, False)])
def test_attention_tcgen05_head_dim_96(B, H, KVH, sq, sk, causal):
    """Phi-3 head_dim 96 on the tcgen05 kernels (64 + 32-column TMA chunks, N=96 accumulators):
    against the mma.sync kernels and torch, forward and backward."""
    from visper_lm_b200 import ops
    hd = 96
    qw, kw = H * hd, KVH * hd
    q = rnd(B * sq, qw, seed=91)
    kv = rnd(B * sk, 2 * kw, seed=92)
    do = rnd(B * sq, qw, seed=93)
    k, v = kv[:, :kw], kv[:, kw:]
    scale = hd ** -0.5

    def run():
        o, lse = ops.attn_fwd(q, k, v, B, H, KVH, sq, sk, hd, scale, causal)
        dq = torch.zeros_like(q)
        dkv = torch.zeros_like(kv)
        ops.attn_bwd(q, k, v, o, do, lse, dq, dkv[:, :kw], dkv[:, kw:], B, H, KVH, sq, sk, hd, scale, causal)
        torch.cuda.synchronize()
        return o, lse, dq, dkv

    ops.set_option(ops.OPT_ATTN_LEGACY_FWD, 1)
    ops.set_option(ops.OPT_ATTN_LEGACY_BWD, 1)
    try:
        o1, lse1, dq1, dkv1 = run()
    finally:
        ops.set_option(ops.OPT_ATTN_LEGACY_FWD, 0)
        ops.set_option(ops.OPT_ATTN_LEGACY_BWD, 0)
    o2, lse2, dq2, dkv2 = run()
    close(o2, o1, name="hd96 tc fwd vs mma.sync")
    assert (lse2 - lse1).abs().max().item() < 2e-2
    close(dq2, dq1, rtol=2e-2, name="hd96 dq vs mma.sync")
    close(dkv2, dkv1, rtol=2e-2, name="hd96 dkv vs mma.sync")
    qf = q.float().view(B, sq, H, hd).transpose(1, 2).detach().requires_grad_(True)
    kvf = kv.float().view(B, sk, 2, KVH, hd).detach().requires_grad_(True)
    kf = kvf[:, :, 0].transpose(1, 2).repeat_interleave(H // KVH, 1)
    vf = kvf[:, :, 1].transpose(1, 2).repeat_interleave(H // KVH, 1)
    ref = _attn_ref(qf, kf, vf, scale, causal).transpose(1, 2).reshape(B * sq, qw)
    close(o2, ref, name="hd96 fwd vs torch")
    ref.backward(do.float())
    close(dq2, qf.grad.transpose(1, 2).reshape(B * sq, qw), rtol=3e-2, name="hd96 dq vs torch")
    close(dkv2, kvf.grad.reshape(B * sk, 2 * kw), rtol=3e-2, name="hd96 dkv vs torch")


@pytest.mark.parametrize("B,H,KVH,S,hd,window", [(2, 4, 4, 700, 96, 200), (1, 4, 2, 1000, 128, 255),
                                                  (1, 2, 2, 300, 64, 17), (2, 4, 4, 4096, 96, 2047),
                                                  (1, 4, 4, 1500, 96, 100), (1, 2, 1, 777, 128, 64)])
@pytest.mark.parametrize("legacy", [False, True])
def test_attention_sliding_window(B, H, KVH, S, hd, window, legacy):
    """Causal sliding-window attention (Phi-3): key j visible iff 0 <= i-j <= window — on the tcgen05
    kernels (hd 96 / 128) and on the mma.sync kernels (legacy=True, and always for hd 64)."""
    from visper_lm_b200 import ops
    qw, kw = H * hd, KVH * hd
    q = rnd(B * S, qw, seed=81)
    kv = rnd(B * S, 2 * kw, seed=82)
    do = rnd(B * S, qw, seed=83)
    k, v = kv[:, :kw], kv[:, kw:]
    scale = hd ** -0.5
    ops.set_option(ops.OPT_ATTN_LEGACY_FWD, int(legacy))
    ops.set_option(ops.OPT_ATTN_LEGACY_BWD, int(legacy))
    try:
        o, lse = ops.attn_fwd(q, k, v, B, H, KVH, S, S, hd, scale, True, window=window)
        dq = torch.empty_like(q)
        dkv = torch.empty_like(kv)
        ops.attn_bwd(q, k, v, o, do, lse, dq, dkv[:, :kw], dkv[:, kw:], B, H, KVH, S, S, hd, scale, True,
                     window=window)
        torch.cuda.synchronize()
    finally:
        ops.set_option(ops.OPT_ATTN_LEGACY_FWD, 0)
        ops.set_option(ops.OPT_ATTN_LEGACY_BWD, 0)
    qf = q.float().view(B, S, H, hd).transpose(1, 2).detach().requires_grad_(True)
    kvf = kv.float().view(B, S, 2, KVH, hd).detach().requires_grad_(True)
    kf = kvf[:, :, 0].transpose(1, 2).repeat_interleave(H // KVH, 1)
    vf = kvf[:, :, 1].transpose(1, 2).repeat_interleave(H // KVH, 1)
    s_ = (qf @ kf.transpose(-1, -2)) * scale
    i = torch.arange(S, device=s_.device)[:, None]
    j = torch.arange(S, device=s_.device)[None, :]
    vis = (j <= i) & (i - j <= window)
    ref = (torch.softmax(s_.masked_fill(~vis, float("-inf")), -1) @ vf).transpose(1, 2).reshape(B * S, qw)
    close(o, ref, name="window fwd")
    ref.backward(do.float())
    close(dq, qf.grad.transpose(1, 2).reshape(B * S, qw), rtol=3e-2, name="window dq")
    close(dkv, kvf.grad.reshape(B * S, 2 * kw), rtol=3e-2, name="window dkv")


@pytest.mark.parametrize("B,H,KVH,sq,sk,causal", [(2, 8, 2, 2048, 2048, True), (1, 4, 4, 333, 333, True),
                                                   (2, 4, 2, 200, 520, False), (1, 4, 1, 128, 128, True),
                                                   (1, 4, 2, 900, 900, True), (1, 2, 2, 257, 1000, True),
                                                   (2, 2, 1, 640, 640, False)])
def test_attention_tcgen05_matches_legacy(B, H, KVH, sq, sk, causal):
    """hd=128 forward: the tcgen05/TMEM kernel against the mma.sync kernel (both vs torch above)."""
    from visper_lm_b200 import ops
    hd = 128
    q = rnd(B * sq, H * hd, seed=55)
    kv = rnd(B * sk, 2 * KVH * hd, seed=56)
    k, v = kv[:, :KVH * hd], kv[:, KVH * hd:]
    ops.set_option(ops.OPT_ATTN_LEGACY_FWD, 1)
    o_ref, lse_ref = ops.attn_fwd(q, k, v, B, H, KVH, sq, sk, hd, hd ** -0.5, causal)
    ops.set_option(ops.OPT_ATTN_LEGACY_FWD, 0)
    o, lse = ops.attn_fwd(q, k, v, B, H, KVH, sq, sk, hd, hd ** -0.5, causal)
    torch.cuda.synchronize()
    close(o, o_ref, name="tc fwd vs legacy")
    assert (lse - lse_ref).abs().max().item() < 2e-2
    # the persistent kernel (the default for launches with many work items, forced here for the small ones) against
    # the one-work-item-per-CTA kernel
    outs = []
    for mode in (1, 2):
        ops.set_option(ops.OPT_ATTN_FWD_NS2, mode)
        try:
            outs.append(ops.attn_fwd(q, k, v, B, H, KVH, sq, sk, hd, hd ** -0.5, causal))
            torch.cuda.synchronize()
        finally:
            ops.set_option(ops.OPT_ATTN_FWD_NS2, 0)
    (o1, lse1), (o2, lse2) = outs
    if sk >= sq:  # same MMAs and softmax; the epilogues differ in rounding only
        close(o2, o1, rtol=4e-3, name="persistent forward vs one-item-per-CTA kernel")
        assert (lse2 - lse1).abs().max().item() < 2e-4
        close(o2, o_ref, name="persistent tc fwd vs legacy")
        assert (lse2 - lse_ref).abs().max().item() < 2e-2


@pytest.mark.parametrize("B,H,KVH,S,hd,causal,window", [(2, 32, 8, 2560, 128, True, 0), (4, 32, 32, 1500, 96, True, 600),
                                                        (8, 20, 4, 1100, 128, False, 0), (16, 16, 16, 577, 64, False, 0)])
def test_persistent_forward_kernel_vs_torch(B, H, KVH, S, hd, causal, window):
    """Launches with >= 8 work items per SM take the persistent forward kernel by default (the decoder's shapes):
    output and LSE against torch fp32 for the first and the last batch entry."""
    from visper_lm_b200 import ops
    g = torch.Generator().manual_seed(S + hd + B)
    qkv = torch.randn(B * S, (H + 2 * KVH) * hd, generator=g).to(BF).to(dev())
    q, k, v = qkv[:, :H * hd], qkv[:, H * hd:(H + KVH) * hd], qkv[:, (H + KVH) * hd:]
    o, lse = ops.attn_fwd(q, k, v, B, H, KVH, S, S, hd, hd ** -0.5, causal, window=window)
    torch.cuda.synchronize()
    i = torch.arange(S, device=q.device)
    vis = torch.ones(S, S, dtype=torch.bool, device=q.device)
    if causal:
        vis = i[None, :] <= i[:, None]
        if window:
            vis = vis & (i[:, None] - i[None, :] <= window)
    for b in (0, B - 1):
        rows = slice(b * S, (b + 1) * S)
        qf = q[rows].float().view(S, H, hd).transpose(0, 1)
        kf = k[rows].float().view(S, KVH, hd).transpose(0, 1).repeat_interleave(H // KVH, 0)
        vf = v[rows].float().view(S, KVH, hd).transpose(0, 1).repeat_interleave(H // KVH, 0)
        sc = (qf @ kf.transpose(1, 2) * hd ** -0.5).masked_fill(~vis[None], float("-inf"))
        ref = (torch.softmax(sc, -1) @ vf).transpose(0, 1).reshape(S, H * hd)
        got = o[rows].float()
        assert ((got - ref).norm() / ref.norm()).item() < 1e-2, (b, ((got - ref).norm() / ref.norm()).item())
        assert (got - ref).abs().max().item() < 3e-2
        assert (lse.view(B, H, S)[b] - torch.logsumexp(sc, -1)).abs().max().item() < 2e-2


@pytest.mark.parametrize("B,H,KVH,sq,sk,causal", [
    (2, 8, 2, 2048, 2048, True), (1, 4, 4, 333, 333, True), (2, 4, 2, 200, 520, False),
    (1, 4, 1, 128, 128, True), (1, 8, 8, 1100, 1100, True), (2, 4, 2, 130, 700, True),
    (1, 2, 1, 64, 64, False), (1, 4, 2, 1537, 1537, False)])
def test_attention_tc_backward_v2_matches_v1_and_torch(B, H, KVH, sq, sk, causal):
    """hd=128 backward: ping-pong (v2) tcgen05 kernels against the v1 tcgen05 kernels and torch."""
    from visper_lm_b200 import ops
    hd = 128
    qw, kw = H * hd, KVH * hd
    q = rnd(B * sq, qw, seed=57)
    kv = rnd(B * sk, 2 * kw, seed=58)
    do = rnd(B * sq, qw, seed=59)
    k, v = kv[:, :kw], kv[:, kw:]
    scale = hd ** -0.5
    o, lse = ops.attn_fwd(q, k, v, B, H, KVH, sq, sk, hd, scale, causal)

    def bwd():
        dq = torch.zeros_like(q)
        dkv = torch.zeros_like(kv)
        ops.attn_bwd(q, k, v, o, do, lse, dq, dkv[:, :kw], dkv[:, kw:], B, H, KVH, sq, sk, hd, scale, causal)
        torch.cuda.synchronize()
        return dq, dkv

    ops.set_option(ops.OPT_ATTN_TC_BWD_V1, 1)
    try:
        dq1, dkv1 = bwd()
    finally:
        ops.set_option(ops.OPT_ATTN_TC_BWD_V1, 0)
    dq2, dkv2 = bwd()
    dq2b, dkv2b = bwd()
    assert torch.equal(dq2, dq2b) and torch.equal(dkv2, dkv2b), "v2 backward is not deterministic"
    # P/dS through shared memory (SS) instead of TMEM (TS, default): same bf16 operands → same bits
    ops.set_option(ops.OPT_ATTN_BWD_SS, 1)
    try:
        dq3, dkv3 = bwd()
    finally:
        ops.set_option(ops.OPT_ATTN_BWD_SS, 0)
    assert torch.equal(dq2, dq3) and torch.equal(dkv2, dkv3), "TS and SS backward variants differ"
    # two softmax groups on alternate iterations instead of the default column split: same bits
    ops.set_option(ops.OPT_ATTN_BWD_PINGPONG, 1)
    try:
        dq4, dkv4 = bwd()
    finally:
        ops.set_option(ops.OPT_ATTN_BWD_PINGPONG, 0)
    assert torch.equal(dq2, dq4) and torch.equal(dkv2, dkv4), "column-split and ping-pong variants differ"
    close(dq2, dq1, rtol=2e-2, name="dq v2 vs v1")
    close(dkv2, dkv1, rtol=2e-2, name="dkv v2 vs v1")
    qf = q.float().view(B, sq, H, hd).transpose(1, 2).detach().requires_grad_(True)
    kvf = kv.float().view(B, sk, 2, KVH, hd).detach().requires_grad_(True)
    kf = kvf[:, :, 0].transpose(1, 2).repeat_interleave(H // KVH, 1)
    vf = kvf[:, :, 1].transpose(1, 2).repeat_interleave(H // KVH, 1)
    ref = _attn_ref(qf, kf, vf, scale, causal).transpose(1, 2).reshape(B * sq, qw)
    ref.backward(do.float())
    close(dq2, qf.grad.transpose(1, 2).reshape(B * sq, qw), rtol=3e-2, name="dq v2 vs torch")
    close(dkv2, kvf.grad.reshape(B * sk, 2 * kw), rtol=3e-2, name="dkv v2 vs torch")


# ------------------------------------------------------------------------------------------- losses
@pytest.mark.parametrize("B,T,V", [(2, 37, 1000), (1, 64, 128256)])
def test_cross_entropy(B, T, V):
    from visper_lm_b200 import ops
    logits = rnd(B * T, V, seed=61, scale=2.0)
    g = torch.Generator().manual_seed(5)
    labels = torch.randint(0, V, (B, T), generator=g)
    labels[:, :5] = -100
    labels = labels.to(dev())
    lf = logits.float().view(B, T, V).requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(lf[:, :-1].reshape(-1, V), labels[:, 1:].reshape(-1),
                                            ignore_index=-100)
    ref.backward()
    cnt = ops.ce_count(labels, T)
    row_loss = torch.empty(B * T, dtype=torch.float32, device=dev())
    work = logits.clone()
    ops.ce_fwd_bwd_(work, labels, 0, T, row_loss, cnt)
    loss = ops.ce_finalize(row_loss, cnt)
    assert abs(loss.item() - ref.item()) < 2e-4 * max(1.0, abs(ref.item())), (loss.item(), ref.item())
    close(work, lf.grad.view(B * T, V), rtol=2e-2, atol=1e-2 * lf.grad.abs().max().item(), name="ce grad")


@pytest.mark.parametrize("B,Bt,off,n", [(4, 4, 0, 1024), (8, 8, 0, 589824), (3, 12, 6, 8200), (8, 64, 16, 24576)])
def test_distill_loss(B, Bt, off, n):
    from visper_lm_b200 import ops
    pred = rnd(B, n, seed=71)
    tgt = rnd(Bt, n, seed=72)
    tgt[off:off + B] = (0.6 * pred.float() + 0.8 * tgt[off:off + B].float()).to(BF)
    tau = torch.tensor(2.0, device=dev())
    mask = torch.tensor([1.0, 0.0, 1.0, 1.0, 1.0, 0.0, 1.0, 1.0][:B], device=dev())
    cw = 0.3
    pf = pred.float().requires_grad_(True)
    tf = tgt.float()
    tauf = tau.clone().requires_grad_(True)
    sl1 = (torch.nn.functional.smooth_l1_loss(pf, tf[off:off + B], reduction="none") * mask[:, None]).mean()
    pn = torch.nn.functional.normalize(pf, dim=-1)
    tn = torch.nn.functional.normalize(tf, dim=-1)
    logits = pn @ tn.t() * torch.clamp(tauf.exp(), max=100)
    ce = torch.nn.functional.cross_entropy(logits, torch.arange(B, device=dev()) + off, reduction="none")
    con = (cw * ce[None, None, :] * mask[:, None, None]).mean()
    total = sl1 + con
    total.backward()
    out4, coef, _ = ops.distill_loss_fwd(pred, tgt, off, tau, mask, cw)
    got = out4.tolist()
    for g_, r_, nm in [(got[0], total.item(), "total"), (got[1], sl1.item(), "sl1"), (got[2], con.item(), "con"),
                       (got[3], tauf.grad.item(), "dtau")]:
        assert abs(g_ - r_) <= 2e-3 * max(abs(r_), 1e-3) + 1e-6, (nm, g_, r_)
    gout = torch.tensor(0.5, device=dev())
    dpred = ops.distill_loss_bwd(pred, tgt, off, coef, gout)
    close(dpred, 0.5 * pf.grad, rtol=2e-2, atol=2e-2 * pf.grad.abs().max().item() * 0.5, name="distill dpred")


# ------------------------------------------------------------------------------------------- misc
def test_gather_and_pool():
    from visper_lm_b200 import ops
    D = 256
    s0, s1 = rnd(40, D, seed=81), rnd(10, D, seed=82)
    kind = torch.tensor([0, 1, -1, 0, 1, 0], dtype=torch.int32, device=dev())
    idx = torch.tensor([3, 9, 0, 39, 0, -1], dtype=torch.int32, device=dev())
    out = ops.gather_rows(idx, [s0, s1], D, kind=kind)
    ref = torch.stack([s0[3], s1[9], torch.zeros_like(s0[0]), s0[39], s1[0], torch.zeros_like(s0[0])])
    assert torch.equal(out, ref)
    sidx = torch.tensor([[0, 1, -1], [5, 5, 6]], dtype=torch.int32, device=dev())
    gs = ops.gather_sum_rows(sidx.flatten(), 3, s0, D)
    close(gs, torch.stack([s0[0].float() + s0[1].float(), 2 * s0[5].float() + s0[6].float()]), name="gather_sum")
    x = rnd(24, D, seed=83)
    gm = ops.group_mean(x, 4, 6)
    close(gm, x.float().view(4, 6, D).mean(1), name="group_mean")
    close(ops.group_mean_bwd(gm, 4, 6), (gm.float() / 6).repeat_interleave(6, 0), name="group_mean_bwd")
    dst = torch.zeros(10, D, dtype=torch.float32, device=dev())
    ops.scatter_add_rows(dst, torch.tensor([1, 1, -1, 7], dtype=torch.int32, device=dev()), s0[:4])
    close(dst[1], s0[0].float() + s0[1].float(), name="scatter_add")
    close(dst[7], s0[3].float(), name="scatter_add")
    t = ops.transpose(s0)
    assert torch.equal(t, s0.t().contiguous())


def test_im2col_clip_embed():
    from visper_lm_b200 import ops
    B, H, P, D = 2, 56, 14, 64
    img = rnd(B, 3, H, H, seed=91)
    w = rnd(D, 3, P, P, seed=92, scale=0.05)
    cls, pos = rnd(D, seed=93), rnd((H // P) ** 2 + 1, D, seed=94)
    K = 3 * P * P
    kpad = 640
    cols = ops.im2col_patches(img, P, kpad)
    wpad = torch.zeros(D, kpad, dtype=BF, device=dev())
    wpad[:, :K] = w.view(D, K)
    patch = ops.gemm(cols, wpad)
    emb = ops.clip_embed(patch, cls, pos, B, (H // P) ** 2)
    ref = torch.nn.functional.conv2d(img.float(), w.float(), stride=P).flatten(2).transpose(1, 2)
    ref = torch.cat([cls.float().expand(B, 1, D), ref], 1) + pos.float()[None]
    close(emb, ref.reshape(-1, D), name="clip embed")


def test_adamw_and_clip():
    from visper_lm_b200 import ops
    n = 10007
    p0 = torch.randn(n, device=dev())
    g = (torch.randn(n, device=dev()) * 0.1).to(BF)
    ref_p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref_p], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    master, m, v = p0.clone(), torch.zeros(n, device=dev()), torch.zeros(n, device=dev())
    param = torch.empty(n, dtype=BF, device=dev())
    ss = ops.grad_sumsq(g)
    assert abs(ss.item() - g.float().pow(2).sum().item()) < 1e-3 * ss.item()
    coef, norm = ops.clip_coef(ss, 1.0, 1.0)
    expect = min(1.0, 1.0 / (g.float().norm().item() + 1e-6))
    assert abs(coef.item() - expect) < 1e-5
    for step in range(1, 4):
        ref_p.grad = g.float() * expect
        opt.step()
        ops.adamw_step_(master, m, v, g, param, 1e-3, 0.9, 0.999, 1e-8, 0.01, step, grad_scale=coef)
    assert (master - ref_p.data).abs().max().item() < 1e-5
    assert torch.equal(param, master.to(BF))


@pytest.mark.parametrize("n,offset", [(1, 0), (7, 1), (8, 0), (9, 3), (2055, 5), (10007, 0), (1 << 20, 2), (5_000_003, 7),
                                      (40_000_000, 0)])
def test_grad_sumsq_alignment_and_sizes(n, offset):
    """vpb_grad_sumsq reads 16 bytes per load: unaligned starts (bucket slices), tails shorter than a vector, sizes from
    one element to more than one pass of the grid."""
    from visper_lm_b200 import ops
    g = torch.Generator().manual_seed(n + offset)
    buf = (torch.randn(n + offset, generator=g) * 0.5).to(BF).to(dev())
    x = buf[offset:]
    got = ops.grad_sumsq(x).item()
    ref = x.double().pow(2).sum().item()
    assert abs(got - ref) <= 2e-5 * ref + 1e-12, (got, ref)


@pytest.mark.parametrize("B,H,KVH,S,hd,window,use_pos,mode", [
    (2, 8, 2, 1024, 128, 0, False, "v2"), (1, 4, 4, 333, 128, 0, True, "v2"),
    (2, 4, 1, 777, 128, 0, True, "v2"), (1, 4, 2, 1300, 128, 500, False, "v2"),
    (1, 4, 2, 640, 128, 0, False, "pingpong"), (1, 4, 2, 640, 128, 0, True, "ss"),
    (1, 4, 2, 500, 128, 0, True, "v1"), (1, 4, 2, 500, 128, 0, False, "legacy"),
    (1, 4, 4, 450, 96, 0, True, "v2"), (1, 4, 2, 300, 64, 0, False, "v2")])
def test_attention_backward_fused_inverse_rope(B, H, KVH, S, hd, window, use_pos, mode):
    """vpb_attn_bwd_rope == vpb_attn_bwd followed by the inverse rope kernel on dQ and dK, bit for bit:
    in the tcgen05 v2 epilogues (hd 128) and through the fallback (v1 / legacy / hd 96 / hd 64)."""
    from visper_lm_b200 import ops
    qw, kw = H * hd, KVH * hd
    qkv = rnd(B * S, qw + 2 * kw, seed=91)
    do = rnd(B * S, qw, seed=92)
    q, k, v = qkv[:, :qw], qkv[:, qw:qw + kw], qkv[:, qw + kw:]
    scale = hd ** -0.5
    cos, sin = ops.rope_tables(4096, hd, 10000.0, qkv.device)
    pos = None
    if use_pos:
        g = torch.Generator().manual_seed(5)
        pos = torch.randint(0, 4096, (B * S,), generator=g).to(torch.int32).cuda()
    opt = {"pingpong": ops.OPT_ATTN_BWD_PINGPONG, "ss": ops.OPT_ATTN_BWD_SS, "v1": ops.OPT_ATTN_TC_BWD_V1,
           "legacy": ops.OPT_ATTN_LEGACY_BWD}.get(mode)
    if opt is not None:
        ops.set_option(opt, 1)
    try:
        o, lse = ops.attn_fwd(q, k, v, B, H, KVH, S, S, hd, scale, True, window=window)
        ref = torch.empty_like(qkv)
        ops.attn_bwd(q, k, v, o, do, lse, ref[:, :qw], ref[:, qw:qw + kw], ref[:, qw + kw:], B, H, KVH, S, S,
                     hd, scale, True, window=window)
        ops.rope_(ref, S, cos, sin, H + KVH, hd, inverse=True, pos_ids=pos)
        got = torch.empty_like(qkv)
        ops.attn_bwd_rope(q, k, v, o, do, lse, got[:, :qw], got[:, qw:qw + kw], got[:, qw + kw:], B, H, KVH,
                          S, hd, scale, True, cos, sin, pos_ids=pos, window=window)
        torch.cuda.synchronize()
    finally:
        if opt is not None:
            ops.set_option(opt, 0)
    assert torch.equal(got[:, :qw], ref[:, :qw]), "dq"
    assert torch.equal(got[:, qw:qw + kw], ref[:, qw:qw + kw]), "dk"
    assert torch.equal(got[:, qw + kw:], ref[:, qw + kw:]), "dv"
    assert ref[:, :qw + kw].float().abs().max() > 0


@pytest.mark.parametrize("B,H,S,nmask", [(8, 6, 144, 4), (3, 2, 144, 0), (18, 4, 16, 9), (5, 1, 16, 0), (4, 3, 100, 2)])
def test_window_attention_with_score_bias(B, H, S, nmask):
    """vpb_attn_fwd_bias (frozen Swin teacher): softmax(q.k/sqrt(32) + bias[h] + mask[b % nmask]) v, head_dim
    32, window sizes 144 (12x12), 16 (4x4) and a ragged 100, against torch fp32."""
    from visper_lm_b200 import ops
    hd = 32
    qkv = rnd(B * S, 3 * H * hd, seed=95)
    g = torch.Generator().manual_seed(96)
    bias = (2.0 * torch.randn(H, S, S, generator=g)).cuda()
    mask = None
    if nmask:
        mask = torch.where(torch.rand(nmask, S, S, generator=g) < 0.3, torch.tensor(-100.0), torch.tensor(0.0))
        mask[:, torch.arange(S), torch.arange(S)] = 0.0   # never a fully masked row (Swin masks keep the diagonal)
        mask = mask.contiguous().cuda()
    W = H * hd
    o = ops.attn_fwd_bias(qkv[:, :W], qkv[:, W:2 * W], qkv[:, 2 * W:], B, H, S, hd, hd ** -0.5, bias, mask)
    torch.cuda.synchronize()
    q, k, v = (qkv.float()[:, i * W:(i + 1) * W].view(B, S, H, hd).transpose(1, 2) for i in range(3))
    sc = q @ k.transpose(-1, -2) * hd ** -0.5 + bias[None]
    if mask is not None:
        sc = sc + mask[torch.arange(B) % nmask][:, None]
    ref = (torch.softmax(sc, -1) @ v).transpose(1, 2).reshape(B * S, W)
    close(o, ref, name="window attention + bias")


@pytest.mark.parametrize("B,Hi,Wi,Ho,Wo,C", [(2, 25, 25, 24, 24, 1536), (1, 4, 4, 24, 24, 256), (3, 7, 5, 3, 9, 64)])
def test_bilinear_half_pixel(B, Hi, Wi, Ho, Wo, C):
    """F.interpolate(mode='bilinear', align_corners=False) on NHWC rows (seg teacher: 25x25 → 24x24)."""
    from visper_lm_b200 import ops
    x = rnd(B * Hi * Wi, C, seed=97)
    y = ops.bilinear(x, B, Hi, Wi, Ho, Wo, C, align_corners=False)
    torch.cuda.synchronize()
    ref = torch.nn.functional.interpolate(x.float().view(B, Hi, Wi, C).permute(0, 3, 1, 2), size=(Ho, Wo),
                                          mode="bilinear", align_corners=False)
    close(y, ref.permute(0, 2, 3, 1).reshape(B * Ho * Wo, C), name="bilinear half-pixel")


@pytest.mark.parametrize("M,D", [(20000, 192), (16384, 384), (33333, 768), (16390, 64), (16384, 776)])
def test_layernorm_many_short_rows(M, D):
    """LayerNorm forward on the warp-per-row kernel (M >= 16384 rows of D <= 768: the Swin teacher's
    shapes; D = 776 stays on the CTA-per-row kernel) against torch fp32, statistics included."""
    from visper_lm_b200 import ops
    x = rnd(M, D, seed=31, scale=2.0)
    w = (1 + 0.1 * rnd(D, seed=32).float()).to(BF)
    b = rnd(D, seed=33, scale=0.1)
    y, mean, rstd = ops.layernorm_fwd(x, w, b, 1e-5)
    torch.cuda.synchronize()
    xf = x.float()
    close(y, torch.nn.functional.layer_norm(xf, (D,), w.float(), b.float(), 1e-5), name="ln fwd (warp rows)")
    assert torch.allclose(mean, xf.mean(1), atol=1e-5)
    assert torch.allclose(rstd, (xf.var(1, unbiased=False) + 1e-5).rsqrt(), rtol=1e-4)


@pytest.mark.parametrize("B,H,KVH,S,hd,window", [(8, 32, 8, 2048, 128, 0), (8, 32, 8, 1000, 128, 0), (4, 32, 32, 2048, 96, 0),
                                                 (4, 32, 32, 1500, 96, 600), (2, 64, 8, 1333, 128, 0)])
def test_persistent_dq_kernel_is_bit_identical(B, H, KVH, S, hd, window):
    """The persistent dQ kernel (one CTA per SM, Q / dO and dQ double-buffered, epilogue warpgroup) against the
    round-1 kernel it replaces (VPB_OPT_ATTN_BWD_DQ_R1): same MMAs in the same order → dq, dk, dv bit for bit; both
    against torch fp32 on a sub-sample."""
    from visper_lm_b200 import ops
    g = torch.Generator().manual_seed(S + hd)
    qkv = torch.randn(B * S, (H + 2 * KVH) * hd, generator=g).to(BF).to(dev())
    q, k, v = qkv[:, :H * hd], qkv[:, H * hd:(H + KVH) * hd], qkv[:, (H + KVH) * hd:]
    o, lse = ops.attn_fwd(q, k, v, B, H, KVH, S, S, hd, hd ** -0.5, True, window=window)
    do = torch.randn(B * S, H * hd, generator=g).to(BF).to(dev())
    outs = []
    for r1 in (1, 0):
        d = torch.zeros_like(qkv)
        ops.set_option(ops.OPT_ATTN_BWD_DQ_R1, r1)
        try:
            ops.attn_bwd(q, k, v, o, do, lse, d[:, :H * hd], d[:, H * hd:(H + KVH) * hd], d[:, (H + KVH) * hd:], B, H, KVH, S, S,
                         hd, hd ** -0.5, True, window=window)
            torch.cuda.synchronize()
        finally:
            ops.set_option(ops.OPT_ATTN_BWD_DQ_R1, 0)
        outs.append(d)
    assert torch.equal(outs[0], outs[1]), f"max diff {(outs[0].float() - outs[1].float()).abs().max().item()}"
    # torch fp32 reference for batch entry 0, two heads
    b0 = slice(0, S)
    for h in (0, H - 1):
        kvh = h // (H // KVH)
        qh = q[b0, h * hd:(h + 1) * hd].float().requires_grad_(True)
        kh = k[b0, kvh * hd:(kvh + 1) * hd].float()
        vh = v[b0, kvh * hd:(kvh + 1) * hd].float()
        sc = qh @ kh.t() * hd ** -0.5
        i = torch.arange(S, device=sc.device)
        vis = i[None, :] <= i[:, None]
        if window:
            vis = vis & (i[:, None] - i[None, :] <= window)
        pr = torch.softmax(sc.masked_fill(~vis, float("-inf")), -1)
        (pr @ vh).backward(do[b0, h * hd:(h + 1) * hd].float())
        got = outs[1][b0, h * hd:(h + 1) * hd].float()
        assert ((got - qh.grad).norm() / qh.grad.norm()).item() < 2e-2


@pytest.mark.parametrize("B,H,W,C,stride,relu", [(2, 12, 12, 64, 1, False), (3, 24, 24, 256, 2, False), (2, 9, 13, 128, 1, True),
                                                 (1, 48, 48, 256, 1, True), (1, 7, 5, 8, 2, False)])
def test_im2col3x3_matches_torch_unfold(B, H, W, C, stride, relu):
    """vpb_im2col3x3_nhwc (3x3, pad 1, K order (ky, kx, c), optional ReLU on load) against F.unfold."""
    from visper_lm_b200 import ops
    x = rnd(B * H * W, C, seed=H * W + C)
    col, Ho, Wo = ops.im2col3x3(x, B, H, W, C, stride, relu)
    torch.cuda.synchronize()
    xi = x.float().view(B, H, W, C).permute(0, 3, 1, 2)
    if relu:
        xi = xi.relu()
    u = torch.nn.functional.unfold(xi, 3, padding=1, stride=stride)          # [B, C*9, Ho*Wo], K order (c, ky, kx)
    ref = u.view(B, C, 9, Ho * Wo).permute(0, 3, 2, 1).reshape(B * Ho * Wo, 9 * C)
    assert (Ho, Wo) == ((H - 1) // stride + 1, (W - 1) // stride + 1)
    assert torch.equal(col.float(), ref)
