"""Kernel variants selected by vpb_set_option: each is checked against torch fp32 AND against the kernel it
replaced (option value 0).  All ran on a B200 in round 2 (profiles/r02_variants_ab.txt); the five that won
their A/B — one-pass window attention, packed-FMA depthwise, tcgen05 ViT attention, eight epilogue warps,
flat gather — are now the library defaults (`DEFAULT` below mirrors csrc/api.cu)."""
import contextlib

import pytest
import torch

pytestmark = pytest.mark.gpu
DEFAULT = {10: 1, 11: 1, 12: 1, 13: 1, 14: 1}


@contextlib.contextmanager
def option(key, value):
    from visper_lm_b200 import ops

    ops.set_option(key, value)
    try:
        yield
    finally:
        torch.cuda.synchronize()
        ops.set_option(key, DEFAULT.get(key, 0))


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("B,H,S,nmask", [(8, 6, 144, 4), (3, 2, 144, 0), (18, 4, 16, 9), (5, 1, 16, 0), (4, 3, 100, 2),
                                         (2, 2, 99, 2)])
def test_one_pass_window_attention(B, H, S, nmask, variant):
    """VPB_OPT_WIN_ATTN_V2: vpb_attn_fwd_bias on the one-pass kernel == torch fp32 and == the default kernel."""
    from visper_lm_b200 import ops

    hd = 32
    g = torch.Generator().manual_seed(95 + S)
    qkv = torch.randn(B * S, 3 * H * hd, generator=g).to(torch.bfloat16).cuda()
    bias = (2.0 * torch.randn(H, S, S, generator=g)).cuda()
    mask = None
    if nmask:
        mask = torch.where(torch.rand(nmask, S, S, generator=g) < 0.3, torch.tensor(-100.0), torch.tensor(0.0))
        mask[:, torch.arange(S), torch.arange(S)] = 0.0
        mask = mask.contiguous().cuda()
    W = H * hd
    with option(ops.OPT_WIN_ATTN_V2, 0):
        base = ops.attn_fwd_bias(qkv[:, :W], qkv[:, W:2 * W], qkv[:, 2 * W:], B, H, S, hd, hd ** -0.5, bias, mask)
    with option(ops.OPT_WIN_ATTN_V2, variant):
        o = ops.attn_fwd_bias(qkv[:, :W], qkv[:, W:2 * W], qkv[:, 2 * W:], B, H, S, hd, hd ** -0.5, bias, mask)
    q, k, v = (qkv.float()[:, i * W:(i + 1) * W].view(B, S, H, hd).transpose(1, 2) for i in range(3))
    sc = q @ k.transpose(-1, -2) * hd ** -0.5 + bias[None]
    if mask is not None:
        sc = sc + mask[torch.arange(B) % nmask][:, None]
    ref = (torch.softmax(sc, -1) @ v).transpose(1, 2).reshape(B * S, W)
    assert ((o.float() - ref).norm() / ref.norm()).item() < 1.6e-2
    assert ((o.float() - base.float()).norm() / base.float().norm()).item() < 1e-2


@pytest.mark.parametrize("B,H,W,C", [(1, 5, 9, 64), (2, 24, 24, 128), (2, 48, 48, 1536)])
def test_dwconv_packed_fma_is_bit_identical(B, H, W, C):
    """VPB_OPT_DWCONV_FFMA2: fma.rn.f32x2 per channel pair == the validated scalar-FMA depthwise kernel, bit for bit."""
    from visper_lm_b200 import ops

    g = torch.Generator().manual_seed(C + H)
    x = torch.randn(B * H * W, C, generator=g).to(torch.bfloat16).cuda()
    w49 = (torch.randn(49, C, generator=g) / 7).to(torch.bfloat16).cuda()
    b = torch.randn(C, generator=g).to(torch.bfloat16).cuda()
    with option(ops.OPT_DWCONV_FFMA2, 0):
        base = ops.dwconv7x7(x, w49, b, B, H, W, C)
    with option(ops.OPT_DWCONV_FFMA2, 1):
        got = ops.dwconv7x7(x, w49, b, B, H, W, C)
    assert torch.equal(got, base)


@pytest.mark.parametrize("B,H,S", [(2, 16, 577), (1, 16, 1370), (3, 4, 128), (2, 2, 70)])
def test_vit_attention_on_tcgen05(B, H, S):
    """VPB_OPT_ATTN_FWD_TC64: non-causal head_dim-64 attention forward (CLIP ViT-L: S = 577, DINOv2-L at 518 px:
    S = 1370) on the tcgen05 kernel == torch fp32, and its LSE == the mma.sync kernel's."""
    from visper_lm_b200 import ops

    hd = 64
    g = torch.Generator().manual_seed(S)
    qkv = torch.randn(B * S, 3 * H * hd, generator=g).to(torch.bfloat16).cuda()
    W = H * hd
    with option(ops.OPT_ATTN_FWD_TC64, 0):
        base, lse0 = ops.attn_fwd(qkv[:, :W], qkv[:, W:2 * W], qkv[:, 2 * W:], B, H, H, S, S, hd, hd ** -0.5, False)
    with option(ops.OPT_ATTN_FWD_TC64, 1):
        o, lse = ops.attn_fwd(qkv[:, :W], qkv[:, W:2 * W], qkv[:, 2 * W:], B, H, H, S, S, hd, hd ** -0.5, False)
    q, k, v = (qkv.float()[:, i * W:(i + 1) * W].view(B, S, H, hd).transpose(1, 2) for i in range(3))
    sc = q @ k.transpose(-1, -2) * hd ** -0.5
    ref = (torch.softmax(sc, -1) @ v).transpose(1, 2).reshape(B * S, W)
    assert ((o.float() - ref).norm() / ref.norm()).item() < 1.6e-2
    assert ((o.float() - base.float()).norm() / base.float().norm()).item() < 1e-2
    assert torch.allclose(lse, torch.logsumexp(sc, -1), atol=2e-2) and torch.allclose(lse, lse0, atol=2e-2)


@pytest.mark.parametrize("M,N,K,act,res", [(20000, 768, 192, 1, False), (18432, 1536, 384, 0, True),
                                           (8192, 520, 200, 1, True), (4096, 6144, 1024, 1, False)])
def test_gemm_eight_epilogue_warps_is_bit_identical(M, N, K, act, res):
    """VPB_OPT_GEMM_EPI8: the CTA-pair GEMM with eight epilogue warps (K <= 1024) == the default kernel bit for bit
    (same MMAs, same per-element epilogue arithmetic; only which warp drains which columns changes)."""
    from visper_lm_b200 import ops

    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16).cuda()
    b = torch.randn(N, generator=g).to(torch.bfloat16).cuda()
    r = torch.randn(M, N, generator=g).to(torch.bfloat16).cuda() if res else None
    with option(ops.OPT_GEMM_EPI8, 0):
        base = ops.gemm(a, w, bias=b, act=act, residual=r)
    with option(ops.OPT_GEMM_EPI8, 1):
        got = ops.gemm(a, w, bias=b, act=act, residual=r)
    assert torch.equal(got, base)
    ref = a.float() @ w.float().t() + b.float()
    ref = torch.nn.functional.gelu(ref) if act == 1 else ref
    ref = ref + r.float() if res else ref
    assert ((got.float() - ref).norm() / ref.norm()).item() < 1e-2


@pytest.mark.parametrize("n,D,nsrc", [(5000, 192, 1), (333, 1536, 1), (1000, 64, 3), (7, 2048, 2)])
def test_flat_gather_is_bit_identical(n, D, nsrc):
    """VPB_OPT_GATHER_FLAT: grid-stride gather == the CTA-per-row kernel (zero rows, several sources, strided out)."""
    from visper_lm_b200 import ops

    g = torch.Generator().manual_seed(n + D)
    srcs = [torch.randn(400 + 10 * i, D, generator=g).to(torch.bfloat16).cuda() for i in range(nsrc)]
    index = torch.randint(-1, 400, (n,), generator=g, dtype=torch.int32).cuda()
    kind = torch.randint(0, nsrc, (n,), generator=g, dtype=torch.int32).cuda() if nsrc > 1 else None
    wide0 = torch.full((n, 2 * D), 7.0, dtype=torch.bfloat16, device="cuda")
    wide1 = wide0.clone()
    with option(ops.OPT_GATHER_FLAT, 0):
        ops.gather_rows(index, srcs, D, kind=kind, out=wide0[:, D:])
    with option(ops.OPT_GATHER_FLAT, 1):
        ops.gather_rows(index, srcs, D, kind=kind, out=wide1[:, D:])
    assert torch.equal(wide0, wide1)
    k = kind.long() if kind is not None else torch.zeros(n, dtype=torch.long, device="cuda")
    ref = torch.stack([srcs[int(k[r])][int(index[r])] if index[r] >= 0 else torch.zeros(D, dtype=torch.bfloat16, device="cuda")
                       for r in range(min(n, 300))])
    assert torch.equal(wide1[:300, D:], ref)
