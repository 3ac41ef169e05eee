"""Thin tensor-level wrappers over the C ABI (no autograd here — see autograd.py).

Every function takes CUDA bf16/fp32 torch tensors, passes raw device pointers, strides and the
current CUDA stream to libvisper_b200.so and returns torch tensors that own the outputs.  torch is
only the allocator / stream provider; there is no torch compute and no CPU fallback.
"""
from __future__ import annotations

import ctypes
import os

import torch

from . import lib as _lib

ACT_NONE, ACT_GELU, ACT_QUICK_GELU, ACT_RELU = 0, 1, 2, 3
BF16 = torch.bfloat16


def _L():
    return _lib.load()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return 0 if t is None else t.data_ptr()


def _rows(t: torch.Tensor):
    """(ptr, ld) of a 2-D row-major view (unit inner stride)."""
    assert t.dim() == 2 and (t.shape[1] == 1 or t.stride(1) == 1), (t.shape, t.stride())
    return t.data_ptr(), t.stride(0)


OPT_ATTN_LEGACY_FWD, OPT_ATTN_LEGACY_BWD, OPT_ATTN_TC_BWD_V1, OPT_GEMM_1CTA, OPT_GEMM_PANEL_MB = 0, 1, 2, 3, 4
OPT_ATTN_BWD_SS, OPT_ATTN_BWD_PINGPONG, OPT_GEMM_L2_HINTS = 5, 7, 8
OPT_ATTN_BWD_DQ_R1 = 6
OPT_ATTN_FWD_NS2 = 9
OPT_NORM_R1 = 15
OPT_WIN_ATTN_V2 = 10
OPT_DWCONV_FFMA2 = 11
OPT_ATTN_FWD_TC64 = 12
OPT_GEMM_EPI8 = 13
OPT_GATHER_FLAT = 14


def set_option(key, value):
    _chk(_L().vpb_set_option(int(key), int(value)), "set_option")


def _chk(status, what):
    if status != 0:
        _lib.check(status, what)


# --------------------------------------------------------------------------------------------- GEMM
class KernelTimer:
    """CUDA-event timing of the named hot kernels on the launching stream (bench.py's live rooflines):
    every record is (start event, end event, category, algorithmic FLOPs, algorithmic bytes)."""

    def __init__(self):
        self.records = []

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for s, e, cat, fl, by in self.records:
            d = out.setdefault(cat, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            d["launches"] += 1
            d["ms"] += s.elapsed_time(e)
            d["flops"] += fl
            d["bytes"] += by
        for d in out.values():
            d["tflops"] = (d["flops"] / d["ms"] / 1e9) if (d["ms"] > 0 and d["flops"]) else None
        return out


KERNEL_TIMER = None


def _timed(flops, cat="gemm", nbytes=0.0):
    timer = KERNEL_TIMER
    if timer is None:
        return None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    return timer, ev0, ev1, cat, flops, nbytes


def _timed_end(t):
    if t is not None:
        timer, ev0, ev1, cat, flops, nbytes = t
        ev1.record()
        timer.records.append((ev0, ev1, cat, flops, nbytes))


def gemm(a, b, *, a_layout=0, b_layout=0, bias=None, act=ACT_NONE, residual=None, out=None,
         want_pre=False):
    """C = act(A·Bᵀ + bias) + residual.  a_layout/b_layout as in include/visper_b200.h."""
    assert a.dtype == BF16 and b.dtype == BF16
    if a_layout == 0:
        M, K = a.shape
    else:
        K, M = a.shape
    if b_layout == 0:
        N, Kb = b.shape
    else:
        Kb, N = b.shape
    assert K == Kb, (a.shape, b.shape, a_layout, b_layout)
    if out is None:
        out = torch.empty((M, N), dtype=BF16, device=a.device)
    pre = torch.empty((M, N), dtype=BF16, device=a.device) if want_pre else None
    pa, lda = _rows(a)
    pb, ldb = _rows(b)
    pc, ldc = _rows(out)
    pr, ldr = _rows(residual) if residual is not None else (0, 0)
    px, ldx = _rows(pre) if pre is not None else (0, 0)
    t = _timed(2.0 * M * N * K)
    _chk(_L().vpb_gemm_bf16(pa, lda, a_layout, pb, ldb, b_layout, pc, ldc, M, N, K, act, _p(bias),
                            pr, ldr, px, ldx, _stream()), "gemm")
    _timed_end(t)
    return (out, pre) if want_pre else out


# A/B switches for the fused SwiGLU GEMM epilogues (0 selects the separate kernels).  Measured in
# the Llama-3-8B step, same box, alternating runs (profiles/r01_swiglu_fusion_ab.txt): the forward
# fusion wins ~1.7 % of the step; the backward fusion (with the tile-major g|u layout) wins 0.4 ms
# per layer in isolation but is a wash inside the power-capped step, so it stays off by default.
FUSE_SWIGLU = os.environ.get("VPB_FUSE_SWIGLU", "1") != "0"
FUSE_SWIGLU_BWD = os.environ.get("VPB_FUSE_SWIGLU_BWD", "0") != "0"


def gemm_swiglu_fwd(a, wgu, want_gu=True, tiled=False):
    """(h, gu): gu = a·wguᵀ (None unless want_gu) and h[M,F] = silu(gate)·up, one launch.
    tiled: gu is the flat tile-major buffer only gemm_swiglu_bwd(tiled=True) can read."""
    M, K = a.shape
    F2, Kb = wgu.shape
    F = F2 // 2
    assert K == Kb and F % 128 == 0
    h = torch.empty((M, F), dtype=BF16, device=a.device)
    gu = None
    if want_gu:
        gu = (torch.empty((((M + 127) // 128) * 128 * F2,), dtype=BF16, device=a.device) if tiled
              else torch.empty((M, F2), dtype=BF16, device=a.device))
    pa, lda = _rows(a)
    pw, ldw = _rows(wgu)
    pg, ldg = (gu.data_ptr(), F2) if gu is not None else (0, 0)
    t = _timed(2.0 * M * F2 * K)
    _chk(_L().vpb_gemm_swiglu_fwd(pa, lda, pw, ldw, pg, ldg, 1 if tiled else 0, h.data_ptr(), F, M, F, K,
                                  _stream()), "gemm_swiglu_fwd")
    _timed_end(t)
    return h, gu


def gemm_swiglu_bwd(dy, w, gu, b_layout=1, tiled=False, F=None):
    """dgu[M,2F] = swiglu'(gu) ∘ (dy·W); w is down_proj [D,F] (b_layout 1) or its transpose [F,D]."""
    M, K = dy.shape
    if F is None:
        F = gu.shape[1] // 2
    assert (w.shape == (K, F)) if b_layout == 1 else (w.shape == (F, K)), (w.shape, K, F)
    dgu = torch.empty((M, 2 * F), dtype=BF16, device=dy.device)
    pd, ldd = _rows(dy)
    pw, ldw = _rows(w)
    t = _timed(2.0 * M * F * K)
    _chk(_L().vpb_gemm_swiglu_bwd(pd, ldd, pw, ldw, b_layout, gu.data_ptr(), 2 * F, 1 if tiled else 0,
                                  dgu.data_ptr(), 2 * F, M, F, K, _stream()), "gemm_swiglu_bwd")
    _timed_end(t)
    return dgu


FUSE_ROPE = os.environ.get("VPB_FUSE_ROPE", "1") != "0"
# inverse RoPE of dQ/dK inside the attention-backward epilogues (head_dim 128).  Bit-identical to the
# separate kernel but not faster: the backward CTAs own the whole SM (512 TMEM columns), so their
# epilogue is exposed, while the stand-alone rope kernel runs at L2 speed on the still-resident dqkv
# (profiles/r01_rope_bwd_fusion_ab.txt) — off by default.
FUSE_ROPE_BWD = os.environ.get("VPB_FUSE_ROPE_BWD", "0") == "1"


def gemm_rope(a, w, seq_len, cos, sin, rope_heads, pos_ids=None):
    """qkv[M,N] = a·wᵀ with RoPE applied in the epilogue to the first rope_heads 128-wide heads."""
    M, K = a.shape
    N, Kb = w.shape
    assert K == Kb and N % 256 == 0
    out = torch.empty((M, N), dtype=BF16, device=a.device)
    pa, lda = _rows(a)
    pw, ldw = _rows(w)
    t = _timed(2.0 * M * N * K)
    _chk(_L().vpb_gemm_rope_bf16(pa, lda, pw, ldw, out.data_ptr(), N, M, N, K, cos.data_ptr(), sin.data_ptr(),
                                 seq_len, _p(pos_ids), rope_heads, _stream()), "gemm_rope")
    _timed_end(t)
    return out


def transpose(x, out=None):
    R, C = x.shape
    if out is None:
        out = torch.empty((C, R), dtype=BF16, device=x.device)
    pi, ldi = _rows(x)
    po, ldo = _rows(out)
    _chk(_L().vpb_transpose(pi, ldi, po, ldo, R, C, _stream()), "transpose")
    return out


# --------------------------------------------------------------------------------------------- norms
def rmsnorm_fwd(x, w, eps, out=None):
    M, D = x.shape
    y = out if out is not None else torch.empty((M, D), dtype=BF16, device=x.device)
    rstd = torch.empty((M,), dtype=torch.float32, device=x.device)
    px, ldx = _rows(x)
    py, ldy = _rows(y)
    t = _timed(0.0, "rmsnorm_fwd", 4.0 * M * D)  # read x + write y, bf16
    _chk(_L().vpb_rmsnorm_fwd(px, ldx, w.data_ptr(), py, ldy, rstd.data_ptr(), M, D, eps, _stream()),
         "rmsnorm_fwd")
    _timed_end(t)
    return y, rstd


def rmsnorm_bwd(dy, x, w, rstd, dres=None, out=None):
    M, D = x.shape
    dx = out if out is not None else torch.empty((M, D), dtype=BF16, device=x.device)
    pdy, lddy = _rows(dy)
    px, ldx = _rows(x)
    pr, ldr = _rows(dres) if dres is not None else (0, 0)
    pdx, lddx = _rows(dx)
    t = _timed(0.0, "rmsnorm_bwd", (8.0 if dres is not None else 6.0) * M * D)  # dy, x (, dres) in; dx out
    _chk(_L().vpb_rmsnorm_bwd(pdy, lddy, px, ldx, w.data_ptr(), rstd.data_ptr(), pr, ldr,
                              pdx, lddx, M, D, _stream()), "rmsnorm_bwd")
    _timed_end(t)
    return dx


def layernorm_fwd(x, w, b, eps, out=None):
    M, D = x.shape
    y = out if out is not None else torch.empty((M, D), dtype=BF16, device=x.device)
    mean = torch.empty((M,), dtype=torch.float32, device=x.device)
    rstd = torch.empty((M,), dtype=torch.float32, device=x.device)
    px, ldx = _rows(x)
    py, ldy = _rows(y)
    _chk(_L().vpb_layernorm_fwd(px, ldx, w.data_ptr(), b.data_ptr(), py, ldy, mean.data_ptr(),
                                rstd.data_ptr(), M, D, eps, _stream()), "layernorm_fwd")
    return y, mean, rstd


def layernorm_bwd(dy, x, w, mean, rstd, dres=None, out=None):
    M, D = x.shape
    dx = out if out is not None else torch.empty((M, D), dtype=BF16, device=x.device)
    pdy, lddy = _rows(dy)
    px, ldx = _rows(x)
    pr, ldr = _rows(dres) if dres is not None else (0, 0)
    pdx, lddx = _rows(dx)
    _chk(_L().vpb_layernorm_bwd(pdy, lddy, px, ldx, w.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                pr, ldr, pdx, lddx, M, D, _stream()), "layernorm_bwd")
    return dx


def colsum(a, b=None, mean=None, rstd=None, out_dtype=BF16):
    """sum over rows of a (* normalised b) → [N] (fp32 accumulate, returned in out_dtype)."""
    M, N = a.shape
    acc = torch.empty((N,), dtype=torch.float32, device=a.device)
    pa, lda = _rows(a)
    pb, ldb = _rows(b) if b is not None else (0, 0)
    _chk(_L().vpb_colsum(pa, lda, pb, ldb, _p(mean), _p(rstd), acc.data_ptr(), M, N, _stream()),
         "colsum")
    if out_dtype == torch.float32:
        return acc
    return cast_bf16(acc)


def cast_bf16(x_f32, scale=1.0):
    out = torch.empty(x_f32.shape, dtype=BF16, device=x_f32.device)
    _chk(_L().vpb_cast_f32_bf16(x_f32.data_ptr(), out.data_ptr(), x_f32.numel(), scale, _stream()),
         "cast")
    return out


# --------------------------------------------------------------------------------------------- elementwise
_rope_cache = {}


def rope_tables(max_pos, head_dim, theta, device):
    key = (max_pos, head_dim, float(theta), str(device))
    if key not in _rope_cache:
        cos = torch.empty((max_pos, head_dim // 2), dtype=torch.float32, device=device)
        sin = torch.empty_like(cos)
        _chk(_L().vpb_rope_table(cos.data_ptr(), sin.data_ptr(), max_pos, head_dim, theta, _stream()),
             "rope_table")
        _rope_cache[key] = (cos, sin)
    return _rope_cache[key]


def rope_(x, seq_len, cos, sin, nheads, head_dim, inverse=False, pos_ids=None):
    """In-place rotary embedding of the first `nheads` heads of packed rows x[M, ld]."""
    px, ld = _rows(x)
    _chk(_L().vpb_rope_inplace(px, ld, x.shape[0], seq_len, _p(pos_ids), cos.data_ptr(),
                               sin.data_ptr(), nheads, head_dim, 1 if inverse else 0, _stream()),
         "rope")
    return x


def swiglu_fwd(gu):
    M, F2 = gu.shape
    F = F2 // 2
    h = torch.empty((M, F), dtype=BF16, device=gu.device)
    pg, ldg = _rows(gu)
    _chk(_L().vpb_swiglu_fwd(pg, ldg, h.data_ptr(), F, M, F, _stream()), "swiglu_fwd")
    return h


def swiglu_bwd(gu, dh):
    M, F2 = gu.shape
    F = F2 // 2
    dgu = torch.empty((M, F2), dtype=BF16, device=gu.device)
    pg, ldg = _rows(gu)
    pd, ldd = _rows(dh)
    t = _timed(0.0, "swiglu_bwd", 10.0 * M * F)  # g|u (2F) + dh (F) in, d_gate|d_up (2F) out, bf16
    _chk(_L().vpb_swiglu_bwd(pg, ldg, pd, ldd, dgu.data_ptr(), F2, M, F, _stream()), "swiglu_bwd")
    _timed_end(t)
    return dgu


def act_bwd(pre, dy, act):
    M, N = pre.shape
    dx = torch.empty((M, N), dtype=BF16, device=pre.device)
    pp, ldp = _rows(pre)
    pd, ldd = _rows(dy)
    _chk(_L().vpb_act_bwd(pp, ldp, pd, ldd, dx.data_ptr(), N, M, N, act, _stream()), "act_bwd")
    return dx


def axpby(a, b=None, alpha=1.0, beta=1.0, out=None):
    a = a.contiguous()
    if b is not None:
        b = b.contiguous()
    out = out if out is not None else torch.empty_like(a)
    _chk(_L().vpb_axpby(a.data_ptr(), _p(b), out.data_ptr(), alpha, beta, a.numel(), _stream()),
         "axpby")
    return out


def scale_dev(x, scale_f32):
    x = x.contiguous()
    out = torch.empty_like(x)
    _chk(_L().vpb_scale_dev(x.data_ptr(), out.data_ptr(), scale_f32.data_ptr(), x.numel(), _stream()),
         "scale_dev")
    return out


# --------------------------------------------------------------------------------------------- CLIP embed
def im2col_patches(images, patch, kpad):
    B, C, H, W = images.shape
    assert C == 3 and images.is_contiguous() and images.dtype == BF16
    n = (H // patch) * (W // patch)
    out = torch.empty((B * n, kpad), dtype=BF16, device=images.device)
    _chk(_L().vpb_im2col_patches(images.data_ptr(), out.data_ptr(), B, H, W, patch, kpad, _stream()),
         "im2col")
    return out


def clip_embed(patch, cls, pos, B, npatch):
    D = patch.shape[1]
    out = torch.empty((B * (npatch + 1), D), dtype=BF16, device=patch.device)
    _chk(_L().vpb_clip_embed(patch.data_ptr(), cls.data_ptr(), pos.data_ptr(), out.data_ptr(), B,
                             npatch, D, _stream()), "clip_embed")
    return out


# --------------------------------------------------------------------------------------------- DPT decoder
def im2col3x3(x, B, H, W, C, stride=1, relu_in=False):
    """NHWC rows [B*H*W, C] → [B*Ho*Wo, 9*C] (3x3, pad 1), K order (ky, kx, c)."""
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    assert x.is_contiguous() and x.shape == (B * H * W, C)
    out = torch.empty((B * Ho * Wo, 9 * C), dtype=BF16, device=x.device)
    _chk(_L().vpb_im2col3x3_nhwc(x.data_ptr(), out.data_ptr(), B, H, W, C, stride, 1 if relu_in else 0,
                                 _stream()), "im2col3x3")
    return out, Ho, Wo


def bilinear(x, B, Hi, Wi, Ho, Wo, C, align_corners=True):
    assert x.is_contiguous() and x.shape == (B * Hi * Wi, C)
    out = torch.empty((B * Ho * Wo, C), dtype=BF16, device=x.device)
    fn = _L().vpb_bilinear_nhwc if align_corners else _L().vpb_bilinear_nhwc_half_pixel
    _chk(fn(x.data_ptr(), out.data_ptr(), B, Hi, Wi, Ho, Wo, C, _stream()), "bilinear")
    return out


def pixel_shuffle(y, bias, B, H, W, C, k):
    assert y.is_contiguous() and y.shape == (B * H * W, k * k * C)
    out = torch.empty((B * H * k * W * k, C), dtype=BF16, device=y.device)
    _chk(_L().vpb_pixel_shuffle_nhwc(y.data_ptr(), _p(bias), out.data_ptr(), B, H, W, C, k, _stream()),
         "pixel_shuffle")
    return out


def conv1x1_to1(x, w, bias, relu=True):
    P, C = x.shape
    assert x.is_contiguous()
    out = torch.empty((P,), dtype=torch.float32, device=x.device)
    _chk(_L().vpb_conv1x1_to1(x.data_ptr(), w.data_ptr(), _p(bias), out.data_ptr(), P, C, 1 if relu else 0,
                              _stream()), "conv1x1_to1")
    return out


def minmax_normalize(x):
    B, n = x.shape[0], x[0].numel()
    x = x.contiguous()
    out = torch.empty_like(x)
    _chk(_L().vpb_minmax_normalize(x.data_ptr(), out.data_ptr(), B, n, _stream()), "minmax_normalize")
    return out


# --------------------------------------------------------------------------------------------- ConvNeXt tower
def dwconv7x7(x, w49, bias, B, H, W, C, out=None):
    """Depthwise 7x7, pad 3, on NHWC rows [B*H*W, C]; w49 = the [C,1,7,7] filter repacked to [49, C]."""
    assert x.is_contiguous() and x.shape == (B * H * W, C) and x.dtype == BF16
    assert w49.is_contiguous() and w49.shape == (49, C) and w49.dtype == BF16
    if out is None:
        out = torch.empty_like(x)
    _chk(_L().vpb_dwconv7x7_nhwc(x.data_ptr(), w49.data_ptr(), _p(bias), out.data_ptr(), B, H, W, C, _stream()),
         "dwconv7x7")
    return out


# --------------------------------------------------------------------------------------------- gathers
def gather_rows(index, srcs, D, kind=None, out=None):
    """out[r] = srcs[kind[r]][index[r]] (negative → zero row). index/kind: int32 CUDA tensors."""
    n = index.numel()
    dev = index.device
    if out is None:
        out = torch.empty((n, D), dtype=BF16, device=dev)
    po, ldo = _rows(out)
    ptrs = []
    for i in range(4):
        if i < len(srcs) and srcs[i] is not None:
            ptrs += list(_rows(srcs[i]))
        else:
            ptrs += [0, 0]
    _chk(_L().vpb_gather_rows(po, ldo, n, D, _p(kind), index.data_ptr(), *ptrs, _stream()),
         "gather_rows")
    return out


def gather_sum_rows(index, cnt, src, D, scale=1.0):
    nslots = index.numel() // cnt
    out = torch.empty((nslots, D), dtype=BF16, device=src.device)
    ps, lds = _rows(src)
    _chk(_L().vpb_gather_sum_rows(out.data_ptr(), D, nslots, cnt, index.data_ptr(), ps, lds, D, scale,
                                  _stream()), "gather_sum_rows")
    return out


def scatter_add_rows(dst_f32, index, src):
    n, D = src.shape
    ps, lds = _rows(src)
    _chk(_L().vpb_scatter_add_rows(dst_f32.data_ptr(), dst_f32.stride(0), n, index.data_ptr(), ps,
                                   lds, D, _stream()), "scatter_add_rows")
    return dst_f32


def add_rows_(dst, index, src):
    n, D = src.shape
    pd, ldd = _rows(dst)
    ps, lds = _rows(src)
    _chk(_L().vpb_add_rows(pd, ldd, n, index.data_ptr(), ps, lds, D, _stream()), "add_rows")
    return dst


def group_mean(x, groups, gsize, out=None):
    D = x.shape[1]
    if out is None:
        out = torch.empty((groups, D), dtype=BF16, device=x.device)
    px, ldx = _rows(x)
    po, ldo = _rows(out)
    _chk(_L().vpb_group_mean(px, ldx, po, ldo, groups, gsize, D, _stream()), "group_mean")
    return out


def group_mean_bwd(dout, groups, gsize):
    D = dout.shape[1]
    din = torch.empty((groups * gsize, D), dtype=BF16, device=dout.device)
    pd, ldd = _rows(dout)
    _chk(_L().vpb_group_mean_bwd(pd, ldd, din.data_ptr(), D, groups, gsize, D, _stream()),
         "group_mean_bwd")
    return din


# --------------------------------------------------------------------------------------------- attention
def _attn_flops(B, H, sq, sk, hd, causal, window):
    """QK^T + PV of the visible (query, key) pairs: causal = lower triangle, sliding window = band."""
    pairs = float(sq) * sk
    if causal:
        pairs = sq * (sq + 1) / 2.0
        if window and window < sq:
            w = window + 1
            pairs = w * (w + 1) / 2.0 + (sq - w) * float(w)
    return 4.0 * B * H * pairs * hd


def attn_fwd(q, k, v, B, H, KVH, sq, sk, head_dim, scale, causal, k2=None, v2=None, sk2=0, out=None,
             window=0):
    """q:[B*sq, >=H*hd] k,v:[B*sk, >=KVH*hd] row views (may alias one packed buffer)."""
    dev = q.device
    o = out if out is not None else torch.empty((B * sq, H * head_dim), dtype=BF16, device=dev)
    lse = torch.empty((B, H, sq), dtype=torch.float32, device=dev)
    pq, ldq = _rows(q)
    pk, ldk = _rows(k)
    pv, ldv = _rows(v)
    pk2, ldk2 = _rows(k2) if k2 is not None else (0, 0)
    pv2, ldv2 = _rows(v2) if v2 is not None else (0, 0)
    po, ldo = _rows(o)
    t = _timed(_attn_flops(B, H, sq, sk + sk2, head_dim, causal, window), f"attn_fwd_hd{head_dim}")
    _chk(_L().vpb_attn_fwd(pq, ldq, pk, ldk, pv, ldv, pk2, ldk2, pv2, ldv2, po, ldo, lse.data_ptr(),
                           B, H, KVH, sq, sk, sk2, head_dim, scale, 1 if causal else 0, int(window),
                           _stream()), "attn_fwd")
    _timed_end(t)
    return o, lse


def attn_fwd_bias(q, k, v, B, H, S, head_dim, scale, bias, mask=None):
    """Window attention with additive bias: softmax(scale*q.k + bias[h] + mask[b % len(mask)]) v.
    q/k/v: row views [B*S, H*head_dim]; bias fp32 [H,S,S]; mask fp32 [nW,S,S] or None."""
    assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.shape == (H, S, S)
    assert mask is None or (mask.dtype == torch.float32 and mask.is_contiguous() and mask.shape[1:] == (S, S))
    o = torch.empty((B * S, H * head_dim), dtype=BF16, device=q.device)
    lse = torch.empty((B, H, S), dtype=torch.float32, device=q.device)
    pq, ldq = _rows(q)
    pk, ldk = _rows(k)
    pv, ldv = _rows(v)
    _chk(_L().vpb_attn_fwd_bias(pq, ldq, pk, ldk, pv, ldv, o.data_ptr(), H * head_dim, lse.data_ptr(), B, H,
                                S, S, head_dim, scale, bias.data_ptr(), _p(mask),
                                0 if mask is None else mask.shape[0], _stream()), "attn_fwd_bias")
    return o


def attn_bwd(q, k, v, o, do, lse, dq, dk, dv, B, H, KVH, sq, sk, head_dim, scale, causal,
             k2=None, v2=None, sk2=0, dk2=None, dv2=None, window=0):
    """Writes dq/dk/dv (and dk2/dv2) row views in place."""
    delta = torch.empty((B, H, sq), dtype=torch.float32, device=q.device)
    pq, ldq = _rows(q)
    pk, ldk = _rows(k)
    pv, ldv = _rows(v)
    pk2, ldk2 = _rows(k2) if k2 is not None else (0, 0)
    pv2, ldv2 = _rows(v2) if v2 is not None else (0, 0)
    po, ldo = _rows(o)
    pdo, lddo = _rows(do)
    pdq, lddq = _rows(dq)
    pdk, lddk = _rows(dk)
    pdv, lddv = _rows(dv)
    pdk2, lddk2 = _rows(dk2) if dk2 is not None else (0, 0)
    pdv2, lddv2 = _rows(dv2) if dv2 is not None else (0, 0)
    t = _timed(2.5 * _attn_flops(B, H, sq, sk + sk2, head_dim, causal, window), f"attn_bwd_hd{head_dim}")
    _chk(_L().vpb_attn_bwd(pq, ldq, pk, ldk, pv, ldv, pk2, ldk2, pv2, ldv2, po, ldo, pdo, lddo,
                           lse.data_ptr(), delta.data_ptr(), pdq, lddq, pdk, lddk, pdv, lddv, pdk2,
                           lddk2, pdv2, lddv2, B, H, KVH, sq, sk, sk2, head_dim, scale,
                           1 if causal else 0, int(window), _stream()), "attn_bwd")
    _timed_end(t)


def attn_bwd_rope(q, k, v, o, do, lse, dq, dk, dv, B, H, KVH, T, head_dim, scale, causal, cos, sin,
                  pos_ids=None, window=0):
    """Self-attention backward whose dq/dk come back with the inverse RoPE already applied
    (== attn_bwd followed by rope_(inverse=True) on the q and k heads, bit for bit)."""
    delta = torch.empty((B, H, T), dtype=torch.float32, device=q.device)
    pq, ldq = _rows(q)
    pk, ldk = _rows(k)
    pv, ldv = _rows(v)
    po, ldo = _rows(o)
    pdo, lddo = _rows(do)
    pdq, lddq = _rows(dq)
    pdk, lddk = _rows(dk)
    pdv, lddv = _rows(dv)
    _chk(_L().vpb_attn_bwd_rope(pq, ldq, pk, ldk, pv, ldv, po, ldo, pdo, lddo, lse.data_ptr(),
                                delta.data_ptr(), pdq, lddq, pdk, lddk, pdv, lddv, B, H, KVH, T,
                                head_dim, scale, 1 if causal else 0, int(window), cos.data_ptr(),
                                sin.data_ptr(), _p(pos_ids), _stream()), "attn_bwd_rope")


# --------------------------------------------------------------------------------------------- losses
def ce_count(labels, T, shift=True):
    cnt = torch.empty((1,), dtype=torch.float32, device=labels.device)
    _chk(_L().vpb_ce_count(labels.data_ptr(), labels.numel(), T, 1 if shift else 0, cnt.data_ptr(),
                           _stream()), "ce_count")
    return cnt


def ce_fwd_bwd_(logits, labels, row0, T, row_loss, count, gscale=1.0, write_grad=True, shift=True):
    R, V = logits.shape
    pl, ld = _rows(logits)
    t = _timed(0.0, "ce_fwd_bwd", (6.0 if write_grad else 2.0) * R * V)  # 2 reads (+1 write) of bf16 logits
    _chk(_L().vpb_ce_fwd_bwd(pl, ld, labels.data_ptr(), row0, R, V, T, 1 if shift else 0,
                             row_loss.data_ptr(), count.data_ptr(), gscale, 1 if write_grad else 0,
                             _stream()), "ce_fwd_bwd")
    _timed_end(t)


def ce_finalize(row_loss, count):
    loss = torch.empty((), dtype=torch.float32, device=row_loss.device)
    _chk(_L().vpb_ce_finalize(row_loss.data_ptr(), row_loss.numel(), count.data_ptr(),
                              loss.data_ptr(), _stream()), "ce_finalize")
    return loss


def distill_loss_fwd(pred, tgt, off, tau, mask, cw, want_stats=False):
    """pred [B,n], tgt [Bt,n] bf16; tau fp32 scalar tensor; mask fp32 [B] or None."""
    B, n = pred.shape
    Bt = tgt.shape[0]
    dev = pred.device
    ws = torch.empty((int(_L().vpb_distill_workspace_floats(B, Bt, n)),), dtype=torch.float32, device=dev)
    out4 = torch.empty((4,), dtype=torch.float32, device=dev)
    coef = torch.empty((2 * B + B * Bt,), dtype=torch.float32, device=dev)
    stats = torch.empty((B * Bt + 2 * B + Bt,), dtype=torch.float32, device=dev) if want_stats else None
    pp, ldp = _rows(pred)
    pt, ldt = _rows(tgt)
    _chk(_L().vpb_distill_loss_fwd(pp, ldp, pt, ldt, n, B, Bt, off, tau.data_ptr(), _p(mask), cw,
                                   ws.data_ptr(), out4.data_ptr(), coef.data_ptr(), _p(stats),
                                   _stream()), "distill_loss_fwd")
    return out4, coef, stats


def distill_loss_bwd(pred, tgt, off, coef, gout):
    B, n = pred.shape
    Bt = tgt.shape[0]
    dpred = torch.empty((B, n), dtype=BF16, device=pred.device)
    pp, ldp = _rows(pred)
    pt, ldt = _rows(tgt)
    _chk(_L().vpb_distill_loss_bwd(pp, ldp, pt, ldt, n, B, Bt, off, coef.data_ptr(), _p(gout),
                                   dpred.data_ptr(), n, _stream()), "distill_loss_bwd")
    return dpred


# --------------------------------------------------------------------------------------------- optimizer
def adamw_step_(master, m, v, grad, param, lr, beta1, beta2, eps, wd, step, grad_scale=None):
    _chk(_L().vpb_adamw_step(master.data_ptr(), m.data_ptr(), v.data_ptr(), grad.data_ptr(),
                             param.data_ptr(), master.numel(), lr, beta1, beta2, eps, wd, step,
                             _p(grad_scale), _stream()), "adamw")


def grad_sumsq(grad, out=None, accumulate=False):
    ws = torch.empty((1024,), dtype=torch.float32, device=grad.device)
    if out is None:
        out = torch.zeros((1,), dtype=torch.float32, device=grad.device)
    _chk(_L().vpb_grad_sumsq(grad.data_ptr(), grad.numel(), ws.data_ptr(), out.data_ptr(),
                             1 if accumulate else 0, _stream()), "grad_sumsq")
    return out


def clip_coef(sumsq, max_norm, extra_scale=1.0):
    coef = torch.empty((1,), dtype=torch.float32, device=sumsq.device)
    norm = torch.empty((1,), dtype=torch.float32, device=sumsq.device)
    _chk(_L().vpb_clip_coef(sumsq.data_ptr(), max_norm, extra_scale, coef.data_ptr(), norm.data_ptr(),
                            _stream()), "clip_coef")
    return coef, norm


if os.environ.get("VPB_GEMM_PANEL_MB"):
    set_option(OPT_GEMM_PANEL_MB, int(os.environ["VPB_GEMM_PANEL_MB"]))
for _name, _key in (("VPB_ATTN_BWD_PINGPONG", OPT_ATTN_BWD_PINGPONG), ("VPB_ATTN_BWD_SS", OPT_ATTN_BWD_SS),
                    ("VPB_GEMM_1CTA", OPT_GEMM_1CTA),
                    ("VPB_GEMM_L2_HINTS", OPT_GEMM_L2_HINTS), ("VPB_WIN_ATTN_V2", OPT_WIN_ATTN_V2),
                    ("VPB_DWCONV_FFMA2", OPT_DWCONV_FFMA2), ("VPB_ATTN_FWD_TC64", OPT_ATTN_FWD_TC64),
                    ("VPB_GEMM_EPI8", OPT_GEMM_EPI8), ("VPB_GATHER_FLAT", OPT_GATHER_FLAT),
                    ("VPB_ATTN_FWD_NS2", OPT_ATTN_FWD_NS2), ("VPB_ATTN_BWD_DQ_R1", OPT_ATTN_BWD_DQ_R1), ("VPB_NORM_R1", OPT_NORM_R1)):
    if os.environ.get(_name):  # A/B switches for bench runs
        set_option(_key, int(os.environ[_name]))
