"""Frozen segmentation teacher on the GPU, batched: the Swin-L backbone of OneFormer
whose last feature map, resized to 24x24, is the `seg` distillation target.

Reference call site: _get_seg_targets (language_model/base_ola_vlm.py:382-397) →
OneFormerHead.forward_features (aux_heads/oneformer_head.py:42-69) →
AuxOneFormerPixelLevelModule.forward(return_features=True) (:15-35) =
`F.interpolate(self.encoder(pixel_values).feature_maps[-1], size=(24, 24), mode="bilinear")`, one image
at a time.  `encoder` is transformers' SwinBackbone (third-party, not in the reference tree); it is
restated here with the HF parameter names (`pixel_level_module.encoder.*`) so the
oneformer_coco_swin_large weights load unchanged (the transformer decoder half is never run on this
path and is not built).

All compute runs through the C ABI: im2col + tcgen05 GEMM for the 4x4 patch conv, LayerNorm, one
gather kernel for  pad → cyclic shift → window partition  (zero rows for the padding, exactly what
F.pad after layernorm_before produces), fused-QKV GEMM, window attention with the relative-position
bias and the shifted-window mask added to the scores in the kernel (vpb_attn_fwd_bias), one gather for
window reverse → un-shift → crop, output-dense / MLP GEMMs with bias, GELU and the residual in their
epilogues, patch merging as four strided gathers into one [N/4, 4C] buffer + LayerNorm + GEMM, and a
half-pixel bilinear kernel for the final 25x25 → 24x24 resize.  Index plans, bias tables and fused
weights are built once per geometry / weight version.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..ops import ACT_GELU, BF16
from .modules import Linear, Norm, Weight, _param

SWIN_L = dict(embed_dim=192, depths=(2, 2, 18, 2), num_heads=(6, 12, 24, 48), window_size=12, patch_size=4)
EPS = 1e-5
MAX_GRID_Z = 65535


def relative_position_index(ws):
    """SwinSelfAttention.create_relative_position_index → [ws*ws, ws*ws] int64."""
    c = torch.stack(torch.meshgrid([torch.arange(ws), torch.arange(ws)], indexing="ij")).flatten(1)
    rel = (c[:, :, None] - c[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def shift_mask(Hp, Wp, ws, shift):
    """SwinLayer.get_attn_mask → fp32 [nW, ws*ws, ws*ws] of {0, -100} (window order = partition order)."""
    img = torch.zeros(Hp, Wp)
    cnt = 0
    for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[hs, wsl] = cnt
            cnt += 1
    mw = img.view(Hp // ws, ws, Wp // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)
    m = mw.unsqueeze(1) - mw.unsqueeze(2)
    return torch.where(m != 0, torch.full_like(m, -100.0), torch.zeros_like(m)).contiguous()


def window_plans(B, H, W, ws, shift):
    """int32 index plans of one Swin layer on a B x H x W token grid.
    part[r]: source token row (or -1 = padding) of window-major row r — F.pad to a multiple of ws,
             torch.roll(-shift), window_partition (modeling_swin.py SwinLayer.forward).
    rev[t]:  window-major row holding token t after window_reverse, torch.roll(+shift) and the crop."""
    Hp, Wp = (H + ws - 1) // ws * ws, (W + ws - 1) // ws * ws
    y = torch.arange(Hp).view(Hp, 1).expand(Hp, Wp)
    x = torch.arange(Wp).view(1, Wp).expand(Hp, Wp)
    sy, sx = (y + shift) % Hp, (x + shift) % Wp           # shifted[y, x] = padded[(y+s)%Hp, (x+s)%Wp]
    src = torch.where((sy < H) & (sx < W), sy * W + sx, torch.full_like(sy, -1))
    src = src.view(Hp // ws, ws, Wp // ws, ws).permute(0, 2, 1, 3).reshape(-1)      # window-major
    bofs = (torch.arange(B) * (H * W)).view(B, 1)
    part = torch.where(src.view(1, -1) >= 0, src.view(1, -1) + bofs, torch.full((1, 1), -1, dtype=torch.long))
    # token (y, x) of the un-shifted grid sits at shifted position ((y - s) % Hp, (x - s) % Wp)
    ty = (torch.arange(H).view(H, 1) - shift) % Hp
    tx = (torch.arange(W).view(1, W) - shift) % Wp
    wrow = ((ty // ws) * (Wp // ws) + (tx // ws)) * (ws * ws) + (ty % ws) * ws + (tx % ws)
    rev = wrow.reshape(1, -1) + (torch.arange(B) * (Hp * Wp)).view(B, 1)
    return part.reshape(-1).to(torch.int32), rev.reshape(-1).to(torch.int32), Hp, Wp


def merge_plans(B, H, W):
    """SwinPatchMerging: four int32 gathers (0::2,0::2), (1::2,0::2), (0::2,1::2), (1::2,1::2); odd sides
    are padded with zeros (-1)."""
    H2, W2 = (H + 1) // 2, (W + 1) // 2
    yy = torch.arange(H2).view(H2, 1) * 2
    xx = torch.arange(W2).view(1, W2) * 2
    bofs = (torch.arange(B) * (H * W)).view(B, 1)
    out = []
    for dy, dx in ((0, 0), (1, 0), (0, 1), (1, 1)):
        y, x = yy + dy, xx + dx
        src = torch.where((y < H) & (x < W), y * W + x, torch.full_like(y * W + x, -1)).reshape(1, -1)
        out.append(torch.where(src >= 0, src + bofs, torch.full((1, 1), -1, dtype=torch.long))
                   .reshape(-1).to(torch.int32))
    return out, H2, W2


class _SelfAttention(nn.Module):
    def __init__(self, C, heads, ws, device):
        super().__init__()
        self.query = Linear(C, C, True, device)
        self.key = Linear(C, C, True, device)
        self.value = Linear(C, C, True, device)
        self.relative_position_bias_table = _param((2 * ws - 1) ** 2, heads, device=device)


class _Dense(nn.Module):
    def __init__(self, i, o, device):
        super().__init__()
        self.dense = Linear(i, o, True, device)


class _Attention(nn.Module):
    def __init__(self, C, heads, ws, device):
        super().__init__()
        self.self = _SelfAttention(C, heads, ws, device)
        self.output = _Dense(C, C, device)


class SwinLayer(nn.Module):
    def __init__(self, C, heads, ws, device):
        super().__init__()
        self.layernorm_before = Norm(C, True, device)
        self.attention = _Attention(C, heads, ws, device)
        self.layernorm_after = Norm(C, True, device)
        self.intermediate = _Dense(C, 4 * C, device)
        self.output = _Dense(4 * C, C, device)
        self.heads, self.ws = heads, ws
        self._cache = None

    def derived(self, rel_index):
        """(fused qkv weight [3C, C], bias [3C], score bias fp32 [heads, S, S]) per weight version."""
        a = self.attention.self
        src = (a.query.weight, a.key.weight, a.value.weight, a.query.bias, a.key.bias, a.value.bias,
               a.relative_position_bias_table)
        key = tuple((t.data_ptr(), t._version) for t in src)
        if self._cache is None or self._cache[0] != key:
            with torch.no_grad():
                w = torch.cat([a.query.weight, a.key.weight, a.value.weight], 0).detach().contiguous()
                b = torch.cat([a.query.bias, a.key.bias, a.value.bias], 0).detach().contiguous()
                S = self.ws * self.ws
                tab = a.relative_position_bias_table.detach().float()[rel_index.view(-1).to(w.device)]
                bias = tab.view(S, S, self.heads).permute(2, 0, 1).contiguous()
            self._cache = (key, (w, b, bias))
        return self._cache[1]


class _Merge(nn.Module):
    def __init__(self, C, device):
        super().__init__()
        self.reduction = Linear(4 * C, 2 * C, False, device)
        self.norm = Norm(4 * C, True, device)


class _Stage(nn.Module):
    def __init__(self, C, depth, heads, ws, downsample, device):
        super().__init__()
        self.blocks = nn.ModuleList([SwinLayer(C, heads, ws, device) for _ in range(depth)])
        if downsample:
            self.downsample = _Merge(C, device)


class _Encoder(nn.Module):
    def __init__(self, cfg, device):
        super().__init__()
        E, n = cfg["embed_dim"], len(cfg["depths"])
        self.layers = nn.ModuleList([_Stage(E * 2 ** s, cfg["depths"][s], cfg["num_heads"][s], cfg["window_size"],
                                            s + 1 < n, device) for s in range(n)])


class _PatchEmbeddings(nn.Module):
    def __init__(self, E, P, device):
        super().__init__()
        self.projection = Weight((E, 3, P, P), (E,), device)


class _Embeddings(nn.Module):
    def __init__(self, E, P, device):
        super().__init__()
        self.patch_embeddings = _PatchEmbeddings(E, P, device)
        self.norm = Norm(E, True, device)


class SwinBackbone(nn.Module):
    """Forward-only HF SwinBackbone (always_partition=True, no absolute position embeddings)."""

    def __init__(self, cfg=None, device=None):
        super().__init__()
        cfg = dict(SWIN_L if cfg is None else cfg)
        self.cfg = cfg
        E, P = cfg["embed_dim"], cfg["patch_size"]
        self.embeddings = _Embeddings(E, P, device)
        self.encoder = _Encoder(cfg, device)
        self.hidden_states_norms = nn.ModuleDict(
            {f"stage{s + 1}": Norm(E * 2 ** s, True, device) for s in range(len(cfg["depths"]))})
        self.requires_grad_(False)
        self._rel_index = relative_position_index(cfg["window_size"])
        self._plans = {}
        self._patch_w = None

    @property
    def device(self):
        return self.embeddings.norm.weight.device

    def _patch_weight(self):
        w = self.embeddings.patch_embeddings.projection.weight
        key = (w.data_ptr(), w._version)
        if self._patch_w is None or self._patch_w[0] != key:
            K = w[0].numel()
            kpad = (K + 63) // 64 * 64
            wp = torch.zeros((w.shape[0], kpad), dtype=BF16, device=w.device)
            wp[:, :K] = w.detach().reshape(w.shape[0], K)
            self._patch_w = (key, wp, kpad)
        return self._patch_w[1], self._patch_w[2]

    def _plan(self, kind, *args):
        key = (kind,) + args
        if key not in self._plans:
            dev = self.device
            if kind == "win":
                part, rev, Hp, Wp = window_plans(*args)
                B, H, W, ws, shift = args
                mask = shift_mask(Hp, Wp, ws, shift).to(dev) if shift else None
                self._plans[key] = (part.to(dev), rev.to(dev), Hp, Wp, mask)
            else:
                idx, H2, W2 = merge_plans(*args)
                self._plans[key] = ([i.to(dev) for i in idx], H2, W2)
        return self._plans[key]

    @torch.no_grad()
    def last_feature_rows(self, pixel_values):
        """pixel_values [B,3,H,W] → (feature_maps[-1] as NHWC rows [B*h*w, C4] bf16, h, w)."""
        cfg = self.cfg
        E, ws, P = cfg["embed_dim"], cfg["window_size"], cfg["patch_size"]
        x = pixel_values.to(device=self.device, non_blocking=True)
        B, _, Hi, Wi = x.shape
        if Hi % P or Wi % P:
            x = F.pad(x, (0, (P - Wi % P) % P, 0, (P - Hi % P) % P))   # SwinPatchEmbeddings.maybe_pad
        x = ops.cast_bf16(x.contiguous()) if x.dtype == torch.float32 else x.to(BF16).contiguous()
        H, W = x.shape[2] // P, x.shape[3] // P
        wp, kpad = self._patch_weight()
        cols = ops.im2col_patches(x, P, kpad)
        pe = self.embeddings.patch_embeddings.projection
        x = ops.gemm(cols, wp, bias=pe.bias)
        del cols
        x, _, _ = ops.layernorm_fwd(x, self.embeddings.norm.weight, self.embeddings.norm.bias, EPS)
        n_stage = len(cfg["depths"])
        for s, stage in enumerate(self.encoder.layers):
            C = E * 2 ** s
            heads = cfg["num_heads"][s]
            hd = C // heads
            for i, blk in enumerate(stage.blocks):
                shift = 0 if i % 2 == 0 else ws // 2
                part, rev, Hp, Wp, mask = self._plan("win", B, H, W, ws, shift)
                nW = (Hp // ws) * (Wp // ws)
                wqkv, bqkv, bias = blk.derived(self._rel_index)
                h, _, _ = ops.layernorm_fwd(x, blk.layernorm_before.weight, blk.layernorm_before.bias, EPS)
                win = ops.gather_rows(part, [h], C)
                qkv = ops.gemm(win, wqkv, bias=bqkv)
                del win, h
                S = ws * ws
                per = max(1, MAX_GRID_Z // nW)                        # images per launch (gridDim.z limit)
                if per >= B:
                    ctx = ops.attn_fwd_bias(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B * nW, heads, S, hd,
                                            hd ** -0.5, bias, mask)
                else:
                    ctx = torch.cat([ops.attn_fwd_bias(q_[:, :C], q_[:, C:2 * C], q_[:, 2 * C:],
                                                       q_.shape[0] // S, heads, S, hd, hd ** -0.5, bias, mask)
                                     for q_ in qkv.split(per * nW * S)], 0)
                del qkv
                tok = ops.gather_rows(rev, [ctx], C)
                od = blk.attention.output.dense
                x = ops.gemm(tok, od.weight, bias=od.bias, residual=x)
                del tok, ctx
                h, _, _ = ops.layernorm_fwd(x, blk.layernorm_after.weight, blk.layernorm_after.bias, EPS)
                f = ops.gemm(h, blk.intermediate.dense.weight, bias=blk.intermediate.dense.bias, act=ACT_GELU)
                x = ops.gemm(f, blk.output.dense.weight, bias=blk.output.dense.bias, residual=x)
                del f, h
            if s + 1 < n_stage:
                idx, H2, W2 = self._plan("merge", B, H, W)
                cat = torch.empty((B * H2 * W2, 4 * C), dtype=BF16, device=x.device)
                for k, ix in enumerate(idx):
                    ops.gather_rows(ix, [x], C, out=cat[:, k * C:(k + 1) * C])
                ds = stage.downsample
                h, _, _ = ops.layernorm_fwd(cat, ds.norm.weight, ds.norm.bias, EPS)
                x = ops.gemm(h, ds.reduction.weight)
                H, W = H2, W2
        nrm = self.hidden_states_norms[f"stage{n_stage}"]
        y, _, _ = ops.layernorm_fwd(x, nrm.weight, nrm.bias, 1e-5)
        return y, H, W


class _PixelLevelModule(nn.Module):
    def __init__(self, cfg, device):
        super().__init__()
        self.encoder = SwinBackbone(cfg, device)


class OneFormerHead(nn.Module):
    """The reference's `self.oneformer` as far as training uses it: forward_features(pixel_values)."""

    def __init__(self, cfg=None, device=None):
        super().__init__()
        self.pixel_level_module = _PixelLevelModule(cfg, device)
        self.requires_grad_(False)

    @torch.no_grad()
    def seg_target_rows(self, pixel_values):
        """[B*576, C4] bf16, token-major (the layout the seg head's predictions have)."""
        enc = self.pixel_level_module.encoder
        y, H, W = enc.last_feature_rows(pixel_values)
        B = pixel_values.shape[0]
        if (H, W) != (24, 24):
            y = ops.bilinear(y, B, H, W, 24, 24, y.shape[1], align_corners=False)
        return y

    @torch.no_grad()
    def forward_features(self, pixel_values, task_inputs=None, **_):
        """oneformer_head.py:42-69 → [B, C4, 24, 24] (task_inputs are unused there too)."""
        y = self.seg_target_rows(pixel_values)
        B, C = pixel_values.shape[0], y.shape[1]
        out = torch.empty((B, C, 576), dtype=BF16, device=y.device)
        for b in range(B):
            ops.transpose(y[b * 576:(b + 1) * 576], out=out[b])
        return out.view(B, C, 24, 24)
