"""Frozen depth teacher on the GPU, batched: the DINOv2 backbone of Depth-Anything-V2
that produces the depth distillation targets.

Reference: ola_vlm/model/aux_heads/depth_anything_v2/dinov2.py (DinoVisionTransformer: patch embed
:70-80 of dinov2_layers/patch_embed.py, interpolate_pos_encoding :179-208, prepare_tokens :210-231,
get_intermediate_layers :278-303), dinov2_layers/block.py:76-100, attention.py:50-66, mlp.py:31-37,
layer_scale.py; DepthAnythingV2 (depth_anything_v2/dpt.py:153-221); called one image at a time,
batch 1, from _get_dav2_feats (language_model/base_ola_vlm.py:348-366).  Here the whole batch goes
through one pass.  Parameter names and shapes equal the reference's (`pretrained.*`, `depth_head.*`),
so depth_anything_v2_vitl.pth loads unchanged.

All compute runs through the C ABI (no torch math on the data path): im2col + tcgen05 GEMM for the
14x14 patch conv, LayerNorm, fused-QKV GEMM with bias, flash attention (head_dim 64, non-causal),
projection / fc2 GEMMs with the LayerScale gain folded into their (frozen) weights and the residual
in the epilogue, fc1 GEMM with erf-GELU in the epilogue, and one gather-sum kernel that drops the cls
row and averages the four normed taps.  Derived weights (LayerScale folds, the pixel normalisation and
the reference's channel reversal folded into the patch conv, the interpolated position table) are
built once per weight version — a one-time layout change of frozen weights, not step work.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..ops import ACT_GELU, BF16
from .dpt import DPTHead
from .modules import Linear, Norm, Weight, _param

ARCH = {"vits": dict(dim=384, depth=12, heads=6, taps=(2, 5, 8, 11)),     # dinov2.py:346-395,
        "vitb": dict(dim=768, depth=12, heads=12, taps=(2, 5, 8, 11)),    # dpt.py:164-169
        "vitl": dict(dim=1024, depth=24, heads=16, taps=(4, 11, 17, 23))}
MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)                  # dpt.py:205
PATCH = 14
EPS = 1e-6
CENTER_NET = tuple(round(255.0 * m) for m in MEAN)   # per NET channel; raw channel r uses CENTER_NET[2 - r]


class _Gamma(nn.Module):  # LayerScale (layer_scale.py:17-28)
    def __init__(self, D, device):
        super().__init__()
        self.gamma = _param(D, device=device)


class _Attention(nn.Module):
    def __init__(self, D, device):
        super().__init__()
        self.qkv = Linear(D, 3 * D, True, device)
        self.proj = Linear(D, D, True, device)


class _Mlp(nn.Module):
    def __init__(self, D, device):
        super().__init__()
        self.fc1 = Linear(D, 4 * D, True, device)
        self.fc2 = Linear(4 * D, D, True, device)


class Block(nn.Module):
    def __init__(self, D, device):
        super().__init__()
        self.norm1 = Norm(D, True, device)
        self.attn = _Attention(D, device)
        self.ls1 = _Gamma(D, device)
        self.norm2 = Norm(D, True, device)
        self.mlp = _Mlp(D, device)
        self.ls2 = _Gamma(D, device)
        self._folded = None

    def folded(self):
        """(proj_w, proj_b, fc2_w, fc2_b) with the LayerScale gains multiplied in (fp32, rounded once)."""
        src = (self.attn.proj.weight, self.attn.proj.bias, self.ls1.gamma, self.mlp.fc2.weight,
               self.mlp.fc2.bias, self.ls2.gamma)
        key = tuple((t.data_ptr(), t._version) for t in src)
        if self._folded is None or self._folded[0] != key:
            pw, pb, g1, fw, fb, g2 = (t.detach().float() for t in src)
            self._folded = (key, ((g1[:, None] * pw).to(BF16), (g1 * pb).to(BF16),
                                  (g2[:, None] * fw).to(BF16), (g2 * fb).to(BF16)))
        return self._folded[1]


class _PatchEmbed(nn.Module):
    def __init__(self, D, device):
        super().__init__()
        self.proj = Weight((D, 3, PATCH, PATCH), (D,), device)


class DinoVisionTransformer(nn.Module):
    """Forward-only DINOv2 ViT (no register tokens, LayerScale, GELU Mlp; interpolate_offset 0.1)."""

    def __init__(self, encoder="vitl", device=None):
        super().__init__()
        a = ARCH[encoder]
        D = a["dim"]
        self.embed_dim, self.num_heads, self.taps = D, a["heads"], a["taps"]
        self.cls_token = _param(1, 1, D, device=device)
        self.pos_embed = _param(1, (518 // PATCH) ** 2 + 1, D, device=device)
        self.mask_token = _param(1, D, device=device)  # unused at inference; kept for the state dict
        self.patch_embed = _PatchEmbed(D, device)
        self.blocks = nn.ModuleList([Block(D, device) for _ in range(a["depth"])])
        self.norm = Norm(D, True, device)
        self.requires_grad_(False)
        self._pos = {}
        self._patch = {}
        self._idx = {}

    # ---- derived, cached ----------------------------------------------------------------------
    def _pos_table(self, w0, h0):
        """[1 + w0*h0, D] bf16: interpolate_pos_encoding (dinov2.py:179-208) — bicubic with
        scale_factor (n + 0.1)/sqrt(N), computed once per grid in fp32."""
        ver = (self.pos_embed.data_ptr(), self.pos_embed._version)
        key = (w0, h0) + ver
        if key not in self._pos:
            self._pos = {k: v for k, v in self._pos.items() if k[2:] == ver}  # drop tables of old weights
            pe = self.pos_embed.detach().float()
            N = pe.shape[1] - 1
            if not (w0 * h0 == N and w0 == h0):
                M = int(math.sqrt(N))
                grid = pe[:, 1:].reshape(1, M, M, -1).permute(0, 3, 1, 2)
                sx, sy = float(w0 + 0.1) / math.sqrt(N), float(h0 + 0.1) / math.sqrt(N)
                grid = F.interpolate(grid, scale_factor=(sx, sy), mode="bicubic", antialias=False)
                assert grid.shape[-2:] == (w0, h0)
                pe = torch.cat([pe[:, :1], grid.permute(0, 2, 3, 1).reshape(1, w0 * h0, -1)], 1)
            self._pos[key] = pe[0].to(BF16).contiguous()
        return self._pos[key]

    def _patch_weight(self, raw):
        """([D, kpad] bf16, bias [D] bf16) of the 14x14 patch conv as a GEMM operand.  raw=True folds
        image2tensor (dpt.py:194-221) into it: the input is then the uint8 RGB image minus CENTER
        (integers, exact in bf16), and  conv(w, (flip(x)/255 - mean)/std) + b  =  conv(w', x - ctr) + b'.
        Centering keeps b' ~ b and halves the input magnitude, so the single bf16 rounding of w'
        costs no more than the rounding of w itself."""
        w, b = self.patch_embed.proj.weight, self.patch_embed.proj.bias
        key = (raw, w.data_ptr(), w._version, b._version)
        if key not in self._patch:
            wf, bf = w.detach().float(), b.detach().float()
            if raw:
                std = torch.tensor(STD, device=w.device).view(1, 3, 1, 1)
                mean = torch.tensor(MEAN, device=w.device).view(1, 3, 1, 1)
                ctr = torch.tensor(CENTER_NET, device=w.device, dtype=torch.float32).view(1, 3, 1, 1)
                bf = bf + (wf * ((ctr - 255.0 * mean) / (255.0 * std))).sum(dim=(1, 2, 3))
                wf = (wf / (255.0 * std)).flip(1)  # net channel c reads raw channel 2-c (the BGR2RGB call)
            K = wf[0].numel()
            kpad = (K + 63) // 64 * 64
            wp = torch.zeros((wf.shape[0], kpad), dtype=BF16, device=w.device)
            wp[:, :K] = wf.reshape(wf.shape[0], K)
            self._patch = {k: v for k, v in self._patch.items() if k[0] != raw}
            self._patch[key] = (wp, bf.to(BF16), kpad)
        return self._patch[key]

    def _tap_mean_index(self, B, S):
        if (B, S) not in self._idx:
            t = torch.arange(B * S, dtype=torch.int32).view(B, S)[:, 1:].reshape(-1, 1)      # drop cls
            idx = t + torch.arange(4, dtype=torch.int32).view(1, 4) * (B * S)                 # 4 taps
            self._idx[(B, S)] = idx.reshape(-1).contiguous().to(self.norm.weight.device)
        return self._idx[(B, S)]

    # ---- compute ------------------------------------------------------------------------------
    @torch.no_grad()
    def normed_taps(self, images, raw=False):
        """images [B,3,H,W] bf16 (raw=True: RGB pixels minus CENTER, see _raw_batch; False: normalised) →
        ([4*B*S, D] bf16 — the four taps after the final LayerNorm, tap-major — , B, S)."""
        B, _, Hi, Wi = images.shape
        assert Hi % PATCH == 0 and Wi % PATCH == 0, "image sides must be multiples of 14"
        D, H = self.embed_dim, self.num_heads
        hd = D // H
        w0, h0 = Hi // PATCH, Wi // PATCH
        npatch = w0 * h0
        S = npatch + 1
        wp, bp, kpad = self._patch_weight(raw)
        cols = ops.im2col_patches(images.contiguous(), PATCH, kpad)
        patch = ops.gemm(cols, wp, bias=bp)
        del cols
        x = ops.clip_embed(patch, self.cls_token.view(-1), self._pos_table(w0, h0), B, npatch)
        del patch
        taps = torch.empty((4 * B * S, D), dtype=BF16, device=x.device)
        nt = 0
        for i, blk in enumerate(self.blocks):
            pw, pb, fw, fb = blk.folded()
            h, _, _ = ops.layernorm_fwd(x, blk.norm1.weight, blk.norm1.bias, EPS)
            qkv = ops.gemm(h, blk.attn.qkv.weight, bias=blk.attn.qkv.bias)
            a, _ = ops.attn_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B, H, H, S, S, hd, hd ** -0.5, False)
            x = ops.gemm(a, pw, bias=pb, residual=x)                      # x + ls1(proj(attn))
            h, _, _ = ops.layernorm_fwd(x, blk.norm2.weight, blk.norm2.bias, EPS)
            f = ops.gemm(h, blk.mlp.fc1.weight, bias=blk.mlp.fc1.bias, act=ACT_GELU)
            x = ops.gemm(f, fw, bias=fb, residual=x)                      # x + ls2(fc2(gelu(fc1)))
            if i in self.taps:
                ops.layernorm_fwd(x, self.norm.weight, self.norm.bias, EPS, out=taps[nt * B * S:(nt + 1) * B * S])
                nt += 1
        return taps, B, S

    @torch.no_grad()
    def get_intermediate_layers(self, x, n=None, reshape=False, return_class_token=False, norm=True):
        """dinov2.py:278-303 for n == this encoder's taps, norm=True, reshape=False."""
        assert (n is None or tuple(n) == tuple(self.taps)) and norm and not reshape
        taps, B, S = self.normed_taps(_as_bf16(x, self.norm.weight.device))
        views = taps.view(4, B, S, self.embed_dim)
        outs = tuple(views[i, :, 1:] for i in range(4))
        return tuple(zip(outs, (views[i, :, 0] for i in range(4)))) if return_class_token else outs

    @torch.no_grad()
    def tap_mean(self, taps, B, S):
        """(f0 + f1 + f2 + f3) / 4 over the patch tokens (base_ola_vlm.py:355) → [B*(S-1), D]."""
        return ops.gather_sum_rows(self._tap_mean_index(B, S), 4, taps, self.embed_dim, 0.25)


def _as_bf16(x, device):
    x = x.to(device=device, non_blocking=True)
    if x.dtype == torch.float32:
        return ops.cast_bf16(x.contiguous())
    return x.to(BF16).contiguous()


class DepthAnythingV2(nn.Module):
    """The reference's `dav2_backbone` (base_ola_vlm.py:69-83): DINOv2 encoder + an (unused on this
    path) DPT head, kept so the published .pth loads with strict=True."""

    def __init__(self, encoder="vitl", features=256, out_channels=(256, 512, 1024, 1024), device=None,
                 with_depth_head=True):
        super().__init__()
        self.encoder = encoder
        self.intermediate_layer_idx = {k: list(v["taps"]) for k, v in ARCH.items()}
        self.pretrained = DinoVisionTransformer(encoder, device)
        if with_depth_head:
            self.depth_head = DPTHead(self.pretrained.embed_dim, features, tuple(out_channels), device)
        self.requires_grad_(False)

    @torch.no_grad()
    def forward(self, x):
        """dpt.py:176-180: ((patch tokens [B,N,D], cls [B,D]) × 4) for a normalised image batch."""
        return self.pretrained.get_intermediate_layers(x, self.intermediate_layer_idx[self.encoder],
                                                       return_class_token=True)

    @torch.no_grad()
    def infer_image(self, raw_image, input_size=336, is_dsg=False):
        """dpt.py:183-192 for one HxWx3 uint8 array whose sides already equal input_size (what
        _get_dav2_feats feeds it after img.resize((336, 336)); the cv2 resize is then the identity)."""
        raw = torch.as_tensor(raw_image)
        feats = self._raw_features(raw[None], input_size)
        return feats if is_dsg else feats[-1][0]

    def _raw_batch(self, raw, input_size):
        raw = torch.as_tensor(raw)
        assert raw.dim() == 4 and raw.shape[-1] == 3 and raw.dtype == torch.uint8, "uint8 [B,H,W,3] expected"
        if raw.shape[1] != input_size or raw.shape[2] != input_size or input_size % PATCH:
            raise NotImplementedError("resize the images to input_size x input_size (a multiple of 14) first, "
                                      "as _get_dav2_feats does")
        dev = self.pretrained.norm.weight.device
        ctr = torch.tensor(CENTER_NET[::-1], dtype=torch.int16, device=dev).view(1, 3, 1, 1)
        x = raw.to(dev, non_blocking=True).permute(0, 3, 1, 2).to(torch.int16) - ctr   # |x| <= 151: exact in bf16
        return x.to(BF16).contiguous()

    def _raw_features(self, raw, input_size):
        vt = self.pretrained
        taps, B, S = vt.normed_taps(self._raw_batch(raw, input_size), raw=True)
        v = taps.view(4, B, S, vt.embed_dim)
        return tuple((v[i, :, 1:], v[i, :, 0]) for i in range(4))

    @torch.no_grad()
    def dsg_targets(self, raw, input_size=336):
        """Batched target features of _get_dav2_feats: uint8 [B,336,336,3] → [B*576, D] bf16 rows
        (the mean of the four normed taps' patch tokens)."""
        vt = self.pretrained
        taps, B, S = vt.normed_taps(self._raw_batch(raw, input_size), raw=True)
        return vt.tap_mean(taps, B, S)
