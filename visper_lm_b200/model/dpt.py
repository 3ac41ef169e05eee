"""Frozen DPT depth decoder behind the `depth_preds` output field.

Reference: ola_vlm/model/aux_heads/da_v2_head.py — DPTHead (:181-293), DAv2_Head (:296-321), the
scratch / fusion blocks (:10-135); called under no_grad from base_ola_vlm.py:462-470, followed by a
per-image min-max normalisation.  Parameters keep the reference names (`da_v2_head.depth_head.*`,
torch conv layouts) so a DepthAnything-V2 head checkpoint loads unchanged; the GEMM-ready copies
([Cout, (ky,kx,cin)] for 3x3 convs, [(ky,kx,cout), cin] for the transposed convs) are derived once
per weight version — a one-time layout change of frozen weights, not step work.

Everything runs NHWC in bf16 through the C ABI: conv = im2col (conv.cu) + tcgen05 GEMM with the
bias / ReLU / skip-add in its epilogue, ConvTranspose2d(k = stride) = GEMM + pixel shuffle,
bilinear(align_corners=True), a 32→1 dot-product kernel and the min-max kernel.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ..ops import ACT_NONE, ACT_RELU, BF16


def _p(*shape, device=None):
    return nn.Parameter(torch.empty(*shape, dtype=BF16, device=device), requires_grad=False)


class Conv(nn.Module):
    def __init__(self, cin, cout, k, bias=True, device=None, transposed=False):
        super().__init__()
        self.cin, self.cout, self.k, self.transposed = cin, cout, k, transposed
        self.weight = _p(cin, cout, k, k, device=device) if transposed else _p(cout, cin, k, k, device=device)
        self.bias = _p(cout, device=device) if bias else None
        self._gemm_w = None
        self._ver = None

    def gemm_weight(self):
        """[N, K] K-major operand for the GEMM (cached per weight version)."""
        ver = (self.weight._version, self.weight.data_ptr())
        if self._gemm_w is None or self._ver != ver:
            w = self.weight.detach()
            if self.transposed:   # [cin, cout, ky, kx] → [(ky,kx,cout), cin]
                g = w.permute(2, 3, 1, 0).reshape(self.k * self.k * self.cout, self.cin)
            else:                 # [cout, cin, ky, kx] → [cout, (ky,kx,cin)]
                g = w.permute(0, 2, 3, 1).reshape(self.cout, self.k * self.k * self.cin)
            self._gemm_w, self._ver = g.contiguous(), ver
        return self._gemm_w


class _RCU(nn.Module):  # ResidualConvUnit (da_v2_head.py:34-75), bn=False
    def __init__(self, f, device):
        super().__init__()
        self.conv1 = Conv(f, f, 3, True, device)
        self.conv2 = Conv(f, f, 3, True, device)


class _Fusion(nn.Module):  # FeatureFusionBlock (da_v2_head.py:78-135)
    def __init__(self, f, device):
        super().__init__()
        self.out_conv = Conv(f, f, 1, True, device)
        self.resConfUnit1 = _RCU(f, device)
        self.resConfUnit2 = _RCU(f, device)


class _Scratch(nn.Module):
    def __init__(self, out_channels, f, device):
        super().__init__()
        for i, c in enumerate(out_channels):
            setattr(self, f"layer{i + 1}_rn", Conv(c, f, 3, False, device))
        for i in range(1, 5):
            setattr(self, f"refinenet{i}", _Fusion(f, device))
        self.output_conv1 = Conv(f, f // 2, 3, True, device)
        self.output_conv2 = nn.ModuleDict({"0": Conv(f // 2, 32, 3, True, device), "2": Conv(32, 1, 1, True, device)})


class DPTHead(nn.Module):
    def __init__(self, in_channels=1024, features=256, out_channels=(256, 512, 1024, 1024), device=None):
        super().__init__()
        self.projects = nn.ModuleList([Conv(in_channels, c, 1, True, device) for c in out_channels])
        self.resize_layers = nn.ModuleDict({
            "0": Conv(out_channels[0], out_channels[0], 4, True, device, transposed=True),
            "1": Conv(out_channels[1], out_channels[1], 2, True, device, transposed=True),
            "3": Conv(out_channels[3], out_channels[3], 3, True, device)})
        self.scratch = _Scratch(out_channels, features, device)


MAX_IM2COL_BYTES = 1 << 30  # im2col scratch is bounded by splitting the batch


class DAv2_Head(nn.Module):
    """forward(features) → relu(depth) [B,336,336] fp32, features = 4 × [B*576, 1024] bf16 rows."""

    patch = 24

    def __init__(self, device=None):
        super().__init__()
        self.depth_head = DPTHead(device=device)

    # ---- building blocks (NHWC rows) ---------------------------------------------------------
    @staticmethod
    def _conv3(x, B, H, W, conv, stride=1, relu_in=False, act=ACT_NONE, residual=None):
        C = conv.cin
        Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
        per = Ho * Wo * 9 * C * 2
        nb = max(1, min(B, MAX_IM2COL_BYTES // per))
        if nb >= B:
            col, _, _ = ops.im2col3x3(x, B, H, W, C, stride, relu_in)
            return ops.gemm(col, conv.gemm_weight(), bias=conv.bias, act=act, residual=residual), Ho, Wo
        out = torch.empty((B * Ho * Wo, conv.cout), dtype=BF16, device=x.device)
        for b0 in range(0, B, nb):
            b1 = min(B, b0 + nb)
            col, _, _ = ops.im2col3x3(x[b0 * H * W:b1 * H * W], b1 - b0, H, W, C, stride, relu_in)
            ops.gemm(col, conv.gemm_weight(), bias=conv.bias, act=act,
                     residual=None if residual is None else residual[b0 * Ho * Wo:b1 * Ho * Wo],
                     out=out[b0 * Ho * Wo:b1 * Ho * Wo])
            del col
        return out, Ho, Wo

    def _rcu(self, x, B, H, W, rcu):
        # relu(x) → conv1 → relu → conv2 → + x : first ReLU on the im2col load, second in conv1's
        # GEMM epilogue, the skip-add in conv2's
        t, _, _ = self._conv3(x, B, H, W, rcu.conv1, relu_in=True, act=ACT_RELU)
        y, _, _ = self._conv3(t, B, H, W, rcu.conv2, residual=x)
        return y

    def _fusion(self, blk, B, H, W, x0, x1=None, size=None):
        out = x0
        if x1 is not None:
            out = ops.axpby(out, self._rcu(x1, B, H, W, blk.resConfUnit1))
        out = self._rcu(out, B, H, W, blk.resConfUnit2)
        Ho, Wo = (2 * H, 2 * W) if size is None else size
        C = blk.out_conv.cin
        out = ops.bilinear(out, B, H, W, Ho, Wo, C)
        return ops.gemm(out, blk.out_conv.gemm_weight(), bias=blk.out_conv.bias), Ho, Wo

    @torch.no_grad()
    def forward(self, features):
        dh, sc, P = self.depth_head, self.depth_head.scratch, self.patch
        B = features[0].shape[0] // (P * P)
        outs = []
        for i, x in enumerate(features):
            pj = dh.projects[i]
            x = ops.gemm(x.contiguous(), pj.gemm_weight(), bias=pj.bias)
            if i in (0, 1):
                rs = dh.resize_layers[str(i)]
                y = ops.gemm(x, rs.gemm_weight())
                x = ops.pixel_shuffle(y, rs.bias, B, P, P, rs.cout, rs.k)
                outs.append((x, P * rs.k, P * rs.k))
            elif i == 3:
                x, Ho, Wo = self._conv3(x, B, P, P, dh.resize_layers["3"], stride=2)
                outs.append((x, Ho, Wo))
            else:
                outs.append((x, P, P))
        ls = []
        for i, (x, H, W) in enumerate(outs):
            y, _, _ = self._conv3(x, B, H, W, getattr(sc, f"layer{i + 1}_rn"))
            ls.append((y, H, W))
        (l1, H1, W1), (l2, H2, W2), (l3, H3, W3), (l4, H4, W4) = ls
        p4, _, _ = self._fusion(sc.refinenet4, B, H4, W4, l4, size=(H3, W3))
        p3, _, _ = self._fusion(sc.refinenet3, B, H3, W3, p4, l3, size=(H2, W2))
        p2, _, _ = self._fusion(sc.refinenet2, B, H2, W2, p3, l2, size=(H1, W1))
        p1, Hf, Wf = self._fusion(sc.refinenet1, B, H1, W1, p2, l1)
        out, _, _ = self._conv3(p1, B, Hf, Wf, sc.output_conv1)
        S = P * 14
        out = ops.bilinear(out, B, Hf, Wf, S, S, sc.output_conv1.cout)
        out, _, _ = self._conv3(out, B, S, S, sc.output_conv2["0"], act=ACT_RELU)
        last = sc.output_conv2["2"]
        depth = ops.conv1x1_to1(out, last.weight.view(-1), last.bias, relu=True)  # ReLU ∘ ReLU ∘ F.relu
        return depth.view(B, S, S)

    @torch.no_grad()
    def normalized(self, features):
        """depth_pred of base_ola_vlm.py:462-470 (min-max normalised per image)."""
        return ops.minmax_normalize(self.forward(features))
