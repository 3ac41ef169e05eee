"""Frozen generation teacher on the GPU, batched: the unCLIP image encoder whose
image_embeds are the `gen` distillation targets.

Reference call site: _get_gen_feats (language_model/base_ola_vlm.py:323-333) —
`self.pipe.image_encoder(clip_ims).image_embeds`, one image at a time.  `pipe` is diffusers'
StableUnCLIPImg2ImgPipeline; its image_encoder is transformers' CLIPVisionModelWithProjection
(OpenCLIP ViT-H/14: 1280 wide, 32 layers, 16 heads of 80, MLP 5120 erf-GELU, 224 px, projection 1024).
That model is third-party code, not part of the reference tree; this class restates it with the HF
parameter names (`vision_model.*`, `visual_projection.weight`) so the published weights load unchanged.

All compute runs through the C ABI: im2col + tcgen05 GEMM patch conv, LayerNorm, fused-QKV GEMM,
flash attention, out-proj / fc2 GEMMs with the residual in the epilogue, fc1 GEMM with GELU in the
epilogue, cls-row gather, LayerNorm, projection GEMM.  head_dim 80 has no attention kernel of its own:
the frozen Q/K/V rows of each head are zero-padded to 96 once per weight version (and out_proj's
columns to match), which leaves every score and every output unchanged (the softmax scale stays
80^-0.5) and runs on the head_dim-96 tcgen05 kernel.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
import torch.nn as nn

from .. import ops
from ..ops import ACT_GELU, ACT_QUICK_GELU, BF16
from .modules import CLIPVisionTransformer, Linear

_ATTN_HD = (32, 64, 96, 128)  # head dims the attention kernels take


class CLIPVisionModelWithProjection(nn.Module):
    def __init__(self, config: dict, device=None):
        super().__init__()
        self.config = SimpleNamespace(**config)
        self.vision_model = CLIPVisionTransformer(config, device)
        self.visual_projection = Linear(config["hidden_size"], config["projection_dim"], False, device)
        self.requires_grad_(False)
        self._patch_w = None
        self._padded = {}
        self._cls_idx = {}

    @property
    def device(self):
        return self.visual_projection.weight.device

    # ---- derived, cached ----------------------------------------------------------------------
    def _patch_weight(self):
        w = self.vision_model.embeddings.patch_embedding.weight
        key = (w.data_ptr(), w._version)
        if self._patch_w is None or self._patch_w[0] != key:
            K = w[0].numel()
            kpad = (K + 63) // 64 * 64
            wp = torch.zeros((w.shape[0], kpad), dtype=BF16, device=w.device)
            wp[:, :K] = w.detach().reshape(w.shape[0], K)
            self._patch_w = (key, wp, kpad)
        return self._patch_w[1], self._patch_w[2]

    def _layer_weights(self, li, heads, hd, hp):
        """(qkv_w [3*heads*hp, D], qkv_b, out_w [D, heads*hp]) with each head padded hd → hp by zeros."""
        a = self.vision_model.encoder.layers[li].self_attn
        src = (a.q_proj.weight, a.k_proj.weight, a.v_proj.weight, a.q_proj.bias, a.k_proj.bias, a.v_proj.bias,
               a.out_proj.weight)
        key = tuple((t.data_ptr(), t._version) for t in src)
        hit = self._padded.get(li)
        if hit is None or hit[0] != key:
            D = a.q_proj.weight.shape[1]
            with torch.no_grad():
                w = torch.zeros((3, heads, hp, D), dtype=BF16, device=self.device)
                b = torch.zeros((3, heads, hp), dtype=BF16, device=self.device)
                for j, (pw, pb) in enumerate(((a.q_proj.weight, a.q_proj.bias), (a.k_proj.weight, a.k_proj.bias),
                                              (a.v_proj.weight, a.v_proj.bias))):
                    w[j, :, :hd] = pw.detach().view(heads, hd, D)
                    b[j, :, :hd] = pb.detach().view(heads, hd)
                o = torch.zeros((D, heads, hp), dtype=BF16, device=self.device)
                o[:, :, :hd] = a.out_proj.weight.detach().view(D, heads, hd)
            hit = (key, (w.view(3 * heads * hp, D), b.view(-1), o.view(D, heads * hp)))
            self._padded[li] = hit
        return hit[1]

    # ---- compute ------------------------------------------------------------------------------
    @torch.no_grad()
    def image_embeds(self, pixel_values):
        """pixel_values [B,3,H,W] (CLIPImageProcessor output, any float dtype) → [B, projection_dim] bf16."""
        vm, cfg = self.vision_model, self.config
        x = pixel_values.to(device=self.device, non_blocking=True).contiguous()
        x = ops.cast_bf16(x) if x.dtype == torch.float32 else x.to(BF16)
        B = x.shape[0]
        D, heads = cfg.hidden_size, cfg.num_attention_heads
        hd = D // heads
        hp = min(h for h in _ATTN_HD if h >= hd)
        act = ACT_GELU if cfg.hidden_act == "gelu" else ACT_QUICK_GELU
        eps = getattr(cfg, "layer_norm_eps", 1e-5)
        npatch = (cfg.image_size // cfg.patch_size) ** 2
        assert x.shape[2] == x.shape[3] == cfg.image_size, "pixel_values must be image_size x image_size"
        S = npatch + 1
        wp, kpad = self._patch_weight()
        cols = ops.im2col_patches(x, cfg.patch_size, kpad)
        patch = ops.gemm(cols, wp)
        del cols
        emb = ops.clip_embed(patch, vm.embeddings.class_embedding, vm.embeddings.position_embedding.weight, B, npatch)
        x, _, _ = ops.layernorm_fwd(emb, vm.pre_layrnorm.weight, vm.pre_layrnorm.bias, eps)
        W = heads * hp
        for li, L in enumerate(vm.encoder.layers):
            qkv_w, qkv_b, out_w = self._layer_weights(li, heads, hd, hp)
            h, _, _ = ops.layernorm_fwd(x, L.layer_norm1.weight, L.layer_norm1.bias, eps)
            qkv = ops.gemm(h, qkv_w, bias=qkv_b)
            a, _ = ops.attn_fwd(qkv[:, :W], qkv[:, W:2 * W], qkv[:, 2 * W:], B, heads, heads, S, S, hp,
                                hd ** -0.5, False)
            x = ops.gemm(a, out_w, bias=L.self_attn.out_proj.bias, residual=x)
            h, _, _ = ops.layernorm_fwd(x, L.layer_norm2.weight, L.layer_norm2.bias, eps)
            f = ops.gemm(h, L.mlp.fc1.weight, bias=L.mlp.fc1.bias, act=act)
            x = ops.gemm(f, L.mlp.fc2.weight, bias=L.mlp.fc2.bias, residual=x)
        if (B, S) not in self._cls_idx:
            self._cls_idx[(B, S)] = (torch.arange(B, dtype=torch.int32) * S).to(self.device)
        cls = ops.gather_rows(self._cls_idx[(B, S)], [x], D)
        pooled, _, _ = ops.layernorm_fwd(cls, vm.post_layernorm.weight, vm.post_layernorm.bias, eps)
        return ops.gemm(pooled, self.visual_projection.weight)

    def forward(self, pixel_values):
        """HF-shaped call: `.image_embeds` of the returned object."""
        return SimpleNamespace(image_embeds=self.image_embeds(pixel_values))


UNCLIP_VIT_H = dict(hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=16,
                    patch_size=14, image_size=224, projection_dim=1024, hidden_act="gelu", layer_norm_eps=1e-5)
