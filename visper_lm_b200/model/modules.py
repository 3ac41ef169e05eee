"""Parameter containers that reproduce the reference's module tree — and therefore its state-dict
ABI — without any torch compute: every forward goes through visper_lm_b200.autograd.

Names mirror HF LlamaModel / Phi3Model / CLIPVisionModel and the reference's
ola_vlm/model/{ola_arch.py, multimodal_projector/{builder,resampler}.py, aux_heads/*}.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
import torch.nn as nn

from .. import autograd as A
from .. import ops
from ..ops import ACT_GELU, ACT_NONE, ACT_QUICK_GELU, ACT_RELU, BF16


def _param(*shape, device=None, dtype=BF16):
    return nn.Parameter(torch.empty(*shape, dtype=dtype, device=device))


class Weight(nn.Module):
    """nn.Linear / nn.LayerNorm / nn.Embedding stand-in: holds `weight` (+ `bias`)."""

    def __init__(self, shape, bias_shape=None, device=None):
        super().__init__()
        self.weight = _param(*shape, device=device)
        if bias_shape is not None:
            self.bias = _param(*bias_shape, device=device)
        else:
            self.bias = None


def Linear(i, o, bias=True, device=None):
    return Weight((o, i), (o,) if bias else None, device)


def Norm(d, bias=True, device=None):
    return Weight((d,), (d,) if bias else None, device)


class Seq(nn.Module):
    """nn.Sequential-compatible naming ("0", "2", ...) for parameter-holding children only."""

    def __init__(self, **children):
        super().__init__()
        for k, v in children.items():
            self.add_module(k.lstrip("_"), v)

    def __getitem__(self, i):
        return getattr(self, str(i))


class FusedRows:
    """Keeps several [rows_i, K] parameters as row-slices of ONE contiguous buffer so a single GEMM
    serves them (q|k|v, gate|up) while each keeps its reference name / requires_grad.

    Storage has ONE owner.  Before an optimizer exists, this object owns a concatenated buffer and the
    parameters are views of it.  Once Zero2Optimizer has moved the parameters into its flat buffer (it
    lays a fused group out adjacently for exactly this purpose) the group is ADOPTED in place: `fused`
    becomes a view of that region, nothing is copied, and AdamW's writes are what the next GEMM reads."""

    def __init__(self, params):
        self.params = list(params)
        self.fused = None

    def _adjacent(self):
        ps = self.params
        ptr = ps[0].data_ptr()
        for p in ps:
            if not p.is_contiguous() or p.data_ptr() != ptr or p.dtype != ps[0].dtype:
                return False
            ptr += p.numel() * p.element_size()
        return ps[0].untyped_storage().data_ptr() == ps[-1].untyped_storage().data_ptr()

    def get(self):
        ps = self.params
        if self.fused is not None and self.fused.data_ptr() == ps[0].data_ptr() and self._adjacent():
            return self.fused
        with torch.no_grad():
            if self._adjacent():
                rows = sum(p.shape[0] for p in ps)
                inner = tuple(ps[0].shape[1:])
                stride = ps[0].data.stride()
                self.fused = torch.as_strided(ps[0].data, (rows,) + inner, stride, ps[0].data.storage_offset())
                return self.fused
            if any(getattr(p, "_vpb_flat_owner", False) for p in ps):
                raise RuntimeError("FusedRows: parameters live in an optimizer's flat buffer but are not adjacent "
                                   "there; re-copying them would detach them from the optimizer")
            fused = torch.cat([p.data.reshape(p.shape[0], -1) for p in ps], 0).contiguous()
            r = 0
            for p in ps:
                n = p.shape[0]
                p.data = fused[r:r + n].view(p.shape)
                r += n
        self.fused = fused
        return self.fused


class FrozenTranspose:
    """K-major copy Wᵀ of a FROZEN weight for the dgrad GEMM (dY·W): trades HBM capacity (the PT stage
    freezes the whole LLM) for the faster K-major B operand.  Trainable weights use
    the MN-major descriptor path instead (no copy to keep in sync)."""

    def __init__(self):
        self.key = None
        self.wt = None

    def get(self, w):
        if w.requires_grad:
            return None
        key = (w.data_ptr(), w._version, tuple(w.shape))
        if self.key != key:
            with torch.no_grad():
                self.wt = ops.transpose(w.detach())
            self.key = key
        return self.wt


# ------------------------------------------------------------------------------------------------ decoder
class Attention(nn.Module):
    def __init__(self, cfg, device):
        super().__init__()
        D, H, KVH = cfg.hidden_size, cfg.num_attention_heads, cfg.num_key_value_heads
        hd = D // H
        if cfg.family == "phi3":
            self.qkv_proj = Linear(D, (H + 2 * KVH) * hd, False, device)
        else:
            self.q_proj = Linear(D, H * hd, False, device)
            self.k_proj = Linear(D, KVH * hd, False, device)
            self.v_proj = Linear(D, KVH * hd, False, device)
        self.o_proj = Linear(H * hd, D, False, device)


class MLP(nn.Module):
    def __init__(self, cfg, device):
        super().__init__()
        D, F = cfg.hidden_size, cfg.intermediate_size
        if cfg.family == "phi3":
            self.gate_up_proj = Linear(D, 2 * F, False, device)
        else:
            self.gate_proj = Linear(D, F, False, device)
            self.up_proj = Linear(D, F, False, device)
        self.down_proj = Linear(F, D, False, device)


class DecoderLayer(nn.Module):
    def __init__(self, cfg, device):
        super().__init__()
        self.family = cfg.family
        self.self_attn = Attention(cfg, device)
        self.mlp = MLP(cfg, device)
        self.input_layernorm = Norm(cfg.hidden_size, False, device)
        self.post_attention_layernorm = Norm(cfg.hidden_size, False, device)
        if self.family != "phi3":
            a, m = self.self_attn, self.mlp
            self._qkv = FusedRows([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight])
            self._gu = FusedRows([m.gate_proj.weight, m.up_proj.weight])
        self._grad_sink = None   # set by Zero2Optimizer (trainer.create_optimizer): wgrads go straight to its buffer
        self._t = [FrozenTranspose() for _ in range(4)]
        # measured on B200: no gain inside the (power-capped) step, costs +14 GB → off by default
        self.use_frozen_transposes = False

    def sink_params(self):
        """{key: [parameters]} of the four weight gradients the hand-written backward produces."""
        a, m = self.self_attn, self.mlp
        if self.family == "phi3":
            return {"qkv": [a.qkv_proj.weight], "o": [a.o_proj.weight], "gu": [m.gate_up_proj.weight],
                    "d": [m.down_proj.weight]}
        return {"qkv": [a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], "o": [a.o_proj.weight],
                "gu": [m.gate_proj.weight, m.up_proj.weight], "d": [m.down_proj.weight]}

    def run(self, x, meta_base):
        a, m = self.self_attn, self.mlp
        if self.family == "phi3":
            wqkv, wgu = a.qkv_proj.weight, m.gate_up_proj.weight
            wq, wk, wv, wg, wu = wqkv, None, None, wgu, None
        else:
            wqkv, wgu = self._qkv.get(), self._gu.get()
            wq, wk, wv = a.q_proj.weight, a.k_proj.weight, a.v_proj.weight
            wg, wu = m.gate_proj.weight, m.up_proj.weight
        meta = SimpleNamespace(**vars(meta_base))
        meta.wqkv, meta.wgu = wqkv.detach(), wgu.detach()
        meta.grad_sink = self._grad_sink
        meta.wqkvT = meta.woT = meta.wguT = meta.wdT = None
        if self.use_frozen_transposes and torch.is_grad_enabled():
            frozen_qkv = not (wq.requires_grad or (wk is not None and (wk.requires_grad or wv.requires_grad)))
            frozen_gu = not (wg.requires_grad or (wu is not None and wu.requires_grad))
            if frozen_qkv:
                meta.wqkvT = self._t[0].get(meta.wqkv)
            meta.woT = self._t[1].get(a.o_proj.weight)
            if frozen_gu:
                meta.wguT = self._t[2].get(meta.wgu)
            meta.wdT = self._t[3].get(m.down_proj.weight)
        return A.DecoderLayerFn.apply(x, self.input_layernorm.weight, wq, wk, wv, a.o_proj.weight,
                                      self.post_attention_layernorm.weight, wg, wu, m.down_proj.weight,
                                      meta)


# ------------------------------------------------------------------------------------------------ CLIP tower
class CLIPAttention(nn.Module):
    def __init__(self, D, device):
        super().__init__()
        self.k_proj = Linear(D, D, True, device)
        self.v_proj = Linear(D, D, True, device)
        self.q_proj = Linear(D, D, True, device)
        self.out_proj = Linear(D, D, True, device)


class CLIPMLP(nn.Module):
    def __init__(self, D, Fi, device):
        super().__init__()
        self.fc1 = Linear(D, Fi, True, device)
        self.fc2 = Linear(Fi, D, True, device)


class CLIPEncoderLayer(nn.Module):
    def __init__(self, D, Fi, device):
        super().__init__()
        self.self_attn = CLIPAttention(D, device)
        self.layer_norm1 = Norm(D, True, device)
        self.mlp = CLIPMLP(D, Fi, device)
        self.layer_norm2 = Norm(D, True, device)
        s = self.self_attn
        self._qkv_w = FusedRows([s.q_proj.weight, s.k_proj.weight, s.v_proj.weight])
        self._qkv_b = FusedRows([s.q_proj.bias, s.k_proj.bias, s.v_proj.bias])


class CLIPEmbeddings(nn.Module):
    def __init__(self, D, patch, npos, device):
        super().__init__()
        self.class_embedding = _param(D, device=device)
        self.patch_embedding = Weight((D, 3, patch, patch), None, device)
        self.position_embedding = Weight((npos, D), None, device)


class CLIPEncoder(nn.Module):
    def __init__(self, D, Fi, L, device):
        super().__init__()
        self.layers = nn.ModuleList([CLIPEncoderLayer(D, Fi, device) for _ in range(L)])


class CLIPVisionTransformer(nn.Module):
    def __init__(self, vc, device):
        super().__init__()
        D = vc["hidden_size"]
        npos = (vc["image_size"] // vc["patch_size"]) ** 2 + 1
        self.embeddings = CLIPEmbeddings(D, vc["patch_size"], npos, device)
        self.pre_layrnorm = Norm(D, True, device)
        self.encoder = CLIPEncoder(D, vc["intermediate_size"], vc["num_hidden_layers"], device)
        self.post_layernorm = Norm(D, True, device)


class CLIPVisionModel(nn.Module):
    def __init__(self, vc, device):
        super().__init__()
        self.vision_model = CLIPVisionTransformer(vc, device)
        self.config = SimpleNamespace(**vc)


class CLIPVisionTower(nn.Module):
    """Frozen CLIP ViT tower: forward-only, hidden_states[select_layer=-2][:, 1:]
    (model/multimodal_encoder/clip_encoder.py:32-59).  The unused last encoder layer is skipped."""

    def __init__(self, vision_cfg, select_layer=-2, select_feature="patch", device=None):
        super().__init__()
        self.is_loaded = True
        self.select_layer = select_layer
        self.select_feature = select_feature
        self.vision_tower = CLIPVisionModel(vision_cfg, device)
        self.vision_tower.requires_grad_(False)
        self.image_processor = None
        self._patch_w = None
        self._drop_cls = {}

    def load_model(self, device_map=None, path=None):
        """clip_encoder.py:27-35: load the frozen tower's weights.  `path` (or `vision_tower_name`) is a
        local HF CLIPVisionModel / CLIPModel directory (model.safetensors, its shards or
        pytorch_model.bin); only the `vision_model.*` tensors are read.  Without a path this is a no-op
        (the modules already exist; tests and benchmarks initialise them directly)."""
        import os

        path = path or getattr(self, "vision_tower_name", None)
        if not path or not os.path.isdir(str(path)):
            return
        from ..train.checkpoint import load_pretrained_weights

        sd = load_pretrained_weights(str(path))
        own = self.vision_tower.state_dict()
        picked = {k: v for k, v in sd.items() if k in own}
        missing = [k for k in own if k not in picked]
        if missing:
            raise KeyError(f"{path}: CLIP vision weights missing {missing[:4]} (+{max(0, len(missing) - 4)} more)")
        self.vision_tower.load_state_dict(picked)
        self.vision_tower.requires_grad_(False)
        self._patch_w = None
        self.is_loaded = True

    @property
    def config(self):
        return self.vision_tower.config

    @property
    def hidden_size(self):
        return self.config.hidden_size

    @property
    def num_patches_per_side(self):
        return self.config.image_size // self.config.patch_size

    @property
    def num_patches(self):
        return self.num_patches_per_side ** 2

    @property
    def dtype(self):
        return self.vision_tower.vision_model.pre_layrnorm.weight.dtype

    @property
    def device(self):
        return self.vision_tower.vision_model.pre_layrnorm.weight.device

    def _padded_patch_weight(self):
        w = self.vision_tower.vision_model.embeddings.patch_embedding.weight
        K = w[0].numel()
        kpad = (K + 63) // 64 * 64
        key = (w.data_ptr(), w._version)
        if self._patch_w is None or self._patch_w[0] != key:
            wp = torch.zeros((w.shape[0], kpad), dtype=BF16, device=w.device)
            wp[:, :K] = w.detach().reshape(w.shape[0], K)
            self._patch_w = (key, wp)
        return self._patch_w[1], kpad

    @torch.no_grad()
    def forward(self, images):
        """images [B,3,H,W] (any float dtype) → [B*npatch, D] rows (B-major), bf16."""
        vm = self.vision_tower.vision_model
        cfg = self.config
        images = images.to(device=self.device, non_blocking=True).contiguous()
        if images.dtype == torch.float32:
            images = ops.cast_bf16(images)
        elif images.dtype != BF16:
            images = images.to(BF16)
        B = images.shape[0]
        D, Hh = cfg.hidden_size, cfg.num_attention_heads
        hd = D // Hh
        npatch = self.num_patches
        S = npatch + 1
        wp, kpad = self._padded_patch_weight()
        cols = ops.im2col_patches(images, cfg.patch_size, kpad)
        patch = ops.gemm(cols, wp)
        del cols
        emb = ops.clip_embed(patch, vm.embeddings.class_embedding, vm.embeddings.position_embedding.weight,
                             B, npatch)
        x, _, _ = ops.layernorm_fwd(emb, vm.pre_layrnorm.weight, vm.pre_layrnorm.bias, 1e-5)
        n_layers = len(vm.encoder.layers)
        sel = self.select_layer if self.select_layer >= 0 else n_layers + 1 + self.select_layer
        for li in range(sel):  # hidden_states[sel] is the output of layer sel-1
            L = vm.encoder.layers[li]
            h, _, _ = ops.layernorm_fwd(x, L.layer_norm1.weight, L.layer_norm1.bias, 1e-5)
            qkv = ops.gemm(h, L._qkv_w.get(), bias=L._qkv_b.get())
            a, _ = ops.attn_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B, Hh, Hh, S, S, hd,
                                hd ** -0.5, False)
            x = ops.gemm(a, L.self_attn.out_proj.weight, bias=L.self_attn.out_proj.bias, residual=x)
            h, _, _ = ops.layernorm_fwd(x, L.layer_norm2.weight, L.layer_norm2.bias, 1e-5)
            f = ops.gemm(h, L.mlp.fc1.weight, bias=L.mlp.fc1.bias, act=ACT_QUICK_GELU)
            x = ops.gemm(f, L.mlp.fc2.weight, bias=L.mlp.fc2.bias, residual=x)
        if self.select_feature != "patch":
            return x
        if B not in self._drop_cls:
            idx = torch.arange(B * S, dtype=torch.int32).view(B, S)[:, 1:].reshape(-1)
            self._drop_cls[B] = idx.to(self.device)
        return ops.gather_rows(self._drop_cls[B], [x], D)


# ------------------------------------------------------------------------------------------------ heads
class PerceiverAttention(nn.Module):
    def __init__(self, dim, dim_head, heads, device):
        super().__init__()
        inner = dim_head * heads
        self.heads, self.dim_head = heads, dim_head
        self.norm1 = Norm(dim, True, device)
        self.norm2 = Norm(dim, True, device)
        self.to_q = Linear(dim, inner, False, device)
        self.to_kv = Linear(dim, 2 * inner, False, device)
        self.to_out = Linear(inner, dim, False, device)


class TaskTokenResampler(nn.Module):
    """model/multimodal_projector/resampler.py:167-224 (depth must be 1, as every shipped config)."""

    def __init__(self, dim, depth, dim_head, heads, num_queries, embedding_dim, output_dim, ff_mult,
                 device=None):
        super().__init__()
        assert depth == 1, "only depth=1 resamplers are on the reference's training path"
        self.num_queries, self.dim, self.heads = num_queries, dim, heads
        self.proj_in = Linear(embedding_dim, dim, True, device)
        self.proj_out = Linear(dim, output_dim, True, device)
        self.norm_out = Norm(output_dim, True, device)
        inner = int(dim * ff_mult)
        ff = Seq(_0=Norm(dim, True, device), _1=Linear(dim, inner, False, device),
                 _3=Linear(inner, dim, False, device))
        self.layers = nn.ModuleList([nn.ModuleList([PerceiverAttention(dim, dim_head, heads, device), ff])])

    def run(self, state, special, plan):
        at, ff = self.layers[0][0], self.layers[0][1]
        return A.ResamplerFn.apply(
            state, special, self.proj_in.weight, self.proj_in.bias, at.norm1.weight, at.norm1.bias,
            at.norm2.weight, at.norm2.bias, at.to_q.weight, at.to_kv.weight, at.to_out.weight,
            ff[0].weight, ff[0].bias, ff[1].weight, ff[3].weight, self.proj_out.weight,
            self.proj_out.bias, self.norm_out.weight, self.norm_out.bias, plan)


def _head_resampler(proj_config, dim, llm_hidden, device):
    return TaskTokenResampler(dim=dim, depth=proj_config["depth"], dim_head=proj_config["dim_head"],
                              heads=proj_config["num_heads"], num_queries=proj_config["num_tokens"],
                              embedding_dim=llm_hidden, output_dim=proj_config["output_dim"],
                              ff_mult=proj_config["ff_mult"], device=device)


class TaskTokenGenHead(nn.Module):
    """model/aux_heads/gen_head.py:39-65"""

    def __init__(self, proj_config, llm_hidden_size, device=None):
        super().__init__()
        self.projector = _head_resampler(proj_config, proj_config["output_dim"], llm_hidden_size, device)


class OneFormerTaskTokenSegHead(nn.Module):
    """model/aux_heads/oneformer_head.py:224-258"""

    def __init__(self, proj_config, llm_hidden_size, device=None):
        super().__init__()
        self.projector = _head_resampler(proj_config, proj_config["output_dim"], llm_hidden_size, device)


class TaskTokenDepthHead(nn.Module):
    """model/aux_heads/da_v2_head.py:418-457 — resampler dim = llm hidden, then three
    Linear→ReLU→Linear MLPs (build_mlp :331-335)."""

    def __init__(self, proj_config, llm_hidden_size, use_intermediate_depth=True, device=None):
        super().__init__()
        od = proj_config["output_dim"]
        self.projector = _head_resampler(proj_config, llm_hidden_size, llm_hidden_size, device)
        self.use_intermediate_depth = use_intermediate_depth
        if use_intermediate_depth:
            for k in (1, 2, 3):
                setattr(self, f"linear_{k}", Seq(_0=Linear(od, od, True, device), _2=Linear(od, od, True, device)))

    def mlp(self, k, x):
        m = getattr(self, f"linear_{k}")
        h = A.linear(x, m[0].weight, m[0].bias, ACT_RELU)
        return A.linear(h, m[2].weight, m[2].bias, ACT_NONE)
