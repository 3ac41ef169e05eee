"""Frozen CLIP-ConvNeXt vision tower on the GPU (BASELINE config 4).

Reference: CLIPConvNextVisionTower (model/multimodal_encoder/clip_convnext_encoder.py:61-218) — `_forward`
(:150-174) runs `trunk.stem`, the four `trunk.stages`, `norm_pre` (identity) and returns
`x.flatten(2,3).permute(0,2,1)` = [B, 576, 3072] for convnext_xxlarge at 768 px.  `trunk` is timm's ConvNeXt
(third-party, not vendored); the module tree below reproduces timm's parameter names, so the
`model.vision_tower.vision_tower.*` keys of a reference checkpoint (and an open_clip `visual.trunk.*`
state dict) load unchanged.

B200 layout: activations stay NHWC bf16 rows [B*H*W, C] from the stem to the output — the layout every 1x1
conv (= Linear) wants and the one the reference permutes into and out of inside every block.  Per block:
  vpb_dwconv7x7_nhwc (depthwise 7x7 + bias; csrc/dwconv.cu) → LayerNorm → tcgen05 GEMM (fc1 + bias + GELU
  epilogue) → tcgen05 GEMM (fc2 with `gamma` folded into its frozen weight/bias, residual add in the
  epilogue).  Four launches, no permutes, no separate layer-scale / add kernels.
The 4x4 stride-4 stem is im2col + GEMM (K = 48 padded to 64); a 2x2 stride-2 downsample is LayerNorm → four
strided row gathers into one [N/4, 4C] buffer → GEMM.  The output rows are already the [B*576, C3] the
projector consumes.  Forward only (the reference's ConvNeXt configs keep the tower frozen).
"""
from __future__ import annotations

import re
from types import SimpleNamespace

import torch
import torch.nn as nn

from .. import ops
from ..ops import ACT_GELU, BF16
from .modules import Linear, Norm, Weight, _param
from .seg_teacher import merge_plans

# timm's registered configs (timm/models/convnext.py: convnext_xxlarge passes norm_eps=1e-5, convnext_large
# keeps the LayerNorm default 1e-6) for the two towers extract_res_interp accepts (clip_convnext_encoder.py:35-38)
CONVNEXT_PRESETS = {
    "CLIP-convnext_xxlarge": dict(depths=(3, 4, 30, 3), dims=(384, 768, 1536, 3072), eps=1e-5),
    "CLIP-convnext_large": dict(depths=(3, 3, 27, 3), dims=(192, 384, 768, 1536), eps=1e-6),
}
MERGE_TAPS = ((0, 0), (1, 0), (0, 1), (1, 1))   # (dy, dx) order of seg_teacher.merge_plans


def extract_res_interp(model_name):
    """clip_convnext_encoder.py:34-60: ('CLIP-convnext_*' preset key, res, interp) from a tower name such as
    'CLIP-convnext_xxlarge-res768' / '...-res768-interp576'."""
    base = None
    for prefix in CONVNEXT_PRESETS:
        if model_name.split("/")[-1].startswith(prefix):
            base = prefix
            break
    if base is None:
        raise ValueError(f"Unknown vision tower: {model_name}")
    res = interp = None
    for part in model_name.split("-"):
        if part.startswith("res") and part[3:].isdigit():
            res = int(part[3:])
        elif part.startswith("interp") and part[6:].isdigit():
            interp = int(part[6:])
    return base, res, interp


def _read_checkpoint(path):
    """A checkpoint file, or a directory in open_clip's hub layout / this package's save layout."""
    import os

    if os.path.isdir(path):
        for name in ("open_clip_model.safetensors", "open_clip_pytorch_model.bin"):
            if os.path.exists(os.path.join(path, name)):
                path = os.path.join(path, name)
                break
        else:
            from ..train.checkpoint import load_pretrained_weights

            return load_pretrained_weights(path)
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file

        return load_file(path)
    sd = torch.load(path, map_location="cpu", weights_only=True)
    return sd.get("state_dict", sd)


OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


class OpenClipEvalTransform:
    """open_clip's inference transform for these towers — Resize(shortest side → size, bicubic) → CenterCrop →
    RGB → ToTensor → Normalize(CLIP mean/std) — with both sizes set to the tower's resolution as
    clip_convnext_encoder.py:113-114 does.  Host-side PIL work (open_clip / torchvision are not required);
    equal to the torchvision Compose on the same image (tests/test_convnext_cpu.py)."""

    def __init__(self, size, mean=OPENAI_CLIP_MEAN, std=OPENAI_CLIP_STD):
        self.size = int(size)
        self.mean = torch.tensor(mean).view(3, 1, 1)
        self.std = torch.tensor(std).view(3, 1, 1)

    def __call__(self, image):
        import numpy as np
        from PIL import Image

        w, h = image.size
        s = self.size
        short, long = (w, h) if w <= h else (h, w)
        if short != s:                                   # torchvision Resize(int): shorter side → s, aspect kept
            new_long = int(s * long / short)
            nw, nh = (s, new_long) if w <= h else (new_long, s)
            image = image.resize((nw, nh), Image.BICUBIC)
            w, h = nw, nh
        top, left = int(round((h - s) / 2.0)), int(round((w - s) / 2.0))   # torchvision CenterCrop
        image = image.crop((left, top, left + s, top + s)).convert("RGB")
        x = torch.from_numpy(np.asarray(image, dtype=np.uint8).copy()).permute(2, 0, 1).float().div_(255.0)
        return (x - self.mean) / self.std


class ProcessorWrapper:
    """multimodal_encoder/base_encoder.py:8-40: what the dataset code sees as `image_processor`."""

    def __init__(self, transform, height=378, width=378, image_mean=None):
        self._crop_size = {"height": height, "width": width}
        self._transforms = transform
        self.image_mean = list(OPENAI_CLIP_MEAN) if image_mean is None else image_mean

    @property
    def crop_size(self):
        return self._crop_size

    def preprocess(self, image, return_tensors="pt"):
        if isinstance(image, list):
            image = image[0]
        return {"pixel_values": [self._transforms(image)]}

    __call__ = preprocess


class _Mlp(nn.Module):
    def __init__(self, C, device):
        super().__init__()
        self.fc1 = Linear(C, 4 * C, True, device)
        self.fc2 = Linear(4 * C, C, True, device)


class ConvNeXtBlock(nn.Module):
    def __init__(self, C, device):
        super().__init__()
        self.gamma = _param(C, device=device)
        self.conv_dw = Weight((C, 1, 7, 7), (C,), device)
        self.norm = Norm(C, True, device)
        self.mlp = _Mlp(C, device)
        self._cache = None

    def derived(self):
        """(w49 [49, C], gamma-folded fc2 weight [C, 4C], gamma-folded fc2 bias [C]) per weight version."""
        src = (self.conv_dw.weight, self.gamma, self.mlp.fc2.weight, self.mlp.fc2.bias)
        key = tuple((t.data_ptr(), t._version) for t in src)
        if self._cache is None or self._cache[0] != key:
            self._cache = (key, fold_block(self.conv_dw.weight.detach(), self.gamma.detach(),
                                           self.mlp.fc2.weight.detach(), self.mlp.fc2.bias.detach()))
        return self._cache[1]


def fold_block(w_dw, gamma, w2, b2):
    """Frozen-weight preparation of one block (fp32 arithmetic, one bf16 rounding):
    depthwise filter [C,1,7,7] → tap-major [49, C];  gamma * (fc2(h)) = (gamma[:,None] * W2) h + gamma * b2."""
    C = w_dw.shape[0]
    w49 = w_dw.reshape(C, 49).t().contiguous()
    g = gamma.float()
    return w49, (g[:, None] * w2.float()).to(w2.dtype).contiguous(), (g * b2.float()).to(b2.dtype).contiguous()


def fold_downsample(w):
    """Conv2d(Cin, Cout, 2, stride 2) weight [Cout, Cin, 2, 2] → [Cout, 4*Cin] with column block k holding tap
    MERGE_TAPS[k], the order in which the four strided gathers lay the 2x2 neighbourhood out."""
    return torch.cat([w[:, :, dy, dx] for dy, dx in MERGE_TAPS], 1).contiguous()


class _Stage(nn.Module):
    def __init__(self, Cin, C, depth, downsample, device):
        super().__init__()
        if downsample:
            self.downsample = nn.ModuleList([Norm(Cin, True, device), Weight((C, Cin, 2, 2), (C,), device)])
        self.blocks = nn.ModuleList([ConvNeXtBlock(C, device) for _ in range(depth)])


class ConvNeXtTrunk(nn.Module):
    """timm ConvNeXt without its head: stem.{0,1}, stages.{i}.downsample.{0,1}, stages.{i}.blocks.{j}.*"""

    def __init__(self, cfg, device=None):
        super().__init__()
        dims, depths = cfg["dims"], cfg["depths"]
        self.cfg = dict(cfg)
        self.stem = nn.ModuleList([Weight((dims[0], 3, 4, 4), (dims[0],), device), Norm(dims[0], True, device)])
        self.stages = nn.ModuleList([_Stage(dims[max(i - 1, 0)], dims[i], depths[i], i > 0, device)
                                     for i in range(len(dims))])
        self.requires_grad_(False)
        self._plans = {}
        self._stem_w = None
        self._ds_w = {}

    @property
    def device(self):
        return self.stem[1].weight.device

    def _stem_weight(self):
        w = self.stem[0].weight
        key = (w.data_ptr(), w._version)
        if self._stem_w is None or self._stem_w[0] != key:
            K = w[0].numel()
            kpad = (K + 63) // 64 * 64
            wp = torch.zeros((w.shape[0], kpad), dtype=BF16, device=w.device)
            wp[:, :K] = w.detach().reshape(w.shape[0], K)
            self._stem_w = (key, wp, kpad)
        return self._stem_w[1], self._stem_w[2]

    def _downsample_weight(self, i):
        w = self.stages[i].downsample[1].weight
        key = (w.data_ptr(), w._version)
        if i not in self._ds_w or self._ds_w[i][0] != key:
            self._ds_w[i] = (key, fold_downsample(w.detach()))
        return self._ds_w[i][1]

    def _merge_plan(self, B, H, W):
        key = (B, H, W)
        if key not in self._plans:
            idx, H2, W2 = merge_plans(B, H, W)
            self._plans[key] = ([i.to(self.device) for i in idx], H2, W2)
        return self._plans[key]

    @torch.no_grad()
    def forward_rows(self, images, return_stages=False):
        """images [B,3,H,W] (H, W multiples of 32) → (last-stage NHWC rows [B*(H/32)*(W/32), C3] bf16, h, w)."""
        eps = self.cfg["eps"]
        x = images.to(device=self.device, non_blocking=True).contiguous()
        x = ops.cast_bf16(x) if x.dtype == torch.float32 else x.to(BF16)
        B, _, Hi, Wi = x.shape
        assert Hi % 32 == 0 and Wi % 32 == 0, "ConvNeXt tower: image sides must be multiples of 32"
        H, W = Hi // 4, Wi // 4
        wp, kpad = self._stem_weight()
        cols = ops.im2col_patches(x, 4, kpad)
        x = ops.gemm(cols, wp, bias=self.stem[0].bias)
        del cols
        x, _, _ = ops.layernorm_fwd(x, self.stem[1].weight, self.stem[1].bias, eps)
        stages = []
        for i, stage in enumerate(self.stages):
            if i > 0:
                ln, conv = stage.downsample
                C = x.shape[1]
                h, _, _ = ops.layernorm_fwd(x, ln.weight, ln.bias, eps)
                idx, H2, W2 = self._merge_plan(B, H, W)
                cat = torch.empty((B * H2 * W2, 4 * C), dtype=BF16, device=x.device)
                for k, ix in enumerate(idx):
                    ops.gather_rows(ix, [h], C, out=cat[:, k * C:(k + 1) * C])
                x = ops.gemm(cat, self._downsample_weight(i), bias=conv.bias)
                del cat, h
                H, W = H2, W2
            C = x.shape[1]
            for blk in stage.blocks:
                w49, w2, b2 = blk.derived()
                h = ops.dwconv7x7(x, w49, blk.conv_dw.bias, B, H, W, C)
                h, _, _ = ops.layernorm_fwd(h, blk.norm.weight, blk.norm.bias, eps)
                f = ops.gemm(h, blk.mlp.fc1.weight, bias=blk.mlp.fc1.bias, act=ACT_GELU)
                x = ops.gemm(f, w2, bias=b2, residual=x)
                del f, h
            if return_stages:
                stages.append((x, H, W))
        return (x, H, W, stages) if return_stages else (x, H, W)


class CLIPConvNextVisionTower(nn.Module):
    """Drop-in for the reference class of the same name (clip_convnext_encoder.py:61): same name parsing,
    `vision_tower` = the timm trunk, `hidden_size` / `image_size` / `num_patches[_per_side]` properties,
    forward(images) → image features.  Here forward returns the B-major rows [B*num_patches, hidden_size]
    (the layout encode_images feeds to the projector), like this package's CLIPVisionTower."""

    def __init__(self, vision_tower, args=None, delay_load=False, device=None, cfg=None):
        super().__init__()
        try:
            base, res, interp = extract_res_interp(vision_tower)
        except ValueError:
            if cfg is None:
                raise
            base, res, interp = None, None, None
        if cfg is not None and res is None:
            res = cfg.get("image_size")
        self.vision_tower_name = vision_tower
        self.ckpt_path = vision_tower.split("-res")[0]          # clip_convnext_encoder.py:77
        self.is_multi_stage = "multi-stage" in vision_tower
        if self.is_multi_stage:
            raise NotImplementedError("multi-stage ConvNeXt features are dead code in the reference (:165-172)")
        self._image_size = res if res is not None else 768
        self._interp_size = interp
        self._reduction = 32
        self.select_layer = getattr(args, "mm_vision_select_layer", -2)
        self.select_feature = getattr(args, "mm_vision_select_feature", "patch")
        self.cfg = dict(CONVNEXT_PRESETS[base] if cfg is None else cfg)
        self.vision_tower = ConvNeXtTrunk(self.cfg, device)
        self._hidden_size = self.cfg["dims"][-1]
        self.vision_model = "convnext"                          # clip_convnext_encoder.py:110
        # clip_convnext_encoder.py:113-115
        self.image_processor = ProcessorWrapper(OpenClipEvalTransform(self._image_size), height=self._image_size,
                                                width=self._image_size)
        self.is_loaded = True

    def load_model(self, device_map=None, path=None):
        """clip_convnext_encoder.py:104-123.  Reads a local open_clip / timm checkpoint directory or file
        (keys `visual.trunk.*`, `trunk.*` or bare timm names); without one this is a no-op (tests and
        benchmarks initialise the modules directly — there is no network for the hf-hub download)."""
        import os

        path = path or getattr(self, "ckpt_path", None)
        if not path or not os.path.exists(str(path)):
            return
        sd = _read_checkpoint(str(path))
        own = self.vision_tower.state_dict()
        picked = {}
        for k, v in sd.items():
            k2 = re.sub(r"^(module\.)?(visual\.)?(trunk\.)?", "", k)
            if k2 in own:
                picked[k2] = v
        missing = [k for k in own if k not in picked]
        if missing:
            raise KeyError(f"{path}: ConvNeXt trunk weights missing {missing[:4]} (+{max(0, len(missing) - 4)} more)")
        self.vision_tower.load_state_dict(picked)
        self.vision_tower.requires_grad_(False)
        self.is_loaded = True

    @property
    def config(self):
        return SimpleNamespace(hidden_size=self._hidden_size, image_size=self._image_size, patch_size=self._reduction)

    @property
    def hidden_size(self):
        return self._hidden_size

    @property
    def image_size(self):
        return self._image_size

    @property
    def num_patches_per_side(self):
        return self._image_size // self._reduction if self._interp_size is None else int(self._interp_size ** 0.5)

    @property
    def num_patches(self):
        return (self._image_size // self._reduction) ** 2 if self._interp_size is None else self._interp_size

    @property
    def dummy_feature(self):
        """base_encoder.py:72-74"""
        return torch.zeros(1, self.hidden_size, device=self.device, dtype=self.dtype)

    @property
    def dtype(self):
        return self.vision_tower.stem[1].weight.dtype

    @property
    def device(self):
        return self.vision_tower.device

    @torch.no_grad()
    def forward(self, images):
        if isinstance(images, (list, tuple)):
            return [self.forward(im.unsqueeze(0)) for im in images]   # base_encoder.py:62-67
        x, _, _ = self.vision_tower.forward_rows(images)
        return x
