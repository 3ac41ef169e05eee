"""CPU tests of the CLIP-ConvNeXt tower: the oracle restatement against the golden
vectors generated from transformers' ConvNextModel (and against that model live), the plain-C depthwise
oracle against torch's conv2d, the product module's state-dict ABI (timm names), and the product's own
forward_rows — its folded weights, stem padding, merge gathers and launch order — with torch standing in
for the C-ABI calls at the same bf16 rounding points the kernels have."""
import ctypes
import subprocess
from pathlib import Path
from types import SimpleNamespace

import pytest
import torch
import torch.nn.functional as F

from parity_utils import restate

from oracle.make_golden_convnext import MINI, PREFIX, hf_model, pixels

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"
BF16 = torch.bfloat16


def test_oracle_matches_library_golden():
    fx = torch.load(GOLDEN / "convnext_mini_96.pt")
    cfg = fx["config"]
    sd = restate.convnext_seeded_state(cfg, PREFIX)
    px = pixels(fx["B"], fx["size"], fx["seed"])
    with torch.no_grad():
        st = restate.convnext_stage_features(sd, px, cfg, PREFIX)
        tower = restate.convnext_tower(sd, px, cfg, PREFIX)
    for i in range(3):
        assert torch.allclose(st[i][:, ::4, ::2, ::2], fx["stages_sub"][i], atol=3e-5)
    assert torch.allclose(st[3], fx["stages_sub"][3], atol=3e-5)
    assert all(abs(float(s.std()) - w) < 1e-4 for s, w in zip(st, fx["stage_std"]))
    # clip_convnext_encoder.py:173 — flatten(2,3).permute(0,2,1): token t = y*w + x, channels last
    assert tower.shape == (fx["B"], 9, cfg["dims"][-1])
    assert torch.equal(tower[1, 5], st[3][1, :, 1, 2])


def test_oracle_matches_library_live():
    cfg = dict(depths=(1, 2, 1, 1), dims=(64, 64, 128, 64), eps=1e-6)
    sd = restate.convnext_seeded_state(cfg, PREFIX)
    px = pixels(1, 64, 5)
    with torch.no_grad():
        want = hf_model(cfg, sd)(px, output_hidden_states=True).hidden_states[1:]
        got = restate.convnext_stage_features(sd, px, cfg, PREFIX)
    for a, b in zip(got, want):
        assert torch.allclose(a, b, atol=2e-5)


def test_c_depthwise_oracle_matches_torch(tmp_path):
    """oracle/c/dwconv_ref.c (the checker of tools/dwconv_check.cu on the GPU box) == F.conv2d(groups=C)."""
    so = tmp_path / "dwconv_ref.so"
    subprocess.run(["gcc", "-O2", "-shared", "-fPIC", str(ROOT / "oracle" / "c" / "dwconv_ref.c"), "-o", str(so)], check=True)
    fn = ctypes.CDLL(str(so)).oracle_dwconv7x7_nhwc
    fn.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int] * 4
    g = torch.Generator().manual_seed(3)
    for B, H, W, C in ((1, 5, 9, 8), (2, 11, 4, 5)):
        x = torch.randn(B, H, W, C, generator=g)
        w = torch.randn(C, 1, 7, 7, generator=g)
        b = torch.randn(C, generator=g)
        out = torch.empty_like(x)
        fn(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), B, H, W, C)
        want = F.conv2d(x.permute(0, 3, 1, 2), w, b, padding=3, groups=C).permute(0, 2, 3, 1)
        assert torch.allclose(out, want, atol=1e-5)


def _product(cfg, name="CLIP-convnext_xxlarge-res96"):
    from visper_lm_b200.model.convnext import CLIPConvNextVisionTower

    m = CLIPConvNextVisionTower(name, args=SimpleNamespace(mm_vision_select_layer=-2), cfg=dict(cfg, image_size=96))
    sd = restate.convnext_seeded_state(cfg, PREFIX)
    own = {PREFIX + n: tuple(p.shape) for n, p in m.vision_tower.named_parameters()}
    assert own == {k: tuple(v.shape) for k, v in sd.items()}          # timm's names and shapes, nothing else
    with torch.no_grad():
        for n, p in m.vision_tower.named_parameters():
            p.copy_(sd[PREFIX + n])
    return m


class _TorchOps:
    """Stand-ins for visper_lm_b200.ops with the kernels' contracts: bf16 storage, fp32 arithmetic."""

    def __init__(self):
        self.calls = []

    def cast_bf16(self, x):
        return x.to(BF16)

    def im2col_patches(self, images, patch, kpad):
        B, C, H, W = images.shape
        cols = F.unfold(images.float(), patch, stride=patch).transpose(1, 2).reshape(-1, C * patch * patch)
        out = torch.zeros(cols.shape[0], kpad, dtype=BF16)
        out[:, :cols.shape[1]] = cols                              # K order (c, ky, kx), zero padded
        return out

    def gemm(self, a, b, bias=None, act=0, residual=None):
        self.calls.append("gemm")
        y = a.float() @ b.float().t()
        if bias is not None:
            y = y + bias.float()
        if act == 1:
            y = F.gelu(y)
        if residual is not None:
            y = y + residual.float()
        return y.to(BF16)

    def layernorm_fwd(self, x, w, b, eps):
        self.calls.append("ln")
        return F.layer_norm(x.float(), (x.shape[1],), w.float(), b.float(), eps).to(BF16), None, None

    def gather_rows(self, index, srcs, D, out=None):
        self.calls.append("gather")
        r = srcs[0][index.long().clamp_min(0)]
        r[index.long() < 0] = 0
        out.copy_(r)
        return out

    def dwconv7x7(self, x, w49, bias, B, H, W, C):
        self.calls.append("dwconv")
        w = w49.float().t().reshape(C, 1, 7, 7)
        y = F.conv2d(x.float().view(B, H, W, C).permute(0, 3, 1, 2), w, bias.float(), padding=3, groups=C)
        return y.permute(0, 2, 3, 1).reshape(B * H * W, C).to(BF16)


def test_product_forward_rows_on_torch_ops(monkeypatch):
    from visper_lm_b200.model import convnext

    fx = torch.load(GOLDEN / "convnext_mini_96.pt")
    cfg = fx["config"]
    m = _product(cfg)
    fake = _TorchOps()
    monkeypatch.setattr(convnext, "ops", fake)
    px = pixels(fx["B"], fx["size"], fx["seed"])
    rows, H, W, stages = m.vision_tower.forward_rows(px, return_stages=True)
    assert (H, W) == (3, 3) and rows.shape == (fx["B"] * 9, cfg["dims"][-1]) and rows.dtype == BF16
    # oracle on the SAME bf16-rounded weights and pixels
    sd = {PREFIX + n: p.detach().float() for n, p in m.vision_tower.named_parameters()}
    with torch.no_grad():
        want = restate.convnext_stage_features(sd, px.to(BF16).float(), cfg, PREFIX)
    for (x, h, w), ref in zip(stages, want):
        got = x.float().view(fx["B"], h, w, -1).permute(0, 3, 1, 2)
        assert got.shape == ref.shape
        rel = ((got - ref).norm() / ref.norm()).item()
        assert rel < 1.5e-2, rel                                   # bf16 storage between kernels; layout bugs are O(1)
    assert torch.equal(m(px), rows)                                # forward() = the last stage's rows, B-major
    want_rows = restate.convnext_tower(sd, px.to(BF16).float(), cfg, PREFIX).reshape(-1, cfg["dims"][-1])
    assert ((rows.float() - want_rows).norm() / want_rows.norm()).item() < 1.5e-2
    # launch budget: stem 2 + per downsample (LN, 4 gathers, GEMM) + 4 per block
    n_blocks = sum(cfg["depths"])
    assert len(fake.calls) // 2 == (2 + 3 * 6 + 4 * n_blocks)      # forward ran twice (rows, m(px))
    assert fake.calls.count("dwconv") == 2 * n_blocks


def test_folded_weights():
    from visper_lm_b200.model.convnext import MERGE_TAPS, fold_block, fold_downsample
    from visper_lm_b200.model.seg_teacher import merge_plans

    g = torch.Generator().manual_seed(0)
    C = 8
    w_dw, gamma = torch.randn(C, 1, 7, 7, generator=g), torch.randn(C, generator=g)
    w2, b2 = torch.randn(C, 4 * C, generator=g), torch.randn(C, generator=g)
    w49, w2f, b2f = fold_block(w_dw, gamma, w2, b2)
    assert w49.shape == (49, C) and torch.equal(w49[3 * 7 + 5], w_dw[:, 0, 3, 5])
    h = torch.randn(5, 4 * C, generator=g)
    assert torch.allclose(F.linear(h, w2f, b2f), gamma * F.linear(h, w2, b2), atol=1e-5)
    # 2x2 stride-2 conv == merge gathers + folded weight
    B, H, W, Co = 2, 6, 4, 5
    x = torch.randn(B, H, W, C, generator=g)
    w = torch.randn(Co, C, 2, 2, generator=g)
    idx, H2, W2 = merge_plans(B, H, W)
    assert len(idx) == len(MERGE_TAPS)
    cat = torch.cat([x.reshape(-1, C)[i.long()] for i in idx], 1)
    got = (cat @ fold_downsample(w).t()).view(B, H2, W2, Co)
    want = F.conv2d(x.permute(0, 3, 1, 2), w, stride=2).permute(0, 2, 3, 1)
    assert torch.allclose(got, want, atol=1e-4)


def test_name_parsing_and_properties():
    """extract_res_interp / the size properties (clip_convnext_encoder.py:34-60, 176-218)."""
    from visper_lm_b200.model.convnext import CLIPConvNextVisionTower, extract_res_interp

    assert extract_res_interp("CLIP-convnext_xxlarge-res768") == ("CLIP-convnext_xxlarge", 768, None)
    assert extract_res_interp("some/dir/CLIP-convnext_large-res320-interp256") == ("CLIP-convnext_large", 320, 256)
    with pytest.raises(ValueError):
        extract_res_interp("openai/clip-vit-large-patch14-336")
    t = CLIPConvNextVisionTower("CLIP-convnext_xxlarge-res768", args=SimpleNamespace(mm_vision_select_layer=-2),
                                device="meta")
    assert (t.hidden_size, t.image_size, t.num_patches_per_side, t.num_patches) == (3072, 768, 24, 576)
    assert t.ckpt_path == "CLIP-convnext_xxlarge" and t.cfg["eps"] == 1e-5
    assert sum(p.numel() for p in t.parameters()) == 843_391_872     # the xxlarge trunk without its head


def test_config_selects_the_tower():
    """multimodal_encoder/builder.py:6-13: 'convnext' in mm_vision_tower picks the ConvNeXt class; the
    projector's input width follows the tower (3072)."""
    from visper_lm_b200.model.convnext import CLIPConvNextVisionTower
    from visper_lm_b200.model.vlm import VisperConfig, VisperModel

    cfg = VisperConfig(vocab_size=64, hidden_size=64, intermediate_size=128, num_hidden_layers=1,
                       num_attention_heads=2, num_key_value_heads=2, mm_vision_tower="CLIP-convnext_xxlarge-res768")
    assert cfg.mm_hidden_size == 3072 and cfg.vision["image_size"] == 768
    m = VisperModel(cfg, device="meta")
    assert isinstance(m.vision_tower, CLIPConvNextVisionTower)
    assert m.mm_projector[0].weight.shape == (64, 3072) and m.vision_tower.num_patches == 576
    assert "vision_tower.vision_tower.stages.2.blocks.29.mlp.fc2.weight" in dict(m.named_parameters())


def test_image_processor_matches_torchvision():
    """ProcessorWrapper(OpenClipEvalTransform) == open_clip's eval transform as the reference resizes it
    (clip_convnext_encoder.py:113-115), spelled with torchvision here (open_clip is not installed)."""
    tv = pytest.importorskip("torchvision.transforms")
    import numpy as np
    from PIL import Image

    from visper_lm_b200.model.convnext import OPENAI_CLIP_MEAN, OPENAI_CLIP_STD, CLIPConvNextVisionTower

    t = CLIPConvNextVisionTower("CLIP-convnext_xxlarge-res96", args=SimpleNamespace(mm_vision_select_layer=-2),
                                device="meta", cfg=dict(MINI))
    proc = t.image_processor
    assert proc.crop_size == {"height": 96, "width": 96} and proc.image_mean == list(OPENAI_CLIP_MEAN)
    ref = tv.Compose([tv.Resize(96, interpolation=tv.InterpolationMode.BICUBIC), tv.CenterCrop((96, 96)),
                      lambda im: im.convert("RGB"), tv.ToTensor(), tv.Normalize(OPENAI_CLIP_MEAN, OPENAI_CLIP_STD)])
    rng = np.random.default_rng(0)
    for (w, h), mode in (((200, 131), "RGB"), ((77, 150), "RGB"), ((96, 96), "RGB"), ((40, 300), "L"), ((640, 480), "RGBA")):
        arr = rng.integers(0, 256, (h, w, {"RGB": 3, "L": 1, "RGBA": 4}[mode]), dtype=np.uint8)
        im = Image.fromarray(arr.squeeze(-1) if mode == "L" else arr, mode)
        got = proc.preprocess(im, return_tensors="pt")["pixel_values"][0]
        want = ref(im)
        assert got.shape == (3, 96, 96) and torch.allclose(got, want, atol=1e-6), (w, h, mode)
    assert torch.equal(proc([im])["pixel_values"][0], got)           # list input / __call__ (base_encoder.py:23-31)


def test_config_json_round_trip_keeps_the_tower():
    """config.json as trainer._save writes it → from_pretrained's constructor call: the ConvNeXt tower survives."""
    import json

    from visper_lm_b200.model.convnext import CLIPConvNextVisionTower
    from visper_lm_b200.model.vlm import VisperConfig, VisperModel
    from visper_lm_b200.train.checkpoint import config_to_dict

    cfg = VisperConfig(vocab_size=64, hidden_size=64, intermediate_size=128, num_hidden_layers=1, num_attention_heads=2,
                       num_key_value_heads=2, mm_vision_tower="CLIP-convnext_xxlarge-res768")
    d = json.loads(json.dumps(config_to_dict(cfg)))
    for k in ("model_type", "family", "architectures"):
        d.pop(k, None)
    again = VisperConfig(**d)
    assert again.mm_vision_tower == cfg.mm_vision_tower and again.mm_hidden_size == 3072
    assert list(again.vision["dims"]) == [384, 768, 1536, 3072] and again.vision["eps"] == 1e-5
    m = VisperModel(again, device="meta")
    assert isinstance(m.vision_tower, CLIPConvNextVisionTower) and m.mm_projector[0].weight.shape == (64, 3072)
