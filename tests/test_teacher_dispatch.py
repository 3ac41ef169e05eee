"""Host logic of the distillation-target plumbing (no GPU): which source feeds each task's targets —
caller-supplied tensors, the on-GPU teachers behind the reference's hooks (base_ola_vlm.py:323,347,382),
or nothing — and init_target_models' failure mode when the (un-downloadable) weights are absent."""
from types import SimpleNamespace

import pytest
import torch

from visper_lm_b200.model.vlm import OlaLlavaLlamaForCausalLM as M


class _Stub(SimpleNamespace):
    def _get_dav2_feats(self, images, device, decode=True):
        self.calls.append(("depth", decode))
        return [(torch.full((len(images), 576, 4), 1.0), None)], None

    def _get_gen_feats(self, images, device):
        self.calls.append(("gen",))
        return torch.full((len(images), 1, 4), 2.0)

    def _get_seg_targets(self, images, seg_preds):
        self.calls.append(("seg",))
        return torch.full((len(images), 4, 24, 24), 3.0)

    def _seg_pixel_values(self, images):
        return images


def _host(**kw):
    return _Stub(calls=[], **kw)


def test_caller_targets_win():
    h = _host(dav2_backbone=object(), pipe=object())
    t = {"depth": torch.zeros(2, 576, 4)}
    assert M._targets(h, "depth", [object(), object()], t, "cpu") is t["depth"] and h.calls == []


def test_teacher_fills_missing_task_when_images_present():
    h = _host(dav2_backbone=object(), pipe=object())
    out = M._targets(h, "depth", torch.zeros(2, 336, 336, 3, dtype=torch.uint8), {"seg": torch.zeros(1)}, "cpu")
    assert out.shape == (2, 576, 4) and h.calls == [("depth", False)]       # the logging-only decode is skipped
    out = M._targets(h, "gen", {"gen": torch.zeros(2, 3, 224, 224)}, {}, "cpu")  # per-task preprocessed tensors
    assert out.shape == (2, 1, 4) and h.calls[-1] == ("gen",)


def test_no_teacher_no_targets():
    h = _host()
    assert M._targets(h, "seg", [None, None], {"depth": torch.zeros(1)}, "cpu") is None   # caller gave others only
    assert M._targets(h, "gen", None, None, "cpu") is None
    assert M._targets(h, "depth", {"gen": torch.zeros(1)}, {}, "cpu") is None              # no depth images in the dict
    assert h.calls == []


def test_reference_hooks_without_caller_targets():
    h = _host()
    out = M._targets(h, "seg", [object()], None, "cpu")      # reference behaviour: hook is called
    assert out.shape == (1, 4, 24, 24) and h.calls == [("seg",)]


def test_seg_teacher_rows_skip_the_nchw_round_trip():
    net = SimpleNamespace(seg_target_rows=lambda px: torch.arange(2 * 576 * 8, dtype=torch.float32).view(2 * 576, 8))
    h = _host(oneformer=net)
    out = M._targets(h, "seg", torch.zeros(2, 3, 8, 8), {}, "cpu")
    assert out.shape == (2, 576, 8) and h.calls == []


def test_hooks_raise_before_init():
    h = SimpleNamespace()
    for fn, args in ((M._get_dav2_feats, ([None], "cpu")), (M._get_gen_feats, ([None], "cpu")),
                     (M._get_seg_targets, ([None], None))):
        with pytest.raises(NotImplementedError):
            fn(h, *args)


def test_init_target_models_needs_weights_or_opt_in():
    cfg = SimpleNamespace(aux_mode="depth", image_depth={"x": 1}, depth_estimator="/nonexistent/depth_anything_v2_vitl.pth")
    h = SimpleNamespace(_device=None)
    with pytest.raises(FileNotFoundError):
        M.init_target_models(h, cfg)
    cfg.random_init_teachers = True
    M.init_target_models(h, cfg)
    assert sum(p.numel() for p in h.dav2_backbone.pretrained.parameters()) == 304_368_640   # DINOv2-L
    assert not any(p.requires_grad for p in h.dav2_backbone.parameters())


def test_target_tiling_matches_reference_emb_loss():
    """Fewer target rows than predictions (base_ola_vlm.py:292-299): product helper and oracle tile the
    way the reference's own _emb_loss does."""
    from oracle import ref_shim, restate
    from visper_lm_b200.model.vlm import tile_targets

    g = torch.Generator().manual_seed(0)
    preds = torch.randn(4, 6, 8, generator=g)
    tgt = torch.randn(2, 6, 8, generator=g)
    mask = torch.tensor([1, 0])
    t2, m2 = tile_targets(tgt, mask, 4)
    assert torch.equal(t2, torch.cat([tgt, tgt])) and m2.tolist() == [1, 0, 1, 0]
    t3, m3 = tile_targets(tgt, torch.ones(4), 4)
    assert m3.shape == (4,) and torch.equal(t3, t2)
    same, msame = tile_targets(preds, mask, 4)
    assert same is preds and msame is mask
    scale = torch.tensor(2.0)
    mine = restate.emb_loss(preds, mask, tgt, scale)
    want = restate.emb_loss(preds, m2, t2, scale)
    assert all(torch.equal(a, b) for a, b in zip(mine, want))
    if ref_shim.available():
        R = ref_shim.load()
        ref = R.base.BaseOLA_VLM._emb_loss(SimpleNamespace(contrastive_loss_weight=0.3), preds, mask, tgt, scale)
        for a, b in zip(mine, ref):
            assert abs(float(a) - float(b)) < 1e-6


def test_freeze_policy_sets():
    """PT (tune_mm_mlp_adapter) trains projector + heads + task tokens + logit scales only — the set
    bench.py selects by name and the golden gradients cover; fine-tuning trains everything but the tower,
    the DPT decoder and the teachers."""
    from parity_utils import build_product, configs
    from visper_lm_b200.train.policy import apply_freeze_policy

    model = build_product(configs.TINY_LLAMA, True, None)
    pt = apply_freeze_policy(model, tune_mm_mlp_adapter=True)
    fx = torch.load(__import__("pathlib").Path(__file__).parent / "golden" / "tiny_llama_dsg.pt")
    # the reference's PT-stage tensors that received a gradient; linear_2 / linear_3 of the depth heads are
    # trainable too but get none, so the golden does not list them
    extra = set(pt) - set(fx["grads"])
    assert set(fx["grads"]) <= set(pt) and all(("linear_2" in n) or ("linear_3" in n) for n in extra)
    assert all(("mm_projector" in n) or ("_heads." in n) or ("special_" in n) or n.endswith("logit_scale") for n in pt)
    ft = apply_freeze_policy(model)
    assert not any(("vision_tower" in n) or ("da_v2_head" in n) for n in ft)
    assert {"lm_head.weight", "model.embed_tokens.weight", "model.layers.0.self_attn.q_proj.weight"} <= set(ft)
    assert set(pt) <= set(ft)
    frozen_tokens = apply_freeze_policy(model, tune_mm_mlp_adapter=True, freeze_task_token=True, freeze_mm_mlp_adapter=True)
    assert not any(("special_" in n) or ("mm_projector" in n) for n in frozen_tokens) and frozen_tokens
