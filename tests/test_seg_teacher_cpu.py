"""CPU tests of the frozen segmentation teacher: the oracle restatement against golden
vectors of transformers' SwinBackbone (the model behind the reference's `oneformer.forward_features`),
the state-dict ABI of the product module, and — with the product's own index plans, masks, bias tables
and fused weights driven through plain torch — the whole pad / shift / partition / reverse / merge logic
the CUDA path executes as gathers."""
import pytest
import torch
import torch.nn.functional as F

from parity_utils import restate

from oracle.make_golden_seg_teacher import PREFIX, seg_param, seg_pixels

GOLDEN = __import__("pathlib").Path(__file__).parent / "golden"


@pytest.mark.parametrize("name", ["seg_teacher_mini_120", "seg_teacher_swinl_800"])
def test_oracle_matches_library_golden(name):
    fx = torch.load(GOLDEN / f"{name}.pt")
    cfg = fx["config"]
    sd = {n: seg_param(n, s) for n, s in fx["state_spec"].items()}
    px = seg_pixels(fx["B"], fx["size"], fx["seed"])
    with torch.no_grad():
        maps = restate.swin_stage_features(sd, px, cfg, PREFIX)
        tgt = F.interpolate(maps[-1], size=(24, 24), mode="bilinear", align_corners=False)
    assert [tuple(m.shape[1:]) for m in maps] == fx["stage_shapes"]
    assert torch.allclose(tgt[:, ::8, ::2, ::2], fx["targets_sub"], atol=3e-5)
    assert torch.allclose(maps[0][:, ::8, ::5, ::5], fx["stage1_sub"], atol=3e-5)
    assert abs(float(tgt.std()) - fx["tgt_std"]) < 1e-5


def _product(cfg, fx):
    from visper_lm_b200.model.seg_teacher import OneFormerHead

    m = OneFormerHead(cfg)
    assert {"oneformer." + n: tuple(p.shape) for n, p in m.named_parameters()} == fx["state_spec"]
    with torch.no_grad():
        for n, p in m.named_parameters():
            p.copy_(seg_param("oneformer." + n, tuple(p.shape)))
    return m


def _emulate(backbone, px):
    """last_feature_rows with torch standing in for the C-ABI calls (fp32), same plans / derived weights."""
    cfg = backbone.cfg
    E, ws, P = cfg["embed_dim"], cfg["window_size"], cfg["patch_size"]
    B = px.shape[0]
    lin, ln = F.linear, F.layer_norm
    f32 = lambda t: t.detach().float()  # noqa: E731

    def gather(idx, src):
        out = src[idx.long().clamp_min(0)]
        out[idx.long() < 0] = 0
        return out

    pe = backbone.embeddings.patch_embeddings.projection
    x = F.conv2d(px, f32(pe.weight), f32(pe.bias), stride=P)
    H, W = x.shape[-2:]
    x = x.flatten(2).transpose(1, 2).reshape(B * H * W, E)
    x = ln(x, (E,), f32(backbone.embeddings.norm.weight), f32(backbone.embeddings.norm.bias), 1e-5)
    for s, stage in enumerate(backbone.encoder.layers):
        C = E * 2 ** s
        heads = cfg["num_heads"][s]
        hd = C // heads
        for i, blk in enumerate(stage.blocks):
            shift = 0 if i % 2 == 0 else ws // 2
            part, rev, Hp, Wp, mask = backbone._plan("win", B, H, W, ws, shift)
            nW = (Hp // ws) * (Wp // ws)
            wqkv, bqkv, bias = blk.derived(backbone._rel_index)
            h = ln(x, (C,), f32(blk.layernorm_before.weight), f32(blk.layernorm_before.bias), 1e-5)
            qkv = lin(gather(part, h), f32(wqkv), f32(bqkv)).view(B * nW, ws * ws, 3, heads, hd).permute(2, 0, 3, 1, 4)
            sc = qkv[0] @ qkv[1].transpose(-1, -2) * hd ** -0.5 + bias[None]
            if mask is not None:
                sc = sc + mask.repeat(B, 1, 1)[:, None]            # window b uses mask[b % nW]
            ctx = (torch.softmax(sc, -1) @ qkv[2]).transpose(1, 2).reshape(B * nW * ws * ws, C)
            od = blk.attention.output.dense
            x = x + lin(gather(rev, ctx), f32(od.weight), f32(od.bias))
            h = ln(x, (C,), f32(blk.layernorm_after.weight), f32(blk.layernorm_after.bias), 1e-5)
            h = F.gelu(lin(h, f32(blk.intermediate.dense.weight), f32(blk.intermediate.dense.bias)))
            x = x + lin(h, f32(blk.output.dense.weight), f32(blk.output.dense.bias))
        if s + 1 < len(cfg["depths"]):
            idx, H2, W2 = backbone._plan("merge", B, H, W)
            cat = torch.cat([gather(ix, x) for ix in idx], 1)
            ds = stage.downsample
            x = lin(ln(cat, (4 * C,), f32(ds.norm.weight), f32(ds.norm.bias), 1e-5), f32(ds.reduction.weight))
            H, W = H2, W2
    nrm = backbone.hidden_states_norms[f"stage{len(cfg['depths'])}"]
    return ln(x, (x.shape[1],), f32(nrm.weight), f32(nrm.bias), 1e-5), H, W


def test_index_plans_reproduce_the_reference_layout():
    """Product plans (window gather with padding + cyclic shift, reverse gather, 4-way merge gathers,
    shift masks, bias tables, fused QKV) in fp32 torch == the oracle, on the miniature whose grid needs
    window padding (30→32, 15→16), shifted-window masks and an odd patch merge (15→8)."""
    fx = torch.load(GOLDEN / "seg_teacher_mini_120.pt")
    cfg = fx["config"]
    m = _product(cfg, fx)
    px = seg_pixels(fx["B"], fx["size"], fx["seed"])
    with torch.no_grad():
        rows, H, W = _emulate(m.pixel_level_module.encoder, px)
    sd = {n: p.detach().float() for n, p in m.named_parameters()}
    with torch.no_grad():
        ref = restate.swin_stage_features(sd, px, cfg)[-1]
    got = rows.view(fx["B"], H, W, -1).permute(0, 3, 1, 2)
    assert got.shape == ref.shape
    assert (got - ref).abs().max().item() < 2e-2 * ref.abs().max().item()   # product params are bf16-rounded...
    assert ((got - ref).norm() / ref.norm()).item() < 1e-4                  # ...but identical on both sides


def test_plans_small_cases():
    from visper_lm_b200.model.seg_teacher import merge_plans, shift_mask, window_plans

    part, rev, Hp, Wp = window_plans(2, 5, 7, 4, 2)
    assert (Hp, Wp) == (8, 8) and part.numel() == 2 * 64 and rev.numel() == 2 * 35
    x = torch.arange(2 * 35, dtype=torch.float32).view(2, 5, 7)
    ref = torch.roll(F.pad(x + 1, (0, 1, 0, 3)), (-2, -2), (1, 2))         # +1 so padding (0) is distinguishable
    ref = ref.view(2, 2, 4, 2, 4).permute(0, 1, 3, 2, 4).reshape(-1)
    got = torch.where(part >= 0, x.reshape(-1)[part.long().clamp_min(0)] + 1, torch.zeros(()))
    assert torch.equal(got, ref)
    assert torch.equal(part[rev.long()].long(), torch.arange(2 * 35))       # reverse undoes partition
    assert torch.equal(shift_mask(8, 8, 4, 2), restate.swin_shift_mask(8, 8, 4, 2))
    idx, H2, W2 = merge_plans(1, 5, 4)
    assert (H2, W2) == (3, 2) and idx[1].tolist() == [4, 6, 12, 14, -1, -1] and idx[3].tolist() == [5, 7, 13, 15, -1, -1]


def test_plans_property_random_grids():
    """hypothesis: for random grids / window sizes / shifts the gather plans equal F.pad → torch.roll →
    window_partition (and its inverse), the shift mask equals HF's get_attn_mask, and the merge plans equal
    SwinPatchMerging's strided slices."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    from visper_lm_b200.model.seg_teacher import merge_plans, shift_mask, window_plans

    @settings(max_examples=60, deadline=None)
    @given(B=st.integers(1, 3), H=st.integers(1, 23), W=st.integers(1, 23), ws=st.sampled_from([2, 3, 4, 7, 12]),
           shifted=st.booleans())
    def check(B, H, W, ws, shifted):
        shift = ws // 2 if shifted else 0
        part, rev, Hp, Wp = window_plans(B, H, W, ws, shift)
        x = torch.arange(B * H * W, dtype=torch.float32).view(B, H, W) + 1
        ref = F.pad(x, (0, Wp - W, 0, Hp - H))
        if shift:
            ref = torch.roll(ref, (-shift, -shift), (1, 2))
        win = ref.view(B, Hp // ws, ws, Wp // ws, ws).permute(0, 1, 3, 2, 4).reshape(-1)
        got = torch.where(part >= 0, x.reshape(-1)[part.long().clamp_min(0)], torch.zeros(()))
        assert torch.equal(got, win)
        back = win.view(B, Hp // ws, Wp // ws, ws, ws).permute(0, 1, 3, 2, 4).reshape(B, Hp, Wp)   # window_reverse
        if shift:
            back = torch.roll(back, (shift, shift), (1, 2))
        assert torch.equal(win[rev.long()], back[:, :H, :W].reshape(-1))
        if shift:
            assert torch.equal(shift_mask(Hp, Wp, ws, shift), restate.swin_shift_mask(Hp, Wp, ws, shift))
        idx, H2, W2 = merge_plans(B, H, W)
        xp = F.pad(x, (0, W % 2, 0, H % 2))
        for ix, sl in zip(idx, (xp[:, 0::2, 0::2], xp[:, 1::2, 0::2], xp[:, 0::2, 1::2], xp[:, 1::2, 1::2])):
            g = torch.where(ix >= 0, x.reshape(-1)[ix.long().clamp_min(0)], torch.zeros(()))
            assert (H2, W2) == tuple(sl.shape[1:]) and torch.equal(g, sl.reshape(-1))

    check()
