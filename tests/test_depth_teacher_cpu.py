"""CPU tests of the frozen depth teacher: the oracle restatement against the golden
vectors of the UNMODIFIED reference DepthAnythingV2 / DAv2_Head, the state-dict ABI of the product
module, and the host-side derivations the CUDA path relies on (weights folded once per version,
position-table interpolation, tap-mean index) — none of which needs a GPU."""
import pytest
import torch
import torch.nn.functional as F

from parity_utils import restate

from oracle.make_golden_dinov2 import teacher_images, teacher_param

GOLDEN = __import__("pathlib").Path(__file__).parent / "golden"
PREFIX = "dav2_backbone.pretrained."


@pytest.mark.parametrize("name", ["dinov2_vits_224", "dav2_teacher_vitl_336"])
def test_oracle_matches_reference_golden(name):
    fx = torch.load(GOLDEN / f"{name}.pt")
    sd = {n: teacher_param(n, s) for n, s in fx["state_spec"].items()}
    hsd = None if fx["head_spec"] is None else {n: restate.seeded_param(n, s) for n, s in fx["head_spec"].items()}
    raw = teacher_images(fx["B"], fx["size"], fx["seed"])
    with torch.no_grad():
        ft, gts = restate.dav2_depth_teacher(sd, hsd, raw, fx["encoder"], prefix=PREFIX)
        taps = restate.dinov2_intermediate(sd, restate.dav2_image_tensor(raw), fx["encoder"], PREFIX)
    assert torch.allclose(ft[:, ::7, ::16], fx["ft_sub"], atol=2e-5)
    assert abs(float(ft.mean()) - fx["ft_mean"]) < 1e-5 and abs(float(ft.std()) - fx["ft_std"]) < 1e-5
    assert torch.allclose(torch.stack([t[1] for t in taps], 1)[:, :, ::16], fx["cls_sub"], atol=2e-5)
    if gts is not None:
        assert torch.allclose(gts[:, ::7, ::7], fx["depth_gts_sub"], atol=2e-5)


def test_teacher_state_dict_abi():
    """Parameter names / shapes equal the reference DepthAnythingV2.pretrained's (the golden's spec),
    so depth_anything_v2_vitl.pth loads unchanged."""
    from visper_lm_b200.model.dinov2 import DepthAnythingV2

    for name, enc in (("dinov2_vits_224", "vits"), ("dav2_teacher_vitl_336", "vitl")):
        fx = torch.load(GOLDEN / f"{name}.pt")
        m = DepthAnythingV2(enc, with_depth_head=(enc == "vitl"))
        mine = {"dav2_backbone." + n: tuple(p.shape) for n, p in m.named_parameters() if n.startswith("pretrained.")}
        assert mine == fx["state_spec"]
        if enc == "vitl":  # the DPT half of the .pth: same keys as DAv2_Head's (strict load, base_ola_vlm.py:81)
            head = {n.replace("da_v2_head.", ""): s for n, s in fx["head_spec"].items()}
            assert {n: tuple(p.shape) for n, p in m.named_parameters() if n.startswith("depth_head.")} == head


def _seeded_vit(enc):
    from visper_lm_b200.model.dinov2 import DinoVisionTransformer

    vt = DinoVisionTransformer(enc)
    with torch.no_grad():
        for n, p in vt.named_parameters():
            p.copy_(teacher_param(PREFIX + n, tuple(p.shape)))
    return vt


def test_patch_conv_folds_image2tensor():
    """conv(w', raw uint8) + b' == conv(w, image2tensor(raw)) + b: pixel scaling, mean/std and the
    reference's channel reversal live in the folded weights."""
    vt = _seeded_vit("vits")
    raw = teacher_images(2, 56, 5)
    wp, bp, kpad = vt._patch_weight(True)
    w, b = vt.patch_embed.proj.weight.float(), vt.patch_embed.proj.bias.float()
    ref = F.conv2d(restate.dav2_image_tensor(raw), w, b, stride=14)
    from visper_lm_b200.model.dinov2 import DepthAnythingV2

    da = DepthAnythingV2("vits", with_depth_head=False)
    x = da._raw_batch(raw, 56)                             # centred pixels, exact in bf16
    assert torch.equal(x.float(), raw.permute(0, 3, 1, 2).float() - torch.tensor([104., 116., 124.]).view(1, 3, 1, 1))
    got = F.conv2d(x.float(), wp[:, :588].float().view(-1, 3, 14, 14), bp.float(), stride=14)
    assert kpad == 640 and float(wp[:, 588:].abs().max()) == 0.0
    plain = F.conv2d(restate.dav2_image_tensor(raw).to(torch.bfloat16).float(), w, b, stride=14)
    assert (got - ref).norm() / ref.norm() < 3e-3          # one bf16 rounding of the folded weights...
    assert (got - ref).norm() <= 1.2 * (plain - ref).norm()  # ...no worse than rounding the normalised pixels
    wn, bn, _ = vt._patch_weight(False)
    assert torch.equal(wn[:, :588], vt.patch_embed.proj.weight.view(-1, 588)) and torch.equal(bn, vt.patch_embed.proj.bias)


def test_position_table_and_tap_mean_index():
    vt = _seeded_vit("vits")
    pe = vt.pos_embed.float()
    for n in (16, 24, 37):
        ref = restate.dinov2_pos_embed(pe, n, n)[0]
        assert torch.equal(vt._pos_table(n, n), ref.to(torch.bfloat16))
    B, S, D = 3, 5, 8
    taps = torch.randn(4 * B * S, D)
    idx = vt._tap_mean_index(B, S).long().view(-1, 4)
    got = taps[idx].sum(1) * 0.25
    ref = taps.view(4, B, S, D)[:, :, 1:].mean(0).reshape(-1, D)
    assert torch.allclose(got, ref, atol=1e-6)


def test_layerscale_fold():
    vt = _seeded_vit("vits")
    blk = vt.blocks[3]
    pw, pb, fw, fb = blk.folded()
    x = torch.randn(7, 384)
    ref = blk.ls1.gamma.float() * F.linear(x, blk.attn.proj.weight.float(), blk.attn.proj.bias.float())
    got = F.linear(x, pw.float(), pb.float())
    assert (got - ref).norm() / ref.norm() < 4e-3
    y = torch.randn(7, 1536)
    ref = blk.ls2.gamma.float() * F.linear(y, blk.mlp.fc2.weight.float(), blk.mlp.fc2.bias.float())
    assert (F.linear(y, fw.float(), fb.float()) - ref).norm() / ref.norm() < 4e-3
    assert blk.folded()[0] is pw                          # cached until a weight changes
    with torch.no_grad():
        blk.ls1.gamma.mul_(2.0)
    assert blk.folded()[0] is not pw
