"""CPU tests of the frozen generation teacher: the oracle restatement against golden
image_embeds of transformers' CLIPVisionModelWithProjection (the model behind the reference's
`pipe.image_encoder`), the state-dict ABI of the product module, and the head_dim 80 → 96 padding."""
import pytest
import torch

from parity_utils import restate

from oracle.make_golden_gen_teacher import gen_pixels

GOLDEN = __import__("pathlib").Path(__file__).parent / "golden"


@pytest.mark.parametrize("name", ["gen_teacher_mini", "gen_teacher_vith_224"])
def test_oracle_matches_library_golden(name):
    fx = torch.load(GOLDEN / f"{name}.pt")
    cfg = fx["config"]
    sd = {n: restate.seeded_param(n, s) for n, s in fx["state_spec"].items()}
    px = gen_pixels(fx["B"], cfg["image_size"], fx["seed"])
    with torch.no_grad():
        emb = restate.gen_teacher_targets(sd, px, cfg["num_attention_heads"], cfg["hidden_act"], "image_encoder.")
    assert emb.shape == fx["image_embeds"].shape == (fx["B"], 1, cfg["projection_dim"])
    assert torch.allclose(emb, fx["image_embeds"], atol=3e-5)


def test_state_dict_abi_and_head_padding():
    from visper_lm_b200.model.gen_teacher import CLIPVisionModelWithProjection

    fx = torch.load(GOLDEN / "gen_teacher_mini.pt")
    cfg = fx["config"]
    m = CLIPVisionModelWithProjection(cfg)
    assert {"image_encoder." + n: tuple(p.shape) for n, p in m.named_parameters()} == fx["state_spec"]
    with torch.no_grad():
        for n, p in m.named_parameters():
            p.copy_(restate.seeded_param("image_encoder." + n, tuple(p.shape)))
    heads, D = cfg["num_attention_heads"], cfg["hidden_size"]
    hd, hp = D // heads, 96
    qkv_w, qkv_b, out_w = m._layer_weights(1, heads, hd, hp)
    a = m.vision_model.encoder.layers[1].self_attn
    x = torch.randn(2, 9, D)

    def attend(q, k, v, h):
        q, k, v = (t.view(2, 9, heads, h).transpose(1, 2) for t in (q, k, v))
        return (torch.softmax(q @ k.transpose(-1, -2) * hd ** -0.5, -1) @ v).transpose(1, 2).reshape(2, 9, heads * h)

    lin = torch.nn.functional.linear
    ref = lin(attend(*(lin(x, getattr(a, f"{n}_proj").weight.float(), getattr(a, f"{n}_proj").bias.float())
                       for n in "qkv"), hd), a.out_proj.weight.float())
    qkv = lin(x, qkv_w.float(), qkv_b.float())
    W = heads * hp
    got = lin(attend(qkv[..., :W], qkv[..., W:2 * W], qkv[..., 2 * W:], hp), out_w.float())
    assert torch.allclose(got, ref, atol=1e-5)             # zero padding changes nothing
    assert m._layer_weights(1, heads, hd, hp)[0] is qkv_w  # cached per weight version
