"""CPU tests (no GPU): the oracle restatement against the golden vectors of the unmodified
reference, the reference itself when /root/reference is present, and the state-dict ABI."""
import pytest
import torch

from parity_utils import configs, restate

GOLDEN = __import__("pathlib").Path(__file__).parent / "golden"
CASES = [("tiny_llama_dsg", "TINY_LLAMA", True), ("tiny_llama_dsg_padded", "TINY_LLAMA", True),
         ("tiny_phi3_dsg", "TINY_PHI3", True), ("tiny_phi3_sw_dsg", "TINY_PHI3_SW", True),
         ("tiny_llama_ntp", "TINY_LLAMA", False), ("tiny_llama_ntp_mixed", "TINY_LLAMA", False),
         ("wide_llama_dsg", "WIDE_LLAMA", True), ("wide_phi3_dsg", "WIDE_PHI3", True)]


def _state(fx):
    return {n: restate.seeded_param(n, s) for n, s in fx["state_spec"].items()}


@pytest.mark.parametrize("name,cfg_name,distill", CASES)
def test_oracle_matches_golden(name, cfg_name, distill):
    fx = torch.load(GOLDEN / f"{name}.pt")
    cfg = getattr(configs, cfg_name)
    sd = {k: v.requires_grad_(k in fx["grads"]) for k, v in _state(fx).items()}
    batch = (configs.synthetic_batch_mixed(cfg, fx["n_text"], seed=fx["seed"]) if fx["pad_rows"] < 0 else
             configs.synthetic_batch(cfg, fx["B"], fx["n_text"], seed=fx["seed"], distill=distill,
                                     pad_rows=fx["pad_rows"]))
    pub = restate.forward_step(sd, cfg, batch, distill=distill, zero_masks_like_reference=True)
    assert abs(pub["loss"].item() - fx["loss_as_published"]) < 1e-5
    out = restate.forward_step(sd, cfg, batch, distill=distill, zero_masks_like_reference=False)
    assert abs(out["text_loss"].item() - fx["text_loss"]) < 1e-5
    assert abs(out["loss"].item() - fx["loss_live"]) < 2e-5, (out["loss"].item(), fx["loss_live"])
    assert torch.allclose(out["logits"][:, ::16, ::8], fx["logits_sub"], atol=2e-5)
    for a, b in zip(out["hidden_states"], fx["hidden_sub"]):
        assert torch.allclose(a[..., ::32, ::8], b, atol=3e-5)
    if distill:
        for task in ("depth", "seg", "gen"):
            for (l, s1, c), (rl, rs1, rc) in zip(out[f"{task}_losses"], fx["emb_losses_live"][task]):
                assert abs(l.item() - rl) < 1e-5 and abs(s1.item() - rs1) < 1e-5 and abs(c.item() - rc) < 1e-5
    out["loss"].backward()
    for n, g in fx["grads"].items():
        mine = sd[n].grad
        if g["norm"] == 0.0:
            assert mine is None or mine.norm().item() < 1e-7, n
            continue
        assert abs(mine.norm().item() - g["norm"]) <= 1e-4 * g["norm"] + 1e-7, n
        assert torch.allclose(mine.flatten()[:16], g["head"], atol=1e-5 + 1e-4 * g["norm"]), n


def test_depth_heads_linear_2_3_get_no_gradient():
    """Reference quirk (Appendix A): only features[0] is supervised."""
    fx = torch.load(GOLDEN / "tiny_llama_dsg.pt")
    live = {n for n, g in fx["grads"].items() if g["norm"] > 0.0}
    assert not any("linear_2" in n or "linear_3" in n for n in live)
    assert any("linear_1" in n for n in live)
    assert any("linear_2" in n for n in fx["state_spec"])


@pytest.mark.parametrize("name,cfg_name,distill", [c for c in CASES if c[0].startswith("tiny") and c[0] not in ("tiny_llama_dsg_padded", "tiny_llama_ntp_mixed")])
def test_state_dict_abi(name, cfg_name, distill):
    from parity_utils import build_product

    fx = torch.load(GOLDEN / f"{name}.pt")
    model = build_product(getattr(configs, cfg_name), distill, None)
    mine = {n: tuple(p.shape) for n, p in model.named_parameters() if "da_v2_head" not in n}
    assert mine == fx["state_spec"]


def test_dpt_head_state_dict_abi_and_oracle():
    """Frozen DPT decoder: parameter names/shapes equal the reference DAv2_Head's,
    and the oracle restatement reproduces the reference's depth map (golden from make_golden_dpt)."""
    from oracle.make_golden_dpt import dpt_inputs
    from visper_lm_b200.model.dpt import DAv2_Head

    fx = torch.load(GOLDEN / "dpt_head.pt")
    mine = {"da_v2_head." + n: tuple(p.shape) for n, p in DAv2_Head().named_parameters()}
    assert mine == fx["state_spec"]
    sd = {n: restate.seeded_param(n, s) for n, s in fx["state_spec"].items()}
    with torch.no_grad():
        depth = restate.dav2_head(sd, dpt_inputs(fx["B"], fx["seed"]))
    assert torch.allclose(depth[:, ::7, ::7], fx["depth_sub"], atol=1e-5)
    assert torch.allclose(restate.depth_pred_normalized(depth)[:, ::7, ::7], fx["depth_norm_sub"], atol=1e-5)


def test_oracle_matches_live_reference():
    """Only where the reference tree is mounted (this container): run the unmodified classes."""
    from oracle import ref_shim

    if not ref_shim.available():
        pytest.skip("/root/reference not mounted")
    cfg = configs.TINY_LLAMA
    model = ref_shim.build_reference_model(cfg, "llama", True, seed_fn=restate.seeded_param)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    batch = configs.synthetic_batch(cfg, 2, 40)
    ref_shim.install_synthetic_teachers(model, batch["targets"])
    masks = {k: v.clone() for k, v in batch["masks"].items()}
    with torch.no_grad():
        out = model(input_ids=batch["input_ids"], labels=batch["labels"], attention_mask=batch["attention_mask"],
                    images=batch["images"], pil_images=[None] * 2, depth_mask=masks["depth"],
                    seg_mask=masks["seg"], gen_mask=masks["gen"])
        mine = restate.forward_step(sd, cfg, batch, distill=True, zero_masks_like_reference=True)
    assert abs(out.loss.item() - mine["loss"].item()) < 1e-5
    assert torch.allclose(out.logits, mine["logits"], atol=3e-5)
    for a, b in zip(out.hidden_states, mine["hidden_states"]):
        assert torch.allclose(a, b, atol=3e-5)
    assert torch.allclose(out.depth_embs[0][0][0], mine["depth_embs"][0][0], atol=3e-5)
    assert torch.allclose(out.seg_embs[1], mine["seg_embs"][1], atol=3e-5)
    assert torch.allclose(out.image_embs[0], mine["gen_embs"][0], atol=3e-5)


@pytest.mark.parametrize("fmt,family", [("emb", "llama"), ("expand_emb", "llama"), ("expand_emb", "phi3")])
def test_oracle_ntp_class_with_task_tokens_matches_live_reference(fmt, family):
    """VPT / IFT stages (vpt.sh, finetune.sh → train.py:933-941): the NTP-only class on the config of a distilled
    checkpoint keeps the task tokens — llava_arch.py:251-293 appends the RAW [576, D] parameters for
    task_token_format "emb" and the 8 pooled rows for "expand_emb"."""
    from oracle import ref_shim

    if not ref_shim.available():
        pytest.skip("/root/reference not mounted")
    base = configs.TINY_LLAMA if family == "llama" else configs.TINY_PHI3
    cfg = dict(base, max_pos=2048, tokenizer_model_max_length=2048)   # room for the 1775-token "emb" rows
    model = ref_shim.build_reference_model(cfg, family, False, seed_fn=restate.seeded_param, ntp_task_token_format=fmt)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    assert sd["model.special_depth_tokens"].shape == (576, cfg["hidden"]) and sd["model.special_gen_tokens"].shape[0] == 8
    batch = configs.synthetic_batch(cfg, 2, 40, distill=False, pad_rows=1)
    with torch.no_grad():
        out = model(input_ids=batch["input_ids"], labels=batch["labels"], attention_mask=batch["attention_mask"],
                    images=batch["images"])
        mine = restate.forward_step(sd, cfg, batch, distill=False, ntp_task_token_format=fmt)
    extra = 576 + 576 + 8 if fmt == "emb" else 24
    assert out.logits.shape[1] == 40 - 1 + 576 + extra == mine["logits"].shape[1]
    assert torch.isfinite(out.loss)
    assert abs(out.loss.item() - mine["loss"].item()) < 1e-5
    real = mine["labels"][0] != -100
    assert torch.allclose(out.logits[0][real], mine["logits"][0][real], atol=5e-5)
