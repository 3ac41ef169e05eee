"""GPU parity of the VPT / IFT configuration (scripts/train/vpt.sh, finetune.sh → train.py): the NTP-only class on
the config of a distilled checkpoint splices the task tokens — pooled for task_token_format "expand_emb", raw
(576 + 576 + 8 rows) for "emb" (llava_arch.py:251-293).  Oracle pinned to the live reference classes in
tests/test_oracle_cpu.py::test_oracle_ntp_class_with_task_tokens_matches_live_reference."""
import pytest
import torch

from parity_utils import bf16_seeded, configs, cos_sim, oracle_state, restate, round_batch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fmt", ["expand_emb", "emb"])
def test_ntp_step_with_task_tokens_vs_oracle(fmt):
    from visper_lm_b200.model import LlavaLlamaForCausalLM, presets

    c = dict(configs.TINY_LLAMA, max_pos=2048, tokenizer_model_max_length=2048)
    cfg = presets.from_dict(c, distill=True)            # a distilled checkpoint's config: aux keys present
    cfg.task_token_format = fmt
    model = LlavaLlamaForCausalLM(cfg, device="cuda:0")
    model.init_weights(seed_fn=bf16_seeded)
    for n, p in model.named_parameters():                # finetune.sh regime; task tokens frozen (train.py:65)
        p.requires_grad_(("vision_tower" not in n) and ("model.special_" not in n))
    batch = round_batch(configs.synthetic_batch(c, 2, 40, seed=77, distill=False, pad_rows=1))
    out = model(input_ids=batch["input_ids"], labels=batch["labels"], attention_mask=batch["attention_mask"],
                images=batch["images"].to("cuda:0"))
    out.loss.backward()
    torch.cuda.synchronize()
    sd = oracle_state(model)
    req = {n: sd[n].clone().requires_grad_(True) for n in ("model.mm_projector.2.weight", "model.layers.0.mlp.down_proj.weight")}
    ref = restate.forward_step({**sd, **req}, c, batch, distill=False, ntp_task_token_format=fmt)
    ref["loss"].backward()
    extra = 24 if fmt == "expand_emb" else 576 + 576 + 8
    assert ref["logits"].shape[1] == 40 - 1 + 576 + extra
    got, want = out.loss.item(), ref["loss"].item()
    print(f"ntp + task tokens ({fmt}): loss {got:.6f} oracle {want:.6f} rel {abs(got - want) / abs(want):.2e}")
    assert abs(got - want) <= 2e-3 * abs(want), (got, want)   # 1e-3 expected (north_star); uncalibrated new case
    params = dict(model.named_parameters())
    for n, r in req.items():
        assert cos_sim(params[n].grad, r.grad) > 0.99, (n, cos_sim(params[n].grad, r.grad))
