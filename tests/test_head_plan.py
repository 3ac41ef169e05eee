"""HeadPlan (the gather indices behind forward_emb_predictor) against the reference's own
method (base_ola_vlm.py:413-443, extracted from the source and run on a stub) over random shapes / orders
/ flags, including the `T < 600` fallback."""
from types import SimpleNamespace

import pytest
import torch

from parity_utils import ROOT  # noqa: F401
from oracle import ref_functions, ref_shim
from visper_lm_b200.model.vlm import HeadPlan

ORDERS = [("gen", "depth", "seg"), ("depth", "seg", "gen"), ("seg", "gen", "depth"), ("depth",), ("gen", "seg")]


def _reference(B, T, S, nt, order, task, pass_text, state, special):
    fn = ref_functions.extract_method("ola_vlm/model/language_model/base_ola_vlm.py", "BaseOLA_VLM",
                                      "forward_emb_predictor")
    captured = {}

    def head(inp, task_tokens=None):
        captured["inp"], captured["lat"] = inp, task_tokens
        return inp

    stub = SimpleNamespace(token_order=list(order), NUM_SYS_TOKENS=S, num_task_tokens=nt, pass_text_to_aux_head=pass_text)
    fn(stub, [state], 0, 0, task, [head], special)
    return captured["inp"], captured["lat"]


@pytest.mark.parametrize("case", range(40))
def test_head_plan_equals_reference_selection(case):
    if not ref_shim.available():
        pytest.skip("/root/reference not mounted")
    g = torch.Generator().manual_seed(case)
    order = ORDERS[case % len(ORDERS)]
    task = order[int(torch.randint(0, len(order), (1,), generator=g))]
    S = [13, 26, 38][case % 3]
    nt = 8
    pass_text = case % 4 != 3
    B = 1 + case % 3
    small = case % 5 == 4                                   # the T < 600 fallback branch
    T = int(torch.randint(S + 4, 599, (1,), generator=g)) if small else \
        S + 576 + nt * len(order) + int(torch.randint(0, 50, (1,), generator=g))
    D = 4
    state = torch.randn(B, T, D, generator=g)
    n_lat = 1 if task == "gen" else 16
    special = torch.randn(nt if task == "gen" else n_lat, D, generator=g)
    if small and task == "gen" and T < S + 576 + nt:
        pytest.skip("reference slices past the sequence end here")
    ref_inp, ref_lat = _reference(B, T, S, nt, order, task, pass_text, state, special)
    from oracle import restate
    o_inp, o_lat = restate.head_inputs({"num_sys_tokens": S, "num_task_tokens": nt, "aux_mode": "-".join(order),
                                        "pass_text_to_aux": pass_text}, state, task, special)
    assert torch.equal(o_inp, ref_inp) and torch.equal(o_lat, ref_lat)               # the oracle, too
    plan = HeadPlan(B, T, S, nt, list(order), task, pass_text, n_lat, 4, "cpu")
    flat = state.reshape(B * T, D)
    mine_inp = flat[plan.ctx_index.long()].view(B, plan.nk, D)
    assert mine_inp.shape == ref_inp.shape and torch.equal(mine_inp, ref_inp)
    inv = plan.inv_ctx.long()
    assert torch.equal(inv[plan.ctx_index.long()], torch.arange(B * plan.nk))      # inverse map for the backward
    if task == "gen":
        mine_lat = flat[plan.gen_index.long()].view(B, -1, D)
        assert mine_lat.shape == ref_lat.shape and torch.equal(mine_lat, ref_lat)
    else:
        assert ref_lat.shape == (B, n_lat, D) and torch.equal(ref_lat[B - 1], special)
        assert plan.lat_index.view(B, n_lat).tolist() == [list(range(n_lat))] * B
