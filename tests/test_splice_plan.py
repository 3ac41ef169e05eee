"""Property test (CPU): the host-built SplicePlan behind prepare_inputs_labels_for_multimodal equals the
oracle's restatement of ola_arch.py:337-444 on random batches — any number of images per row (0, 1,
several), right padding, truncation to tokenizer_model_max_length, with and without task tokens."""
import numpy as np
import pytest
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from parity_utils import restate
from visper_lm_b200.model.vlm import IGNORE_INDEX, IMAGE_TOKEN_INDEX, SplicePlan

D, V, NIMG = 8, 50, 6  # tiny widths: 6 "image tokens" per image


@st.composite
def batches(draw):
    B = draw(st.integers(1, 4))
    N = draw(st.integers(3, 24))
    task_rows = draw(st.sampled_from([0, 3, 17]))   # 17 > the image rows: the NTP-only "emb" case appends more task rows than image rows
    max_len = draw(st.sampled_from([None, 20, 64]))
    rng = np.random.default_rng(draw(st.integers(0, 2**31 - 1)))
    ids = rng.integers(0, V, (B, N))
    am = np.ones((B, N), bool)
    n_images = 0
    for b in range(B):
        length = int(rng.integers(1, N + 1))
        am[b, length:] = False
        k = int(rng.integers(0, 3))
        for p in rng.choice(length, size=min(k, length), replace=False):
            ids[b, p] = IMAGE_TOKEN_INDEX
        n_images += max(1, int((ids[b, :length] == IMAGE_TOKEN_INDEX).sum()))  # text-only rows use a slot too
    labels = np.where(rng.random((B, N)) < 0.3, IGNORE_INDEX, ids)
    return ids, labels, am, n_images, task_rows, max_len, int(rng.integers(0, 2**31 - 1))


@settings(max_examples=150, deadline=None)
@given(batches())
def test_splice_plan_matches_oracle(case):
    ids, labels, am, n_images, task_rows, max_len, seed = case
    g = torch.Generator().manual_seed(seed)
    embed = torch.randn(V, D, generator=g)
    img = torch.randn(n_images, NIMG, D, generator=g)
    task = torch.randn(task_rows, D, generator=g)
    plan = SplicePlan(torch.from_numpy(ids), torch.from_numpy(labels), torch.from_numpy(am), NIMG, task_rows,
                      max_len, "right")
    # materialise the plan with plain indexing (the CUDA gather does exactly this)
    srcs = [embed, img.reshape(-1, D), task]
    kind, index = plan.np["kind"], plan.np["index"]
    rows = torch.zeros(len(kind), D)
    for r, (k, i) in enumerate(zip(kind, index)):
        if k >= 0:
            rows[r] = srcs[k][i]
    mine = rows.view(plan.B, plan.T, D)

    # oracle: restate.splice with the pooled task rows injected
    sd = {"model.embed_tokens.weight": embed}
    cfg = {"num_task_tokens": 0, "tokenizer_model_max_length": max_len, "aux_mode": ""}
    orig = restate.pooled_task_tokens
    restate.pooled_task_tokens = lambda sd_, cfg_: [task] if task_rows else []
    try:
        want, want_lab, want_mask = restate.splice(sd, cfg, torch.from_numpy(ids), torch.from_numpy(labels),
                                                   torch.from_numpy(am), img)
    finally:
        restate.pooled_task_tokens = orig
    assert mine.shape == want.shape
    assert torch.equal(mine, want)
    assert torch.equal(torch.from_numpy(plan.np["labels"]), want_lab)
    assert torch.equal(torch.from_numpy(plan.np["mask"]), want_mask)
    # the inverse maps used by the backward are consistent with the forward plan
    inv = plan.np["inv_img"]
    for slot, r in enumerate(inv):
        if r >= 0:
            assert kind[r] == 1 and index[r] == slot
    valid = plan.np["ce_rows"]
    lab = plan.np["labels"]
    shifted = np.full_like(lab, IGNORE_INDEX)
    shifted[:, :-1] = lab[:, 1:]
    assert np.array_equal(np.nonzero(shifted.reshape(-1) != IGNORE_INDEX)[0], valid)
    assert np.array_equal(plan.np["ce_targets"], shifted.reshape(-1)[valid])
