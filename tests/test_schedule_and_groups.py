"""Optimizer-side host logic against the libraries the reference's Trainer uses: the cosine schedule with
linear warm-up (HF get_cosine_schedule_with_warmup, warmup_ratio 0.03 as in pretrain.sh:47-48) and the
decay / no-decay × projector-lr parameter grouping of LLaVATrainer.create_optimizer
(llava_trainer.py:903-976)."""
import math

import pytest
import torch

from visper_lm_b200.train.trainer import _no_decay, cosine_with_warmup


@pytest.mark.parametrize("total,ratio", [(100, 0.03), (2180, 0.03), (7, 0.5), (50, 0.0)])
def test_cosine_with_warmup_equals_hf(total, ratio):
    from transformers import get_cosine_schedule_with_warmup

    warm = math.ceil(total * ratio)
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=1.0)
    sched = get_cosine_schedule_with_warmup(opt, num_warmup_steps=warm, num_training_steps=total)
    for step in range(total):
        want = opt.param_groups[0]["lr"]
        assert abs(cosine_with_warmup(step, total, warm) - want) < 1e-12, step
        opt.step()
        sched.step()


def test_no_decay_names_match_hf_rule():
    """HF: decay parameters = all except LayerNorm weights and anything named *bias
    (get_parameter_names(model, ALL_LAYERNORM_LAYERS) minus 'bias')."""
    from transformers.trainer_pt_utils import get_parameter_names

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.proj = torch.nn.Linear(4, 4)
            self.norm_out = torch.nn.LayerNorm(4)
            self.layers = torch.nn.ModuleList([torch.nn.Sequential(torch.nn.LayerNorm(4), torch.nn.Linear(4, 4, bias=False))])
            self.logit_scale = torch.nn.Parameter(torch.tensor(2.0))

    m = Tiny()
    decay = [n for n in get_parameter_names(m, [torch.nn.LayerNorm]) if "bias" not in n]
    for n, _ in m.named_parameters():
        hf_no_decay = n not in decay
        if "layers.0.0" in n:   # an unnamed LayerNorm inside a Sequential: only the module TYPE identifies it
            continue            # (the product's norms all carry 'norm' in their reference names)
        assert _no_decay(n) == hf_no_decay, n


def test_trainer_sampler_and_compute_loss_surface():
    """LLaVATrainer._get_train_sampler (llava_trainer.py:219-232) and HF's compute_loss contract."""
    import torch

    from visper_lm_b200.train.data import LengthGroupedSampler
    from visper_lm_b200.train.trainer import LLaVATrainer, TrainingArguments

    class DS:
        modality_lengths = [5, -3, 7, 9, -2, 4, 8, 6]

        def __len__(self):
            return len(self.modality_lengths)

    t = LLaVATrainer(model=None, args=TrainingArguments(per_device_train_batch_size=2, group_by_modality_length=True,
                                                        gradient_accumulation_steps=2), train_dataset=DS())
    s = t._get_train_sampler()
    assert isinstance(s, LengthGroupedSampler) and (s.batch_size, s.world_size, s.group_by_modality) == (2, 2, True)
    assert sorted(s) == list(range(8))
    t.args.group_by_modality_length = False
    assert sorted(t._get_train_sampler()) == list(range(8))
    assert LLaVATrainer(model=None, args=TrainingArguments())._get_train_sampler() is None

    class Out:
        loss = torch.tensor(1.5)

    class M:
        device = torch.device("cpu")

        def __call__(self, **kw):
            return (torch.tensor(2.5), None) if kw.get("return_dict") is False else Out()

    t = LLaVATrainer(model=M(), args=TrainingArguments())
    assert float(t.compute_loss(t.model, {"x": torch.zeros(1)})) == 1.5
    loss, out = t.compute_loss(t.model, {"return_dict": False}, return_outputs=True)
    assert float(loss) == 2.5 and isinstance(out, tuple)
