"""Checkpoint → resume with GPU-resident state: the real tiny model on cuda:0, the CUDA AdamW /
grad-norm kernels, fp32 master / moment shards on the device.  A run interrupted after checkpoint-3 and resumed
must end where the uninterrupted run ends: same step counter, same data order, same LR schedule, same optimizer
state.  (Bias gradients are reduced with fp32 atomics, so equality is to ~1e-6 rather than bit for bit; a lost
moment, a repeated batch or a shifted schedule would show at the 1e-3 level after three more steps at lr 1e-2.)"""
import os
import types

import pytest
import torch

from parity_utils import build_product, configs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class _Interrupted(Exception):
    pass


def _trainer(out_dir, max_steps, save_steps):
    from visper_lm_b200.train.data import DataCollatorForSupervisedDataset, SyntheticSupervisedDataset
    from visper_lm_b200.train.trainer import LLaVATrainer, TrainingArguments

    cfg = configs.TINY_LLAMA
    model = build_product(cfg, False, DEV)
    for n, p in model.named_parameters():
        p.requires_grad_("mm_projector" in n)
    ds = SyntheticSupervisedDataset(48, vocab=cfg["vocab"], n_sys=cfg["num_sys_tokens"], min_text=40, max_text=56,
                                    image_size=336, distill=False, seed=11)
    tok = types.SimpleNamespace(pad_token_id=cfg["vocab"] - 1, model_max_length=1024)
    args = TrainingArguments(output_dir=str(out_dir), per_device_train_batch_size=2, learning_rate=1e-2,
                             max_steps=max_steps, save_steps=save_steps, tune_mm_mlp_adapter=True, logging_steps=1,
                             warmup_ratio=0.0)
    return LLaVATrainer(model=model, args=args, train_dataset=ds, data_collator=DataCollatorForSupervisedDataset(tok))


def test_resume_on_gpu_matches_uninterrupted_run(tmp_path):
    from visper_lm_b200.train import checkpoint as C

    full = _trainer(tmp_path / "a", 6, 3)
    full.train()
    part = _trainer(tmp_path / "b", 6, 3)
    real = part.step

    def step(batch):
        if part.state["global_step"] >= 3:
            raise _Interrupted()
        return real(batch)

    part.step = step
    with pytest.raises(_Interrupted):
        part.train()
    ck = C.get_last_checkpoint(str(tmp_path / "b"))
    assert os.path.basename(ck) == "checkpoint-3" and os.path.exists(os.path.join(ck, "zero2_rank0_of1.pt"))
    resumed = _trainer(tmp_path / "b", 6, 3)
    resumed.train(resume_from_checkpoint=True)
    torch.cuda.synchronize()
    assert resumed.state["global_step"] == full.state["global_step"] == 6
    assert resumed.optimizer.step_count == full.optimizer.step_count == 6
    assert resumed.optimizer.master.is_cuda and resumed.optimizer.m.is_cuda
    for name in ("master", "m", "v"):
        a, b = getattr(full.optimizer, name), getattr(resumed.optimizer, name)
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-6), f"optimizer {name} differs after resume: {(a - b).abs().max().item()}"
    la = [h["loss"] for h in full.state["log_history"]]
    lb = [h["loss"] for h in resumed.state["log_history"]]
    assert len(la) == len(lb) == 6 and all(abs(x - y) <= 1e-4 * abs(x) for x, y in zip(la, lb)), (la, lb)
    moved = (full.optimizer.master - _trainer(tmp_path / "c", 6, 0).create_optimizer().master).abs().max().item()
    assert moved > 1e-3, "the run did not train (test would be vacuous)"
