"""Checkpoint / resume on CPU: the trainer's save → resume round trip is
bit-identical to an uninterrupted run (world 1 and world-2 gloo), the adapter-only files follow
the reference layout (llava_trainer.py:997-1016, ola_vlm_train.py:228-249) and load back through
the --pretrain_mm_mlp_adapter path (ola_arch.py:139-144).  The optimizer kernels are torch TEST
DOUBLES here (tests/test_dist_gloo.py); what is under test is the host-side state handling."""
import os
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from test_dist_gloo import _install_test_doubles
from visper_lm_b200.train import checkpoint as C
from visper_lm_b200.train.data import DataCollatorForSupervisedDataset, SyntheticSupervisedDataset


class ToyVLM(torch.nn.Module):
    """Parameter names of the real model's trainable PT-stage set; a differentiable toy forward."""

    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(3)
        mk = lambda *s: torch.nn.Parameter((0.1 * torch.randn(*s, generator=g)).to(torch.bfloat16))
        self.model = torch.nn.Module()
        self.model.mm_projector = torch.nn.Module()
        l0, l2 = torch.nn.Module(), torch.nn.Module()
        l0.weight, l0.bias = mk(16, 12), mk(16)
        l2.weight, l2.bias = mk(16, 16), mk(16)
        self.model.mm_projector.add_module("0", l0)
        self.model.mm_projector.add_module("2", l2)
        self.model.embed_tokens = torch.nn.Module()
        self.model.embed_tokens.weight = torch.nn.Parameter(mk(300, 16).data, requires_grad=False)
        self.depth_logit_scale = torch.nn.Parameter(torch.tensor(2.0))
        self.config = types.SimpleNamespace(model_type="ola_llama", hidden_size=16, mm_hidden_size=12)

    @property
    def device(self):
        return torch.device("cpu")

    def forward(self, input_ids=None, labels=None, attention_mask=None, images=None, **kw):
        pj = self.model.mm_projector
        x = images.float().mean((2, 3))[:, :, None].expand(-1, -1, 4).reshape(images.shape[0], 12)
        l0, l2 = getattr(pj, "0"), getattr(pj, "2")
        h = torch.nn.functional.gelu(x @ l0.weight.float().t() + l0.bias.float())
        h = h @ l2.weight.float().t() + l2.bias.float()
        e = self.model.embed_tokens.weight.float()[input_ids.clamp_min(0)]
        loss = ((e.mean(1) - h) ** 2).mean() * self.depth_logit_scale.float().exp()
        return types.SimpleNamespace(loss=loss)


def _make_trainer(out_dir, max_steps, save_steps, **kw):
    from visper_lm_b200.train.trainer import LLaVATrainer, TrainingArguments

    torch.manual_seed(0)
    ds = SyntheticSupervisedDataset(64, vocab=300, n_sys=13, min_text=20, max_text=40, image_size=8,
                                    distill=False, text_only_every=4, seed=5)
    tok = types.SimpleNamespace(pad_token_id=0, model_max_length=64)
    args = TrainingArguments(output_dir=str(out_dir), per_device_train_batch_size=4, learning_rate=1e-2,
                             max_steps=max_steps, save_steps=save_steps, tune_mm_mlp_adapter=True,
                             group_by_modality_length=True, logging_steps=1, **kw)
    return LLaVATrainer(model=ToyVLM(), args=args, train_dataset=ds, data_collator=DataCollatorForSupervisedDataset(tok))


class _Interrupted(Exception):
    pass


def _train_until(tr, n_steps):
    """Run `tr.train()` and kill it when step n_steps+1 starts (a crash after checkpoint-n)."""
    real, real_acc = tr.step, tr.accumulated_step

    def step(batch):
        if tr.state["global_step"] >= n_steps:
            raise _Interrupted()
        return real(batch)

    def acc_step(micro):
        if tr.state["global_step"] >= n_steps:
            raise _Interrupted()
        return real_acc(micro)

    tr.step, tr.accumulated_step = step, acc_step
    with pytest.raises(_Interrupted):
        tr.train()


def _params(tr):
    return {n: p.detach().float().clone() for n, p in tr.model.named_parameters()}


def test_resume_is_bit_identical_and_layout(tmp_path):
    _install_test_doubles()
    full = _make_trainer(tmp_path / "a", 6, 3, save_total_limit=1)
    full.train()
    # rotation kept only the newest checkpoint
    assert [os.path.basename(p) for p in C.list_checkpoints(str(tmp_path / "a"))] == ["checkpoint-6"]
    part = _make_trainer(tmp_path / "b", 6, 3)
    _train_until(part, 3)
    ck = C.get_last_checkpoint(str(tmp_path / "b"))
    assert os.path.basename(ck) == "checkpoint-3"
    assert sorted(os.listdir(ck)) == ["config.json", "mm_projector.bin", "rng_state.pth", "trainable.bin",
                                      "trainer_state.json", "zero2_rank0_of1.pt"]
    adapter = torch.load(os.path.join(ck, "mm_projector.bin"))
    assert sorted(adapter) == ["model.mm_projector.0.bias", "model.mm_projector.0.weight",
                               "model.mm_projector.2.bias", "model.mm_projector.2.weight"]
    resumed = _make_trainer(tmp_path / "b", 6, 3)
    resumed.train(resume_from_checkpoint=True)  # ola_vlm_train.py:1306-1309
    assert resumed.state["global_step"] == 6
    a, b = _params(full), _params(resumed)
    for n in a:
        assert torch.equal(a[n], b[n]), n
    assert [h["loss"] for h in full.state["log_history"]] == [h["loss"] for h in resumed.state["log_history"]]


def test_final_adapter_save_and_pretrain_load(tmp_path):
    _install_test_doubles()
    tr = _make_trainer(tmp_path / "run", 2, 0)
    tr.train()
    C.safe_save_model_for_hf_trainer(tr, str(tmp_path / "run"))
    assert os.path.exists(tmp_path / "run" / "mm_projector.bin") and os.path.exists(tmp_path / "run" / "config.json")
    C.safe_save_model_for_hf_trainer(tr, str(tmp_path / "run" / "checkpoint-2"))
    assert os.path.exists(tmp_path / "run" / "mm_projector" / "checkpoint-2.bin")
    fresh = ToyVLM()
    C.load_mm_projector(fresh, str(tmp_path / "run" / "mm_projector.bin"))
    for (n, p), (_, q) in zip(fresh.model.mm_projector.named_parameters(), tr.model.model.mm_projector.named_parameters()):
        assert torch.equal(p, q), n
    # full (non-adapter) branch goes through trainer._save
    tr.args.tune_mm_mlp_adapter = False
    C.safe_save_model_for_hf_trainer(tr, str(tmp_path / "full"))
    assert os.path.exists(tmp_path / "full" / "model.safetensors")         # HF save_pretrained layout
    sd = C.load_pretrained_weights(str(tmp_path / "full"))
    assert "model.embed_tokens.weight" in sd and "depth_logit_scale" in sd
    for k, v in tr.model.state_dict().items():
        assert torch.equal(sd[k], v.detach().cpu()), k


def test_sharded_safetensors_layout_equals_hf_save_pretrained(tmp_path):
    """save_pretrained_weights writes what transformers' own save_pretrained writes for the same state dict
    and shard size (file names, index weight_map / total_size, tensors), and HF from_pretrained loads it."""
    import json

    from transformers import LlamaConfig, LlamaForCausalLM

    torch.manual_seed(0)
    hf = LlamaForCausalLM(LlamaConfig(vocab_size=160, hidden_size=32, intermediate_size=64, num_hidden_layers=3,
                                      num_attention_heads=4, num_key_value_heads=2, tie_word_embeddings=False))
    hf.save_pretrained(tmp_path / "hf", max_shard_size="20KB", safe_serialization=True)
    files = C.save_pretrained_weights(hf.state_dict(), str(tmp_path / "mine"), max_shard_size="20KB")
    ref_files = sorted(f for f in os.listdir(tmp_path / "hf") if f.endswith(".safetensors"))
    assert files == ref_files and len(files) > 2
    a = json.load(open(tmp_path / "hf" / C.SAFE_WEIGHTS_INDEX_NAME))
    b = json.load(open(tmp_path / "mine" / C.SAFE_WEIGHTS_INDEX_NAME))
    assert a["weight_map"] == b["weight_map"] and a["metadata"]["total_size"] == b["metadata"]["total_size"]
    back = C.load_pretrained_weights(str(tmp_path / "mine"))
    for k, v in hf.state_dict().items():
        assert torch.equal(back[k], v), k
    hf.config.save_pretrained(tmp_path / "mine")
    again = LlamaForCausalLM.from_pretrained(tmp_path / "mine")
    for (k, v), (_, w) in zip(hf.state_dict().items(), again.state_dict().items()):
        assert torch.equal(v, w), k
    one = C.save_pretrained_weights(hf.state_dict(), str(tmp_path / "one"))   # under the limit: single file, no index
    assert one == ["model.safetensors"] and not os.path.exists(tmp_path / "one" / C.SAFE_WEIGHTS_INDEX_NAME)


def _worker(rank, world, port, root, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _install_test_doubles()
    full = _make_trainer(os.path.join(root, "a"), 4, 2)
    full.train()
    part = _make_trainer(os.path.join(root, "b"), 4, 2)
    _train_until(part, 2)
    dist.barrier()
    resumed = _make_trainer(os.path.join(root, "b"), 4, 2)
    resumed.train(resume_from_checkpoint=True)
    a, b = _params(full), _params(resumed)
    ret[rank] = all(torch.equal(a[n], b[n]) for n in a) and os.path.exists(
        os.path.join(root, "b", "checkpoint-2", f"zero2_rank{rank}_of{world}.pt"))
    dist.destroy_process_group()


def test_resume_world2_gloo(tmp_path):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29547, str(tmp_path), ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def test_full_save_round_trips_through_from_pretrained(tmp_path):
    """config.json + sharded safetensors written the way trainer._save does → from_pretrained rebuilds the
    same model (heads, task tokens, logit scales, DPT decoder included)."""
    from parity_utils import build_product, configs
    from visper_lm_b200.model import OlaLlavaLlamaForCausalLM

    torch.manual_seed(3)
    model = build_product(configs.TINY_LLAMA, True, None)
    with torch.no_grad():
        for p in model.parameters():
            p.copy_(torch.randn(p.shape))
    C.save_config(model.config, str(tmp_path))
    files = C.save_pretrained_weights(model.state_dict(), str(tmp_path), max_shard_size="300KB")
    assert len(files) > 1
    again = OlaLlavaLlamaForCausalLM.from_pretrained(str(tmp_path))
    a, b = model.state_dict(), again.state_dict()
    assert list(a) == list(b)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    assert again.config.aux_mode == model.config.aux_mode and again.depth_layer_indices == model.depth_layer_indices
    os.remove(tmp_path / files[0])
    import pytest
    with pytest.raises(Exception):
        OlaLlavaLlamaForCausalLM.from_pretrained(str(tmp_path))


def test_prefetching_workers_keep_batch_order_and_results(tmp_path):
    """dataloader_num_workers > 0: batches are built on host threads ahead of the step, delivered in the
    same order — the trained parameters are bit-identical to the synchronous loader's, also after a resume
    in the middle of an epoch."""
    _install_test_doubles()
    sync = _make_trainer(tmp_path / "s", 6, 0)
    sync.train()
    pre = _make_trainer(tmp_path / "p", 6, 3, dataloader_num_workers=3)
    a = [b["input_ids"].clone() for b in pre._batches(0, 2)]
    b = [b["input_ids"].clone() for b in sync._batches(0, 2)]
    assert len(a) == len(b) > 4 and all(torch.equal(x, y) for x, y in zip(a, b))
    _train_until(pre, 3)
    resumed = _make_trainer(tmp_path / "p", 6, 3, dataloader_num_workers=2)
    resumed.train(resume_from_checkpoint=True)
    pa, pb = _params(sync), _params(resumed)
    for n in pa:
        assert torch.equal(pa[n], pb[n]), n
    it = pre._batches(0, 0)   # abandoning the iterator early must not hang or leak the pool
    next(it)
    it.close()


def test_loads_plain_hf_llama_and_clip_checkpoints(tmp_path):
    """A stock HF Llama checkpoint (sharded safetensors) loads into the product's decoder by name, the
    multimodal modules keep their fresh init; a stock HF CLIP vision checkpoint loads into the tower."""
    from transformers import CLIPVisionConfig, CLIPVisionModel, LlamaConfig, LlamaForCausalLM

    from visper_lm_b200.model import LlavaLlamaForCausalLM

    torch.manual_seed(1)
    hf = LlamaForCausalLM(LlamaConfig(vocab_size=160, hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                                      num_attention_heads=4, num_key_value_heads=2, tie_word_embeddings=False,
                                      max_position_embeddings=256, rope_theta=500000.0))
    hf.save_pretrained(tmp_path / "llm", max_shard_size="60KB", safe_serialization=True)
    vis = dict(hidden_size=32, intermediate_size=64, num_hidden_layers=3, num_attention_heads=2, image_size=28, patch_size=14)
    model = LlavaLlamaForCausalLM.from_pretrained(str(tmp_path / "llm"), vision=vis)
    own = model.state_dict()
    for k, v in hf.state_dict().items():
        assert torch.equal(own[k].float(), v.to(torch.bfloat16).float()), k
    assert model.config.num_key_value_heads == 2 and model.config.rope_theta == 500000.0
    clip = CLIPVisionModel(CLIPVisionConfig(**vis))
    clip.save_pretrained(tmp_path / "clip", safe_serialization=True)
    tower = model.get_vision_tower()
    tower.load_model(path=str(tmp_path / "clip"))
    for k, v in clip.state_dict().items():
        if k in tower.vision_tower.state_dict():
            assert torch.equal(tower.vision_tower.state_dict()[k].float(), v.to(torch.bfloat16).float()), k
    assert tower.is_loaded and not any(p.requires_grad for p in tower.parameters())
    import pytest
    with pytest.raises(KeyError):
        tower.load_model(path=str(tmp_path / "llm"))


def test_distill_model_on_top_of_a_plain_llm_checkpoint(tmp_path):
    """The train() recipe (ola_vlm_train.py:1007-1021,1149-1266): base LLM from a stock checkpoint, aux
    config injected, task tokens + heads created and initialised without touching the loaded weights."""
    from transformers import LlamaConfig, LlamaForCausalLM

    from visper_lm_b200.model import OlaLlavaLlamaForCausalLM

    hf = LlamaForCausalLM(LlamaConfig(vocab_size=160, hidden_size=64, intermediate_size=128, num_hidden_layers=4,
                                      num_attention_heads=4, num_key_value_heads=2, tie_word_embeddings=False))
    hf.save_pretrained(tmp_path, safe_serialization=True)
    vis = dict(hidden_size=32, intermediate_size=64, num_hidden_layers=3, num_attention_heads=2, image_size=28, patch_size=14)
    m = OlaLlavaLlamaForCausalLM.from_pretrained(str(tmp_path), vision=vis)
    assert not hasattr(m, "image_depth_heads")
    m.config.inject_aux(mode="gen-depth-seg", layer_indices="d2-3_s1-2_g2-3", num_task_tokens=8, gen_dim=32,
                        seg_dim=48, depth_dim=32)
    m.get_model().initialize_special_tokens(m.config)
    m.init_heads(m.config)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if any(t in n for t in m.NEW_MODULE_KEYS):
                p.fill_(float("nan"))
    m.init_weights(seed=3, only=m.NEW_MODULE_KEYS)
    sd = m.state_dict()
    assert all(torch.isfinite(v.float()).all() for v in sd.values())
    for k, v in hf.state_dict().items():
        assert torch.equal(sd[k].float(), v.to(torch.bfloat16).float()), k       # loaded LLM untouched
    assert m.depth_layer_indices == [1, 2] and float(m.depth_logit_scale) == 2.0
    assert abs(float(m.model.special_depth_tokens.float().std()) - 1.0) < 0.1    # randn task tokens (ola_arch.py:77-93)


def test_gradient_accumulation_equals_the_big_batch(tmp_path):
    """gradient_accumulation_steps=2 at per-device batch 2 takes the same optimizer steps as batch 4 without
    accumulation: every step covers the same length-grouped mega-batch, the losses are averaged over the
    micro-batches (HF Trainer semantics) — parameters agree to bf16 gradient rounding; resume works."""
    from visper_lm_b200.train.trainer import LLaVATrainer, TrainingArguments

    _install_test_doubles()

    def make(out, B, ga, **kw):
        torch.manual_seed(0)
        # equal text lengths: the toy loss averages over the padded length, so padding must not differ
        ds = SyntheticSupervisedDataset(64, vocab=300, n_sys=13, min_text=30, max_text=30, image_size=8,
                                        distill=False, text_only_every=4, seed=5)
        tok = types.SimpleNamespace(pad_token_id=0, model_max_length=64)
        args = TrainingArguments(output_dir=str(out), per_device_train_batch_size=B, gradient_accumulation_steps=ga,
                                 learning_rate=1e-2, max_steps=6, save_steps=3, tune_mm_mlp_adapter=True,
                                 group_by_modality_length=True, logging_steps=1, **kw)
        return LLaVATrainer(model=ToyVLM(), args=args, train_dataset=ds, data_collator=DataCollatorForSupervisedDataset(tok))

    big = make(tmp_path / "big", 4, 1)
    acc = make(tmp_path / "acc", 2, 2)
    assert big.steps_per_epoch() == acc.steps_per_epoch() == 16
    for k in range(3):   # each optimizer step sees the same 4 samples
        a = sorted(big._index_order(0)[k * 4:(k + 1) * 4])
        b = sorted(acc._index_order(0)[k * 4:(k + 1) * 4])
        assert a == b
    big.train()
    acc.train()
    assert acc.state["global_step"] == big.state["global_step"] == 6
    pa, pb = _params(big), _params(acc)
    for n in pa:
        assert torch.allclose(pa[n].float(), pb[n].float(), atol=2e-2, rtol=5e-2), n
    la, lb = [h["loss"] for h in big.state["log_history"]], [h["loss"] for h in acc.state["log_history"]]
    assert all(abs(x - y) <= 0.05 * abs(x) + 1e-3 for x, y in zip(la, lb))
    part = make(tmp_path / "acc2", 2, 2)
    _train_until(part, 3)
    resumed = make(tmp_path / "acc2", 2, 2)
    resumed.train(resume_from_checkpoint=True)
    pc = _params(resumed)
    for n in pb:
        assert torch.equal(pb[n], pc[n]), n
