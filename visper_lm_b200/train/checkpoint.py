"""Checkpoint / resume for the trainer, in the reference's on-disk layout:

  * adapter-only saves of the PT stage — `checkpoint-N/mm_projector.bin` during training
    (ola_vlm/train/llava_trainer.py:997-1016: keys matching mm_projector / vision_resampler, plus
    embed_tokens / embed_in with --use_im_start_end) and `mm_projector.bin` (or
    `mm_projector/checkpoint-N.bin`) at the end (ola_vlm/train/ola_vlm_train.py:228-249), so
    `--pretrain_mm_mlp_adapter` (ola_arch.py:139-144) can load what this trainer wrote;
  * what a resume needs on top of that, which the reference gets from HF Trainer + DeepSpeed:
    every trainable tensor (`trainable.bin`), each rank's ZeRO-2 shard of fp32 masters / Adam
    moments (`zero2_rank{r}_of{w}.pt`), `trainer_state.json`, RNG state; `save_total_limit`
    rotation and auto-resume from the newest `checkpoint-*` (ola_vlm_train.py:1306-1309).
All host-side; tensors cross to the CPU once per save.
"""
from __future__ import annotations

import json
import os
import re
import shutil
from typing import Dict, Iterable, Optional

import torch

PREFIX_CHECKPOINT_DIR = "checkpoint"


def get_mm_adapter_state(named_params: Iterable, keys_to_match) -> Dict[str, torch.Tensor]:
    """CPU copies of the parameters whose name contains one of `keys_to_match`
    (get_mm_adapter_state_maybe_zero_3 without the ZeRO-3 gather: ZeRO-2 keeps full parameters)."""
    return {k: v.detach().cpu().clone() for k, v in named_params if any(m in k for m in keys_to_match)}


def adapter_keys(args, in_training: bool):
    keys = ["mm_projector", "vision_resampler"] if in_training else ["mm_projector"]
    if getattr(args, "use_im_start_end", False):
        keys += ["embed_tokens", "embed_in"]
    return keys


def config_to_dict(config) -> dict:
    out = {}
    for k, v in vars(config).items():
        try:
            json.dumps(v)
            out[k] = v
        except TypeError:
            out[k] = repr(v)
    out.setdefault("model_type", getattr(config, "model_type", "visper"))
    return out


def save_config(config, output_dir: str):
    os.makedirs(output_dir, exist_ok=True)
    with open(os.path.join(output_dir, "config.json"), "w") as f:
        json.dump(config_to_dict(config), f, indent=2, sort_keys=True)


def list_checkpoints(run_dir: str):
    if not os.path.isdir(run_dir):
        return []
    found = []
    for name in os.listdir(run_dir):
        m = re.fullmatch(rf"{PREFIX_CHECKPOINT_DIR}-(\d+)", name)
        if m and os.path.isdir(os.path.join(run_dir, name)):
            found.append((int(m.group(1)), os.path.join(run_dir, name)))
    return [p for _, p in sorted(found)]


def get_last_checkpoint(run_dir: str) -> Optional[str]:
    cps = list_checkpoints(run_dir)
    return cps[-1] if cps else None


def rotate_checkpoints(run_dir: str, save_total_limit: Optional[int]):
    if not save_total_limit or save_total_limit <= 0:
        return
    cps = list_checkpoints(run_dir)
    for old in cps[: max(0, len(cps) - save_total_limit)]:
        shutil.rmtree(old, ignore_errors=True)


def save_checkpoint(trainer, output_dir: str):
    """One resumable checkpoint of `trainer` into output_dir (every rank writes its optimizer shard;
    rank 0 writes the rest)."""
    os.makedirs(output_dir, exist_ok=True)
    opt = trainer.optimizer
    rank, world = trainer.rank, trainer.world
    if opt is not None:
        opt.wait_params()   # a parameter all-gather of the last step may still be in flight on the comm stream
        torch.save({"step": opt.step_count, "master": opt.master.cpu(), "m": opt.m.cpu(), "v": opt.v.cpu(),
                    "rank": rank, "world": world, "total": opt.total,
                    "names": [n for n, _ in opt.named]},
                   os.path.join(output_dir, f"zero2_rank{rank}_of{world}.pt"))
    if rank == 0:
        if getattr(trainer.args, "tune_mm_mlp_adapter", False):
            save_config(trainer.model.config, output_dir)
            torch.save(get_mm_adapter_state(trainer.model.named_parameters(), adapter_keys(trainer.args, True)),
                       os.path.join(output_dir, "mm_projector.bin"))
        torch.save({k: v.detach().cpu() for k, v in trainer.model.named_parameters() if v.requires_grad},
                   os.path.join(output_dir, "trainable.bin"))
        with open(os.path.join(output_dir, "trainer_state.json"), "w") as f:
            json.dump(trainer.state, f)
        torch.save({"cpu": torch.get_rng_state()}, os.path.join(output_dir, "rng_state.pth"))


def load_checkpoint(trainer, ckpt_dir: str):
    """Restores trainable weights, this rank's optimizer shard, the step counters and the RNG."""
    dev = trainer.model.device
    sd = torch.load(os.path.join(ckpt_dir, "trainable.bin"), map_location="cpu")
    params = dict(trainer.model.named_parameters())
    missing = [k for k in sd if k not in params]
    if missing:
        raise KeyError(f"checkpoint has parameters the model lacks: {missing[:5]}")
    with torch.no_grad():
        for k, v in sd.items():
            params[k].copy_(v.to(dev, params[k].dtype))
    opt = trainer.create_optimizer()
    f = os.path.join(ckpt_dir, f"zero2_rank{trainer.rank}_of{trainer.world}.pt")
    if not os.path.exists(f):
        raise FileNotFoundError(f"{f}: optimizer shards are per (rank, world size); resume with the same world size")
    st = torch.load(f, map_location="cpu")
    if st["total"] != opt.total or st["names"] != [n for n, _ in opt.named]:
        raise ValueError("optimizer shard layout differs from the current trainable set")
    opt.step_count = st["step"]
    opt.master.copy_(st["master"].to(dev))
    opt.m.copy_(st["m"].to(dev))
    opt.v.copy_(st["v"].to(dev))
    with open(os.path.join(ckpt_dir, "trainer_state.json")) as fh:
        trainer.state = json.load(fh)
    rng = os.path.join(ckpt_dir, "rng_state.pth")
    if os.path.exists(rng):
        torch.set_rng_state(torch.load(rng)["cpu"])


def safe_save_model_for_hf_trainer(trainer, output_dir: str):
    """End-of-run save (ola_vlm_train.py:228-263): adapter-only when tune_mm_mlp_adapter, else
    the full state dict through trainer._save."""
    if getattr(trainer.args, "tune_mm_mlp_adapter", False):
        weights = get_mm_adapter_state(trainer.model.named_parameters(), adapter_keys(trainer.args, False))
        save_config(trainer.model.config, output_dir)
        current = os.path.basename(os.path.normpath(output_dir))
        parent = os.path.dirname(os.path.normpath(output_dir))
        if trainer.rank == 0:
            if current.startswith(f"{PREFIX_CHECKPOINT_DIR}-"):
                folder = os.path.join(parent, "mm_projector")
                os.makedirs(folder, exist_ok=True)
                torch.save(weights, os.path.join(folder, f"{current}.bin"))
            else:
                torch.save(weights, os.path.join(output_dir, "mm_projector.bin"))
        # no `return` here in the reference either (ola_vlm_train.py:249-251): under DeepSpeed — every shipped
        # script — it falls through to trainer.save_model(output_dir), so the PT output directory also holds the
        # FULL model (trained heads, task tokens, logit scales) that finetune.sh / vpt.sh load with from_pretrained
    if trainer.rank == 0:
        trainer._save(output_dir, state_dict={k: v.detach().cpu() for k, v in trainer.model.state_dict().items()})


def load_mm_projector(model, path: str):
    """--pretrain_mm_mlp_adapter (ola_arch.py:139-144): keys are matched after the 'mm_projector.'
    component, whatever prefix the saving run used."""
    weights = torch.load(path, map_location="cpu")
    sub = {k.split("mm_projector.")[1]: v for k, v in weights.items() if "mm_projector" in k}
    pj = dict(model.model.mm_projector.named_parameters())
    if set(sub) != set(pj):
        raise KeyError(f"mm_projector keys differ: {sorted(set(sub) ^ set(pj))}")
    with torch.no_grad():
        for k, v in sub.items():
            pj[k].copy_(v.to(pj[k].device, pj[k].dtype))


# ---- HF-layout full-model weights (what `trainer._save` → `save_pretrained` writes in the reference) ----
SAFE_WEIGHTS_NAME = "model.safetensors"
SAFE_WEIGHTS_INDEX_NAME = "model.safetensors.index.json"


def _parse_size(size) -> int:
    if isinstance(size, int):
        return size
    s = str(size).upper().strip()
    for unit, mul in (("GIB", 2 ** 30), ("MIB", 2 ** 20), ("KIB", 2 ** 10), ("GB", 10 ** 9), ("MB", 10 ** 6), ("KB", 10 ** 3)):
        if s.endswith(unit):
            return int(float(s[:-len(unit)]) * mul)
    return int(s)


def shard_state_dict(state_dict: Dict[str, torch.Tensor], max_shard_size="5GB"):
    """Greedy split in key order with huggingface_hub's rules (split_state_dict_into_shards_factory, which
    transformers' save_pretrained calls): a tensor larger than max_shard_size gets a shard of its own
    WITHOUT closing the shard being filled; any other tensor that would push the current shard over the
    limit closes it first.  → ({file: {name: tensor}}, index-or-None)."""
    limit = _parse_size(max_shard_size)
    shards, cur, cur_size, total = [], {}, 0, 0
    for k, v in state_dict.items():
        n = v.numel() * v.element_size()
        total += n
        if n > limit:
            shards.append({k: v})
            continue
        if cur_size + n > limit:
            shards.append(cur)
            cur, cur_size = {}, 0
        cur[k] = v
        cur_size += n
    if cur:
        shards.append(cur)
    if len(shards) <= 1:
        return {SAFE_WEIGHTS_NAME: shards[0] if shards else {}}, None
    files, weight_map = {}, {}
    for i, sh in enumerate(shards):
        name = f"model-{i + 1:05d}-of-{len(shards):05d}.safetensors"
        files[name] = sh
        for k in sh:
            weight_map[k] = name
    return files, {"metadata": {"total_size": total}, "weight_map": weight_map}


def save_pretrained_weights(state_dict: Dict[str, torch.Tensor], output_dir: str, max_shard_size="5GB"):
    """model.safetensors, or model-0000i-of-0000N.safetensors + model.safetensors.index.json — the
    layout `from_pretrained` (builder.py:58-138) reads."""
    from safetensors.torch import save_file

    os.makedirs(output_dir, exist_ok=True)
    files, index = shard_state_dict({k: v.detach().cpu().contiguous() for k, v in state_dict.items()}, max_shard_size)
    for name, sh in files.items():
        save_file(sh, os.path.join(output_dir, name), metadata={"format": "pt"})
    if index is not None:
        with open(os.path.join(output_dir, SAFE_WEIGHTS_INDEX_NAME), "w") as fh:
            fh.write(json.dumps(index, indent=2, sort_keys=True) + "\n")
    return sorted(files)


def load_pretrained_weights(model_dir: str) -> Dict[str, torch.Tensor]:
    """Reads either layout back (also a legacy pytorch_model.bin)."""
    from safetensors.torch import load_file

    idx = os.path.join(model_dir, SAFE_WEIGHTS_INDEX_NAME)
    if os.path.exists(idx):
        with open(idx) as fh:
            wm = json.load(fh)["weight_map"]
        out = {}
        for name in sorted(set(wm.values())):
            out.update(load_file(os.path.join(model_dir, name)))
        return {k: out[k] for k in wm}
    one = os.path.join(model_dir, SAFE_WEIGHTS_NAME)
    if os.path.exists(one):
        return load_file(one)
    return torch.load(os.path.join(model_dir, "pytorch_model.bin"), map_location="cpu")
