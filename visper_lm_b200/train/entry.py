"""`train()` — the caller of the training-step path (ola_vlm/train/ola_vlm_train.py:977-1326), so the
reference's launch scripts (scripts/train/pretrain.sh, finetune.sh) work with their own flags:

    torchrun --nproc-per-node 8 -m ola_vlm.train.ola_vlm_train --model_name_or_path /ckpt/Meta-Llama-3-8B-Instruct \
        --version llava_llama_3 --mode gen-depth-seg --layer_indices d18-20_s10-18_g12-20 ... (pretrain.sh:16-58)

Same order of operations as the reference: parse (ModelArguments, DataArguments, TrainingArguments) → load
the LLM into the Ola* class picked from the name → tokenizer / pad token → vision tower + projector →
tokenizer settings onto the config → adapter-tuning freeze → aux config from the --mode / --layer_indices /
--loss_weights DSLs and the head arguments → task tokens, heads, frozen teachers → requires_grad policy →
data module → LLaVATrainer.train(resume if checkpoint-* exists) → save_state → safe_save_model_for_hf_trainer.

What differs, all stated where it happens: DeepSpeed / wandb / LoRA / bitsandbytes flags are accepted and
reported as ignored (ZeRO-2 is the trainer's own; `--deepspeed .../zero3.json` is refused); the three
attributes the reference's train() reads but its ModelArguments never declares (task_token_format, use_ce,
sample_tokens — ola_vlm_train.py:1154,1157,1231) are declared here with the values the model code defaults
to; there is no network, so every path must be local.
"""
from __future__ import annotations

import os
import pathlib
import re
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch

from .data import DataArguments, make_supervised_data_module
from .trainer import LLaVATrainer, TrainingArguments


@dataclass
class ModelArguments:
    """ola_vlm_train.py:55-109, field for field."""
    model_name_or_path: Optional[str] = "facebook/opt-125m"
    version: Optional[str] = "v0"
    freeze_backbone: bool = False
    tune_mm_mlp_adapter: bool = False
    vision_tower: Optional[str] = None
    mm_vision_select_layer: Optional[int] = -1
    pretrain_mm_mlp_adapter: Optional[str] = None
    mm_projector_type: Optional[str] = "linear"
    mm_use_im_start_end: bool = False
    mm_use_im_patch_token: bool = True
    mm_patch_merge_type: Optional[str] = "flat"
    mm_vision_select_feature: Optional[str] = "patch"
    contrastive_loss_weight: Optional[float] = 0.3
    image_generator: Optional[str] = "stabilityai/stable-diffusion-2-1-unclip"
    image_segmentor: Optional[str] = "shi-labs/oneformer_coco_swin_large"
    depth_estimator: Optional[str] = "depth_anything_v2_vitl.pth"
    mode: Optional[str] = "gen-depth-seg"
    num_task_tokens: Optional[int] = 8
    pass_text_to_aux: Optional[bool] = True
    use_contrastive: Optional[bool] = True
    layer_indices: Optional[str] = "d8-20_s10-18_g12-20"
    loss_weights: Optional[str] = "d0.5_s0.5_g0.5"
    img_head_depth: Optional[int] = 1
    img_head_dim_head: Optional[int] = 32
    img_head_num_heads: Optional[int] = 4
    img_head_num_tokens: Optional[int] = 1
    img_head_output_dim: Optional[int] = 1024
    img_head_ff_mult: Optional[int] = 1
    seg_head_depth: Optional[int] = 1
    seg_head_dim_head: Optional[int] = 32
    seg_head_num_heads: Optional[int] = 4
    seg_head_num_tokens: Optional[int] = 576
    seg_head_output_dim: Optional[int] = 1536
    seg_head_ff_mult: Optional[int] = 1
    seg_teacher: Optional[str] = "oneformer"
    depth_head_depth: Optional[int] = 1
    depth_head_dim_head: Optional[int] = 32
    depth_head_num_heads: Optional[int] = 4
    depth_head_num_tokens: Optional[int] = 576
    depth_head_output_dim: Optional[int] = 1024
    depth_head_ff_mult: Optional[int] = 1
    freeze_task_token: Optional[bool] = False
    # read by the reference's train() but never declared there; the model code's defaults
    task_token_format: Optional[str] = "emb"
    use_ce: Optional[bool] = False
    sample_tokens: Optional[bool] = False
    # no network here: build the teachers' geometry with random weights when their files are absent
    random_init_teachers: bool = False


def parse_args(argv: Optional[List[str]] = None) -> Tuple[ModelArguments, DataArguments, TrainingArguments, List[str]]:
    """HfArgumentParser over the three dataclasses (ola_vlm_train.py:980-982).  Each class is parsed on its
    own because train() copies some flags across them (tune_mm_mlp_adapter, mm_use_im_start_end, version);
    returns also the flags none of them knows — HF / DeepSpeed / wandb switches of the launch scripts."""
    from transformers import HfArgumentParser

    import sys

    argv = list(sys.argv[1:] if argv is None else argv)
    parsed, unknown = [], None
    for dc in (ModelArguments, DataArguments, TrainingArguments):
        args, rest = HfArgumentParser(dc, allow_abbrev=False).parse_args_into_dataclasses(args=argv, return_remaining_strings=True)
        flags = {r for r in rest if r.startswith("--")}
        unknown = flags if unknown is None else unknown & flags
        parsed.append(args)
    ignored = [a for a in argv if a in unknown]
    for i, a in enumerate(argv):
        if a == "--deepspeed" and i + 1 < len(argv) and "zero3" in argv[i + 1]:
            raise NotImplementedError("ZeRO-3 (scripts/zero3.json) is not built: 180 GB parts hold the 8B configs "
                                      "under the trainer's ZeRO-2 (DESIGN.md §8)")
    return parsed[0], parsed[1], parsed[2], ignored


def parse_dsl(layer_indices: str, loss_weights: str):
    """--layer_indices 'd18-20_s10-18_g12-20' / --loss_weights 'd0.5_s0.5_g0.5' (ola_vlm_train.py:1159-1194)."""
    li = {"d": "0", "s": "0", "g": "0"}
    for m in re.findall(r"[a-zA-Z]\d+(?:-\d+)?", layer_indices):
        if m[0] in li:
            li[m[0]] = m[1:]
    lw = {"d": 0.5, "s": 0.5, "g": 0.5}
    for m in re.findall(r"[a-zA-Z]\d+\.\d+", loss_weights):
        if m[0] in lw:
            lw[m[0]] = float(m[1:])
    return li, lw


def configure_aux(model, model_args: ModelArguments):
    """ola_vlm_train.py:1147-1237: the aux-head configuration written onto model.config, then task tokens,
    heads and frozen teachers."""
    cfg = model.config
    cfg.aux_mode = model_args.mode
    cfg.contrastive_loss_weight = model_args.contrastive_loss_weight
    cfg.num_task_tokens = model_args.num_task_tokens
    cfg.task_token_format = model_args.task_token_format
    cfg.pass_text_to_aux = model_args.pass_text_to_aux
    cfg.use_contrastive = model_args.use_contrastive
    cfg.use_ce = model_args.use_ce
    li, lw = parse_dsl(model_args.layer_indices, model_args.loss_weights)

    def head(prefix, a, key):
        g = lambda f: getattr(model_args, f"{a}_head_{f}")  # noqa: E731
        return {"depth": g("depth"), "dim_head": g("dim_head"), "num_heads": g("num_heads"),
                "num_tokens": g("num_tokens"), "output_dim": g("output_dim"), "ff_mult": g("ff_mult"),
                f"{prefix}_layer_indices": li[key], f"{prefix}_loss_weight": lw[key]}

    cfg.image_gen = head("img", "img", "g")
    cfg.image_generator = model_args.image_generator
    cfg.image_seg = head("seg", "seg", "s")
    cfg.image_segmentor = model_args.image_segmentor
    cfg.image_depth = head("depth", "depth", "d")
    cfg.depth_estimator = model_args.depth_estimator
    cfg.sample_tokens = model_args.sample_tokens
    cfg.random_init_teachers = bool(model_args.random_init_teachers)
    model.img_gen_loss_weight, model.img_seg_loss_weight, model.img_depth_loss_weight = lw["g"], lw["s"], lw["d"]
    if model_args.num_task_tokens > 0:
        model.get_model().initialize_special_tokens(cfg)
    model.init_heads(cfg)
    fresh = ("model.special_", "image_gen_heads.", "image_depth_heads.", "image_seg_heads.", "_logit_scale")
    if getattr(model, "da_v2_head", None) is not None and not getattr(model, "_da_v2_head_loaded", False):
        if not model_args.random_init_teachers:   # the reference downloads it (base_ola_vlm.py:141-146); no network here
            raise FileNotFoundError(f"DPT depth decoder weights not found: {model_args.depth_estimator!r}")
        fresh += ("da_v2_head.",)
    model.init_weights(only=fresh)                  # task tokens, heads, logit scales: freshly created modules
    model.init_target_models(cfg)


def model_class(model_args: ModelArguments, distill: bool = True):
    """ola_vlm_train.py:1007-1021 ('phi' in the name picks the Phi-3 class); distill=False gives the NTP-only
    classes that ola_vlm/train/train.py:933-941 uses for the IFT / VPT stages."""
    from .. import model as pm

    phi = "phi" in model_args.model_name_or_path.lower()
    if distill:
        return pm.OlaLlavaPhi3ForCausalLM if phi else pm.OlaLlavaLlamaForCausalLM
    return pm.LlavaPhi3ForCausalLM if phi else pm.LlavaLlamaForCausalLM


def build_model(model_args: ModelArguments, data_args: DataArguments, training_args: TrainingArguments,
                tokenizer, device=None, distill: bool = True, config_overrides: Optional[dict] = None):
    """Everything train() does between argument parsing and the data module (ola_vlm_train.py:1007-1258).
    config_overrides: extra config keys (tests use a miniature `vision` geometry)."""
    from .checkpoint import load_mm_projector
    from .policy import apply_freeze_policy

    if model_args.vision_tower is None:
        raise ValueError("--vision_tower is required: this package builds the multimodal training path only")
    cls = model_class(model_args, distill)
    overrides = dict(mm_vision_tower=model_args.vision_tower, mm_vision_select_layer=model_args.mm_vision_select_layer,
                     mm_vision_select_feature=model_args.mm_vision_select_feature,
                     mm_projector_type=model_args.mm_projector_type)
    if model_args.mm_projector_type != "mlp2x_gelu":
        raise NotImplementedError(f"mm_projector_type {model_args.mm_projector_type!r}: every shipped script uses "
                                  "mlp2x_gelu, the only projector on the built path")
    overrides.update(config_overrides or {})
    model = cls.from_pretrained(model_args.model_name_or_path, device=device, **overrides)
    model.config.use_cache = False
    # tokenizer (:1077-1092)
    if tokenizer.pad_token is None:
        tokenizer.pad_token = tokenizer.unk_token
    if tokenizer.pad_token_id is None:
        n_new = tokenizer.add_special_tokens(dict(pad_token="<pad>"))
        model.resize_token_embeddings(len(tokenizer))
        if n_new > 0:   # smart_tokenizer_and_embedding_resize (:939-975): new rows ~ N(mean, std) of the table
            with torch.no_grad():
                n_tok = len(tokenizer)   # rows past len(tokenizer) are kernel padding (vocab rounded up to 8), not tokens
                for emb in (model.get_input_embeddings().weight, model.get_output_embeddings().weight):
                    old = emb[:n_tok - n_new].float()
                    emb[n_tok - n_new:n_tok] = torch.normal(old.mean().item(), old.std().item(),
                                                            size=emb[n_tok - n_new:n_tok].shape).to(emb.dtype)
    # vision tower + projector (:1099-1121): built with the model; a fresh projector needs initial values
    tower = model.get_vision_tower()
    missing = getattr(model, "_missing_from_checkpoint", [])
    if model_args.vision_tower and os.path.isdir(str(model_args.vision_tower)):
        tower.load_model(path=model_args.vision_tower)
    elif any(k.startswith("model.vision_tower.") for k in missing):
        # neither a local tower directory nor a checkpoint that carries the tower: the parameters are uninitialised
        # memory (the reference would download the hub id here; there is no network)
        if not model_args.random_init_teachers:
            raise FileNotFoundError(f"vision tower weights not found: --vision_tower {model_args.vision_tower!r} is not a "
                                    "local directory and the loaded checkpoint has no model.vision_tower.* tensors")
        model.init_weights(only=("model.vision_tower.",))
    # initialize_vision_modules builds a projector only if the model has none (llava_arch.py:126-127): a projector that
    # came with a multimodal checkpoint (finetune.sh / vpt.sh load the PT output) is kept
    if any(k.startswith("model.mm_projector.") for k in missing):
        model.init_weights(only=("model.mm_projector.",))
    data_args.image_processor = getattr(tower, "image_processor", None) or data_args.image_processor
    data_args.is_multimodal = True
    data_args.version = model_args.version
    cfg = model.config
    cfg.image_aspect_ratio = data_args.image_aspect_ratio
    cfg.tokenizer_padding_side = tokenizer.padding_side
    cfg.tokenizer_model_max_length = tokenizer.model_max_length
    cfg.tune_mm_mlp_adapter = training_args.tune_mm_mlp_adapter = model_args.tune_mm_mlp_adapter
    cfg.freeze_mm_mlp_adapter = getattr(training_args, "freeze_mm_mlp_adapter", False)
    cfg.mm_use_im_start_end = data_args.mm_use_im_start_end = model_args.mm_use_im_start_end
    cfg.mm_projector_lr = training_args.mm_projector_lr
    training_args.use_im_start_end = model_args.mm_use_im_start_end
    cfg.mm_use_im_patch_token = model_args.mm_use_im_patch_token
    if model_args.pretrain_mm_mlp_adapter:
        load_mm_projector(model, model_args.pretrain_mm_mlp_adapter)           # ola_arch.py:139-144
    model.initialize_vision_tokenizer(model_args, tokenizer=tokenizer)
    if distill and "ola" not in model_args.model_name_or_path.split("/")[-1]:     # :1147
        configure_aux(model, model_args)
    trainable = apply_freeze_policy(model, tune_mm_mlp_adapter=model_args.tune_mm_mlp_adapter,
                                    freeze_mm_mlp_adapter=cfg.freeze_mm_mlp_adapter,
                                    freeze_task_token=bool(model_args.freeze_task_token),
                                    freeze_backbone=model_args.freeze_backbone)
    return model, trainable


def train(argv: Optional[List[str]] = None, attn_implementation=None, distill: bool = True):
    """ola_vlm_train.py:977.  `attn_implementation` is accepted for the call made by ola_vlm_train_mem.py:5
    (flash attention is this package's only attention)."""
    import torch.distributed as dist
    import transformers

    model_args, data_args, training_args, ignored = parse_args(argv)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=device)
    rank = dist.get_rank() if dist.is_initialized() else 0
    if rank == 0 and ignored:
        print(f"[visper_lm_b200] flags accepted and ignored (DeepSpeed / HF Trainer / wandb switches): {' '.join(ignored)}")
    tokenizer = transformers.AutoTokenizer.from_pretrained(model_args.model_name_or_path,
                                                           model_max_length=training_args.model_max_length,
                                                           padding_side="right", use_fast=False)
    model, trainable = build_model(model_args, data_args, training_args, tokenizer, device=device, distill=distill)
    if rank == 0:
        n = sum(p.numel() for p in model.parameters() if p.requires_grad)
        print(f"[visper_lm_b200] {len(trainable)} trainable tensors, {n / 1e6:.1f} M parameters")
    data_module = make_supervised_data_module(tokenizer=tokenizer, data_args=data_args)
    trainer = LLaVATrainer(model=model, tokenizer=tokenizer, args=training_args, **data_module)
    if list(pathlib.Path(training_args.output_dir).glob("checkpoint-*")):
        trainer.train(resume_from_checkpoint=True)
    else:
        trainer.train()
    trainer.save_state()
    model.config.use_cache = True
    from .checkpoint import safe_save_model_for_hf_trainer

    safe_save_model_for_hf_trainer(trainer=trainer, output_dir=training_args.output_dir)
    return trainer
