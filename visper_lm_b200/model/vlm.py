"""B200-native stand-ins for the reference's model classes, behind the same Python surface:

  OlaLlavaLlamaForCausalLM / OlaLlavaPhi3ForCausalLM   (ola_vlm/model/language_model/ola_llama.py:58,
                                                         ola_phi3.py:58 — NTP + distillation heads)
  LlavaLlamaForCausalLM / LlavaPhi3ForCausalLM         (llava_llama.py:51, llava_phi3.py — NTP only)

forward(...) takes the collator's batch and returns an object with .loss, .logits,
.hidden_states, .image_embs / .seg_embs / .depth_embs / .depth_preds like OlaCausalLLMOutputWithPast
(ola_llama.py:36-44).  prepare_inputs_labels_for_multimodal keeps the reference signature
(ola_arch.py:256-259) but is sync-free on the device: the splice is planned on the host from the
token ids and executed as one gather kernel.

Teachers: the depth teacher (Depth-Anything-V2's DINOv2-L, model/dinov2.py) and the frozen DPT decoder
behind `depth_preds` (model/dpt.py), the unCLIP image encoder (model/gen_teacher.py) and OneFormer's
Swin-L backbone (model/seg_teacher.py) run on the GPU, batched, behind the reference's hooks
_get_dav2_feats / _get_gen_feats / _get_seg_targets (base_ola_vlm.py:323,347,382); targets can also be
supplied by the caller (`distill_targets=`).  Out of scope: generation, wandb logging.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from types import SimpleNamespace
from typing import List, Optional

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

from .. import autograd as A
from .. import ops
from ..ops import ACT_GELU, ACT_NONE, BF16
from . import modules as M

IGNORE_INDEX = -100
IMAGE_TOKEN_INDEX = -200


# ------------------------------------------------------------------------------------------------ config
class VisperConfig:
    """Attribute bag with the HF-style names the reference reads (hidden_size, vocab_size, …) plus
    the aux-head injection of ola_vlm/train/ola_vlm_train.py:1149-1229."""

    model_type = "visper"

    def __init__(self, family="llama", vocab_size=128256, hidden_size=4096, intermediate_size=14336,
                 num_hidden_layers=32, num_attention_heads=32, num_key_value_heads=8,
                 max_position_embeddings=4096, rope_theta=500000.0, rms_norm_eps=1e-5,
                 vision=None, mm_projector_type="mlp2x_gelu", mm_vision_select_layer=-2,
                 mm_vision_select_feature="patch", tokenizer_model_max_length=4096,
                 tokenizer_padding_side="right", **extra):
        self.family = family
        self.vocab_size = vocab_size
        self.hidden_size = hidden_size
        self.intermediate_size = intermediate_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.num_key_value_heads = num_key_value_heads
        self.max_position_embeddings = max_position_embeddings
        self.rope_theta = rope_theta
        self.rms_norm_eps = rms_norm_eps
        self.vision = vision or dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24,
                                     num_attention_heads=16, image_size=336, patch_size=14)
        # multimodal_encoder/builder.py:6-13 picks the tower class from this name: "convnext" in it selects
        # CLIPConvNextVisionTower (e.g. "CLIP-convnext_xxlarge-res768", BASELINE config 4), else the CLIP ViT
        self.mm_vision_tower = extra.pop("mm_vision_tower", None) or "openai/clip-vit-large-patch14-336"
        if "convnext" in self.mm_vision_tower.lower():
            from .convnext import CONVNEXT_PRESETS, extract_res_interp

            base, res, _ = extract_res_interp(self.mm_vision_tower)
            if vision is None or "dims" not in vision:
                self.vision = dict(CONVNEXT_PRESETS[base], image_size=res or 768)
            self.mm_hidden_size = self.vision["dims"][-1]
        else:
            self.mm_hidden_size = self.vision["hidden_size"]
        self.mm_projector_type = mm_projector_type
        self.mm_vision_select_layer = mm_vision_select_layer
        self.mm_vision_select_feature = mm_vision_select_feature
        self.tokenizer_model_max_length = tokenizer_model_max_length
        self.tokenizer_padding_side = tokenizer_padding_side
        self.use_return_dict = True
        self.output_hidden_states = False
        # Phi-3-mini-4k ships sliding_window=2047 (HF Phi3Config); HF 4.41.1 hands it to flash_attn as
        # window_size=(2047, 2047): key j visible to query i iff 0 <= i - j <= 2047.  None = full causal.
        self.sliding_window = 2047 if family == "phi3" else None
        self.zero_masks_like_reference = False  # off for training value
        self.materialize_logits = False         # .logits only on request while training
        self.num_task_tokens = 0
        for k, v in extra.items():
            setattr(self, k, v)

    def inject_aux(self, mode="gen-depth-seg", layer_indices="d18-20_s10-18_g12-20",
                   loss_weights="d0.5_s0.5_g0.5", num_task_tokens=8, contrastive_loss_weight=0.3,
                   gen_dim=1024, seg_dim=1536, depth_dim=1024, pass_text_to_aux=True,
                   use_contrastive=True):
        """The string DSLs of --layer_indices / --loss_weights (ola_vlm_train.py:1159-1194)."""
        import re

        li = {"d": "0", "s": "0", "g": "0"}
        for m in re.findall(r"[a-zA-Z]\d+(?:-\d+)?", layer_indices):
            li[m[0]] = m[1:]
        lw = {"d": 0.5, "s": 0.5, "g": 0.5}
        for m in re.findall(r"[a-zA-Z]\d+\.\d+", loss_weights):
            lw[m[0]] = float(m[1:])

        def head(prefix, nt, od, key):
            return {"depth": 1, "dim_head": 32, "num_heads": 4, "num_tokens": nt, "output_dim": od,
                    "ff_mult": 1, f"{prefix}_layer_indices": li[key], f"{prefix}_loss_weight": lw[key]}

        self.aux_mode = mode
        self.contrastive_loss_weight = contrastive_loss_weight
        self.num_task_tokens = num_task_tokens
        self.task_token_format = "emb"
        self.pass_text_to_aux = pass_text_to_aux
        self.use_contrastive = use_contrastive
        self.use_ce = False
        self.sample_tokens = False
        self.image_gen = head("img", 1, gen_dim, "g")
        self.image_seg = head("seg", 576, seg_dim, "s")
        self.image_seg["seg_teacher"] = "oneformer"
        self.image_depth = head("depth", 576, depth_dim, "d")
        return self


class OlaLlavaLlamaConfig(VisperConfig):
    model_type = "ola_llama"


class OlaLlavaPhi3Config(VisperConfig):
    model_type = "ola_phi3"


class LlavaConfig(VisperConfig):
    model_type = "llava_llama"


class LlavaPhi3Config(VisperConfig):
    model_type = "llava_phi3"


@dataclass
class OlaCausalLLMOutputWithPast:
    loss: Optional[torch.Tensor] = None
    logits: Optional[torch.Tensor] = None
    past_key_values: Optional[tuple] = None
    hidden_states: Optional[tuple] = None
    attentions: Optional[tuple] = None
    image_embs: Optional[list] = None
    seg_embs: Optional[list] = None
    depth_embs: Optional[list] = None
    depth_preds: Optional[list] = None
    # extras (not in the reference's dataclass): the loss components as tensors instead of the
    # wandb logging done inside the reference's forward (ola_llama.py:146-168)
    text_loss: Optional[torch.Tensor] = None
    loss_terms: Optional[dict] = None

    def __getitem__(self, k):
        if isinstance(k, str):
            return getattr(self, k)
        return tuple(v for v in (self.loss, self.logits, self.past_key_values, self.hidden_states)
                     if v is not None)[k]


# ------------------------------------------------------------------------------------------------ splice plan
class SplicePlan:
    """Host-built index plan for prepare_inputs_labels_for_multimodal (ola_arch.py:337-444)."""

    def __init__(self, input_ids, labels, attention_mask, n_img_tok, task_rows, max_len, pad_side):
        ids = input_ids.numpy() if isinstance(input_ids, torch.Tensor) else np.asarray(input_ids)
        B = ids.shape[0]
        am = np.ones_like(ids, dtype=bool) if attention_mask is None else attention_mask.numpy().astype(bool)
        lab = np.full_like(ids, IGNORE_INDEX) if labels is None else labels.numpy()
        kinds, idxs, labs = [], [], []
        img_slot = 0
        for b in range(B):
            cid, clab = ids[b][am[b]], lab[b][am[b]]
            pos = np.nonzero(cid == IMAGE_TOKEN_INDEX)[0]
            if len(pos) == 0:
                kinds.append(np.zeros(len(cid), np.int32))
                idxs.append(cid.astype(np.int32))
                labs.append(clab)
                img_slot += 1  # ola_arch.py:354 consumes an image slot without using it
                continue
            k_parts, i_parts, l_parts = [], [], []
            bounds = [-1] + pos.tolist() + [len(cid)]
            for i in range(len(bounds) - 1):
                seg = slice(bounds[i] + 1, bounds[i + 1])
                n = bounds[i + 1] - bounds[i] - 1
                k_parts.append(np.zeros(n, np.int32))
                i_parts.append(cid[seg].astype(np.int32))
                l_parts.append(clab[seg])
                if i < len(pos):
                    k_parts.append(np.full(n_img_tok, 1, np.int32))
                    i_parts.append(np.arange(img_slot * n_img_tok, (img_slot + 1) * n_img_tok, dtype=np.int32))
                    l_parts.append(np.full(n_img_tok, IGNORE_INDEX, clab.dtype))
                    img_slot += 1
                    if task_rows:
                        k_parts.append(np.full(task_rows, 2, np.int32))
                        i_parts.append(np.arange(task_rows, dtype=np.int32))
                        l_parts.append(np.full(task_rows, IGNORE_INDEX, clab.dtype))
            kinds.append(np.concatenate(k_parts))
            idxs.append(np.concatenate(i_parts))
            labs.append(np.concatenate(l_parts))
        if max_len is not None:
            kinds = [k[:max_len] for k in kinds]
            idxs = [i[:max_len] for i in idxs]
            labs = [l[:max_len] for l in labs]
        T = max(len(k) for k in kinds)
        if pad_side != "right":
            raise NotImplementedError("left padding is not on the training path (tokenizer padding_side='right')")
        self.B, self.T = B, T
        self.n_images = img_slot
        kind = np.full((B, T), -1, np.int32)
        index = np.full((B, T), -1, np.int32)
        out_lab = np.full((B, T), IGNORE_INDEX, np.int64)
        mask = np.zeros((B, T), bool)
        for b in range(B):
            n = len(kinds[b])
            kind[b, :n], index[b, :n], out_lab[b, :n], mask[b, :n] = kinds[b], idxs[b], labs[b], True
        flat_kind, flat_index = kind.reshape(-1), index.reshape(-1)
        rows = np.arange(B * T, dtype=np.int32)
        inv_img = np.full(img_slot * n_img_tok, -1, np.int32)
        sel = flat_kind == 1
        inv_img[flat_index[sel]] = rows[sel]
        inv_task = None
        if task_rows:
            inv_task = np.full((task_rows, B), -1, np.int32)
            sel = np.nonzero(flat_kind == 2)[0]
            # at most one image (→ one task block) per sample contributes per column; extra images
            # of a multi-image sample add further columns
            cols = {}
            for r in sel:
                s = flat_index[r]
                c = cols.get(s, 0)
                if c >= inv_task.shape[1]:
                    inv_task = np.concatenate([inv_task, np.full((task_rows, 1), -1, np.int32)], 1)
                inv_task[s, c] = r
                cols[s] = c + 1
        embed_scatter = np.where(flat_kind == 0, flat_index, -1).astype(np.int32)
        self.np = dict(kind=flat_kind, index=flat_index, inv_img=inv_img, inv_task=inv_task,
                       embed_scatter=embed_scatter, labels=out_lab, mask=mask)
        self.all_valid = bool(mask.all())
        # Rows that carry a next-token target (shifted label != -100): lm_head + CE run on these
        # only — ignored rows contribute neither loss nor gradient (ola_llama.py:126-136).
        shifted = np.full((B, T), IGNORE_INDEX, np.int64)
        shifted[:, :-1] = out_lab[:, 1:]
        shifted = shifted.reshape(-1)
        valid = np.nonzero(shifted != IGNORE_INDEX)[0].astype(np.int32)
        inv = np.full(B * T, -1, np.int32)
        inv[valid] = np.arange(len(valid), dtype=np.int32)
        self.np.update(ce_rows=valid, ce_targets=shifted[valid], ce_inv=inv)
        self.n_ce_rows = int(len(valid))

    def ce_compaction(self):
        """(rows, targets, inverse, n) device index tensors for the label-carrying rows, or None when
        compaction would not pay (nearly every row is scored, or none is)."""
        if self.n_ce_rows == 0 or self.n_ce_rows > 0.9 * self.B * self.T:
            return None
        return self.ce_rows, self.ce_targets, self.ce_inv, self.n_ce_rows

    def to(self, device):
        n = self.np
        t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(device, non_blocking=True)
        self.kind, self.index = t(n["kind"]), t(n["index"])
        self.inv_img, self.embed_scatter = t(n["inv_img"]), t(n["embed_scatter"])
        self.inv_task = t(n["inv_task"])
        if n["inv_task"] is not None:
            self.B_cols = n["inv_task"].shape[1]
        self.labels = t(n["labels"])
        self.mask = t(n["mask"])
        self.ce_rows, self.ce_targets, self.ce_inv = t(n["ce_rows"]), t(n["ce_targets"]), t(n["ce_inv"])
        return self


class HeadPlan:
    """Token selection of forward_emb_predictor (base_ola_vlm.py:413-443) as gather indices."""

    _cache = {}

    @classmethod
    def get(cls, B, T, S, nt, order, task, pass_text, n_latents, heads, device):
        key = (B, T, S, nt, tuple(order), task, pass_text, n_latents, heads, str(device))
        if key not in cls._cache:
            cls._cache[key] = cls(B, T, S, nt, order, task, pass_text, n_latents, heads, device)
        return cls._cache[key]

    def __init__(self, B, T, S, nt, order, task, pass_text, n_latents, heads, device):
        k = order.index(task)
        t0 = S + 576 + nt * k
        end = S + 576 + nt * len(order)
        if nt == 0 or T < 600:
            cols = np.arange(T) if pass_text else np.arange(min(S + 576, T))  # the slice clamps at T
        else:
            parts = [np.arange(S + 576), np.arange(t0, t0 + nt)]
            if pass_text:
                parts.append(np.arange(end, T))
            cols = np.concatenate(parts)
        nk = len(cols)
        ctx = (np.arange(B)[:, None] * T + cols[None, :]).astype(np.int32)  # [B, nk]
        inv_ctx = np.full(B * T, -1, np.int32)
        inv_ctx[ctx.reshape(-1)] = np.arange(B * nk, dtype=np.int32)
        self.B, self.nk, self.nt, self.heads = B, nk, nt, heads
        dev = device
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a.astype(np.int32))).to(dev)
        self.ctx_index, self.inv_ctx = t(ctx.reshape(-1)), t(inv_ctx)
        if task == "gen":
            # latents = the 8 hidden states of this task's tokens INSIDE inp_tokens
            # (base_ola_vlm.py:439: inp_tokens[:, S+576 : S+576+nt]), mean-pooled to 1 query
            # (:437 without text in the head input: the LAST nt tokens of inp_tokens)
            lat_cols = cols[S + 576:S + 576 + nt] if pass_text else cols[-nt:]
            gi = (np.arange(B)[:, None] * T + lat_cols[None, :])
            self.gen_index = t(gi.reshape(-1))
            self.nq = 1
            self.lat_index = self.inv_lat = None
        else:
            self.nq = n_latents
            self.lat_index = t(np.tile(np.arange(n_latents), B))
            # inv_lat[r, b] = row of (b, r) inside the latent block
            self.inv_lat = t((np.arange(n_latents)[:, None] + np.arange(B)[None, :] * n_latents).reshape(-1))
            self.gen_index = None


def tile_targets(tgt, mask, B):
    """_emb_loss (base_ola_vlm.py:292-299): teacher targets (and their mask) with fewer rows than the
    predictions — fewer images than samples — are tiled along the batch and cut to B rows."""
    if tgt is not None and tgt.shape[0] != B:
        rf = max(1, B // tgt.shape[0])
        tgt = tgt.repeat(rf, *([1] * (tgt.dim() - 1)))[:B]
        if mask is not None and mask.shape[0] != B:
            mask = mask.reshape(-1).repeat(rf)[:B]
    return tgt, mask


_TARGET_STREAMS = {}


def gather_targets(tgt_flat, group=None):
    """dist_collect (ola_utils.py:96-106) for the InfoNCE negatives: all-gather of the rank-local
    targets [B, n] → ([world·B, n], offset of the local rows = rank·B (ola_utils.py:111), event or None).
    Targets carry no gradient, so a plain collective suffices (no backward collective).  On the GPU the
    collective runs on a side stream — the forward pass gets no extra point at which the ranks wait for each
    other — and the consumer waits for the returned event right before the loss kernel reads the targets."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        world = dist.get_world_size(group)
        src = tgt_flat.contiguous()
        out = torch.empty((world * src.shape[0], src.shape[1]), dtype=src.dtype, device=src.device)
        off = dist.get_rank(group) * src.shape[0]
        if src.is_cuda:
            st = _TARGET_STREAMS.get(src.device)
            if st is None:
                st = _TARGET_STREAMS[src.device] = torch.cuda.Stream(device=src.device)
            ready = torch.cuda.Event()
            ready.record()
            st.wait_event(ready)
            with torch.cuda.stream(st):
                dist.all_gather_into_tensor(out, src, group=group)
                done = torch.cuda.Event()
                done.record()
            src.record_stream(st)
            out.record_stream(st)
            return out, off, done
        dist.all_gather_into_tensor(out, src, group=group)
        return out, off, None
    return tgt_flat, 0, None


# ------------------------------------------------------------------------------------------------ model
class VisperModel(nn.Module):
    """OlaLlavaLlamaModel / OlaLlavaPhi3Model (ola_llama.py:51-55 + OlaLlavaMetaModel ola_arch.py:35-94)."""

    def __init__(self, config, device=None):
        super().__init__()
        self.config = config
        D = config.hidden_size
        self.embed_tokens = M.Weight((config.vocab_size, D), None, device)
        self.layers = nn.ModuleList([M.DecoderLayer(config, device) for _ in range(config.num_hidden_layers)])
        self.norm = M.Norm(D, False, device)
        if "convnext" in config.mm_vision_tower.lower():
            from .convnext import CLIPConvNextVisionTower

            self.vision_tower = CLIPConvNextVisionTower(config.mm_vision_tower, args=config, device=device,
                                                        cfg=config.vision)
        else:
            self.vision_tower = M.CLIPVisionTower(config.vision, config.mm_vision_select_layer,
                                                  config.mm_vision_select_feature, device)
        self.mm_projector = M.Seq(_0=M.Linear(config.mm_hidden_size, D, True, device),
                                  _2=M.Linear(D, D, True, device))
        self.aux_tokens = "depth-seg-gen"
        self.token_order = ["depth", "seg", "gen"]
        self.num_task_tokens = 0
        self._device = device
        if getattr(config, "num_task_tokens", 0) and hasattr(config, "aux_mode"):
            self.initialize_special_tokens(config)

    def get_vision_tower(self):
        return self.vision_tower

    def get_special_tokens(self):
        return (getattr(self, "special_depth_tokens", None), getattr(self, "special_seg_tokens", None),
                getattr(self, "special_gen_tokens", None))

    def initialize_special_tokens(self, config):
        """ola_arch.py:67-94"""
        self.num_task_tokens = config.num_task_tokens
        self.task_token_format = getattr(config, "task_token_format", "emb")
        self.aux_tokens = config.aux_mode
        self.token_order = config.aux_mode.split("-")
        D = config.hidden_size
        if self.num_task_tokens > 0:
            if "depth" in config.aux_mode:
                assert config.image_depth["num_tokens"] % self.num_task_tokens == 0
                self.special_depth_tokens = M._param(config.image_depth["num_tokens"], D, device=self._device)
            if "seg" in config.aux_mode:
                assert config.image_seg["num_tokens"] % self.num_task_tokens == 0
                self.special_seg_tokens = M._param(config.image_seg["num_tokens"], D, device=self._device)
            if "gen" in config.aux_mode:
                self.special_gen_tokens = M._param(self.num_task_tokens, D, device=self._device)

    def initialize_vision_modules(self, model_args=None, fsdp=None):
        return  # tower + projector are built in __init__ (config.mm_vision_tower is always set here)


class VisperForCausalLM(nn.Module):
    """Shared implementation; the four public classes below only pick family / distillation."""

    family = "llama"
    distill = True
    config_class = VisperConfig
    supports_param_sync = True   # forward calls self._pre_trainable_hook() before its first trainable weight
    _pre_trainable_hook = None

    def __init__(self, config, device=None):
        super().__init__()
        config.family = self.family
        self.config = config
        self.steps = 0
        self.model = VisperModel(config, device)
        self.vocab_size = config.vocab_size
        if self.family == "phi3":
            self.NUM_SYS_TOKENS = 13  # ola_phi3.py:68
        else:
            self.NUM_SYS_TOKENS = 26 if self.vocab_size < 128000 else 38  # ola_llama.py:65-68
        self.lm_head = M.Linear(config.hidden_size, config.vocab_size, False, device)
        self._device = device
        if self.distill and hasattr(config, "image_gen"):
            self.init_heads(config)

    # parameters a plain HF Llama-3 / Phi-3 checkpoint does not have: they keep their fresh init, as with
    # the reference's `from_pretrained(base_llm)` (ola_vlm_train.py:1007-1021)
    NEW_MODULE_KEYS = ("model.mm_projector.", "model.vision_tower.", "model.special_", "image_gen_heads.",
                       "image_depth_heads.", "image_seg_heads.", "_logit_scale", "da_v2_head.")

    @classmethod
    def from_pretrained(cls, model_dir, device=None, **config_overrides):
        """Load what `trainer._save` / `save_pretrained` wrote (config.json + model.safetensors or its
        5 GB shards + index; builder.py:58-138 loads the reference's models the same way) — or a plain
        HF Llama-3 / Phi-3 checkpoint, whose config keys and tensor names are the same; the multimodal
        modules it lacks stay freshly initialised.  Teacher submodules are not part of the architecture:
        call init_target_models afterwards if needed."""
        import json
        import os

        from ..train.checkpoint import load_pretrained_weights

        with open(os.path.join(model_dir, "config.json")) as fh:
            d = json.load(fh)
        d.pop("model_type", None)
        d.pop("family", None)
        d.pop("architectures", None)
        d.update(config_overrides)
        model = cls(cls.config_class(**d), device=device)
        sd = load_pretrained_weights(model_dir)
        own = model.state_dict()
        missing = [k for k in own if k not in sd and not any(t in k for t in cls.NEW_MODULE_KEYS)]
        # teachers are not part of the architecture; the NTP-only classes drop a distilled checkpoint's heads the way
        # HF from_pretrained drops unexpected keys (vpt.sh / finetune.sh load the PT output into LlavaLlama*)
        skip = ("dav2_backbone.", "oneformer.")
        if not cls.distill:
            skip += ("image_gen_heads.", "image_depth_heads.", "image_seg_heads.", "da_v2_head.", "gen_logit_scale",
                     "depth_logit_scale", "seg_logit_scale")
        unexpected = [k for k in sd if k not in own and not k.startswith(skip)]
        if missing or unexpected:
            raise KeyError(f"checkpoint does not match {cls.__name__}: missing {missing[:5]}, unexpected {unexpected[:5]}")
        model.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=False)
        # what the checkpoint did NOT carry (fresh modules the caller must initialise or load): a plain LLM
        # checkpoint lacks tower / projector / heads, a multimodal one (PT / VPT output) brings them
        model._missing_from_checkpoint = [k for k in own if k not in sd]
        return model

    # ---- reference accessors ---------------------------------------------------------------
    def get_model(self):
        return self.model

    def get_vision_tower(self):
        return self.model.get_vision_tower()

    # ---- tokenizer / embedding surface (ola_arch.py:446-488; HF PreTrainedModel names) ------------
    def get_input_embeddings(self):
        return self.model.embed_tokens

    def get_output_embeddings(self):
        return self.lm_head

    @torch.no_grad()
    def resize_token_embeddings(self, new_num_tokens, pad_to_multiple_of=None):
        """HF PreTrainedModel.resize_token_embeddings for embed_tokens + lm_head (setup time, not step
        work): old rows kept, new rows ~ N(0, 0.02).  The GEMM / CE kernels need a row count that is a
        multiple of 8, so the size is rounded up to one (what HF's pad_to_multiple_of=8 does: the padding
        rows are ordinary vocabulary entries that never occur as labels)."""
        mult = max(8, int(pad_to_multiple_of or 1))
        mult = mult if mult % 8 == 0 else mult * 8
        n = (int(new_num_tokens) + mult - 1) // mult * mult
        for holder in (self.model.embed_tokens, self.lm_head):
            old = holder.weight
            if old.shape[0] == n:
                continue
            w = torch.empty((n, old.shape[1]), dtype=old.dtype, device=old.device)
            w.normal_(0.0, 0.02)
            k = min(n, old.shape[0])
            w[:k] = old[:k]
            holder.weight = nn.Parameter(w, requires_grad=old.requires_grad)
        self.config.vocab_size = self.vocab_size = n
        self._lm_head_t = M.FrozenTranspose()
        return self.model.embed_tokens

    def initialize_vision_tokenizer(self, model_args, tokenizer):
        """ola_arch.py:446-488, statement for statement: optional <im_patch> / <im_start>,<im_end>
        tokens, new rows set to the mean of the old ones, the adapter-tuning freeze policy and the
        pretrain_mm_mlp_adapter embedding hand-over.  (Both flags are False in every shipped script.)"""
        from ola_vlm.constants import DEFAULT_IM_END_TOKEN, DEFAULT_IM_START_TOKEN, DEFAULT_IMAGE_PATCH_TOKEN

        if model_args.mm_use_im_patch_token:
            tokenizer.add_tokens([DEFAULT_IMAGE_PATCH_TOKEN], special_tokens=True)
            self.resize_token_embeddings(len(tokenizer))
        if model_args.mm_use_im_start_end:
            num_new_tokens = tokenizer.add_tokens([DEFAULT_IM_START_TOKEN, DEFAULT_IM_END_TOKEN], special_tokens=True)
            self.resize_token_embeddings(len(tokenizer))
            n_tok = len(tokenizer)        # rows beyond len(tokenizer) are kernel padding, not new tokens
            if num_new_tokens > 0:
                input_embeddings = self.get_input_embeddings().weight.data
                output_embeddings = self.get_output_embeddings().weight.data
                lo = n_tok - num_new_tokens
                input_embeddings[lo:n_tok] = input_embeddings[:lo].float().mean(dim=0, keepdim=True).to(input_embeddings.dtype)
                output_embeddings[lo:n_tok] = output_embeddings[:lo].float().mean(dim=0, keepdim=True).to(output_embeddings.dtype)
            if model_args.tune_mm_mlp_adapter:
                self.get_input_embeddings().weight.requires_grad = True
                self.get_output_embeddings().weight.requires_grad = False
            if model_args.pretrain_mm_mlp_adapter:
                mm_projector_weights = torch.load(model_args.pretrain_mm_mlp_adapter, map_location="cpu")
                embed_tokens_weight = mm_projector_weights["model.embed_tokens.weight"]
                assert num_new_tokens == 2
                input_embeddings = self.get_input_embeddings().weight.data
                lo = n_tok - num_new_tokens
                if embed_tokens_weight.shape[0] == num_new_tokens:
                    input_embeddings[lo:n_tok] = embed_tokens_weight.to(input_embeddings)
                elif embed_tokens_weight.shape[1] == input_embeddings.shape[1] and embed_tokens_weight.shape[0] >= n_tok:
                    input_embeddings[lo:n_tok] = embed_tokens_weight[lo:n_tok].to(input_embeddings)
                else:
                    raise ValueError(f"Unexpected embed_tokens_weight shape. Pretrained: {embed_tokens_weight.shape}. "
                                     f"Current: {input_embeddings.shape}. Numer of new tokens: {num_new_tokens}.")
        elif model_args.mm_use_im_patch_token:
            if model_args.tune_mm_mlp_adapter:
                self.get_input_embeddings().weight.requires_grad = False
                self.get_output_embeddings().weight.requires_grad = False

    @property
    def depth_tokens(self):
        return self.model.get_special_tokens()[0]

    @property
    def seg_tokens(self):
        return self.model.get_special_tokens()[1]

    @property
    def gen_tokens(self):
        return self.model.get_special_tokens()[2]

    @property
    def num_task_tokens(self):
        return self.model.num_task_tokens

    @property
    def token_order(self):
        return self.model.token_order

    @property
    def device(self):
        return self.lm_head.weight.device

    def init_target_models(self, config):
        """base_ola_vlm.py:61-95: builds the frozen teachers, all of which run batched on the GPU —
        `dav2_backbone` (Depth-Anything-V2 DINOv2-L, model/dinov2.py), `pipe.image_encoder` (unCLIP
        ViT-H/14, model/gen_teacher.py) and `oneformer` (Swin-L backbone, model/seg_teacher.py).
        Weights come from the reference's paths (`config.depth_estimator`, `config.image_generator`,
        `config.image_segmentor`); there is no network here, so a missing file raises unless
        `config.random_init_teachers` is set (benchmarks / tests)."""
        import os

        mode = getattr(config, "aux_mode", "gen-depth-seg")
        if hasattr(config, "image_gen") and "gen" in mode:
            self._init_gen_teacher(config)
        if hasattr(config, "image_seg") and "seg" in mode:
            self._init_seg_teacher(config)
        if not (hasattr(config, "image_depth") and "depth" in mode):
            return
        from .dinov2 import DepthAnythingV2
        self.dav2_backbone = DepthAnythingV2(encoder="vitl", features=256, out_channels=(256, 512, 1024, 1024),
                                             device=self._device)
        path = getattr(config, "depth_estimator", None)
        if path and os.path.exists(path):
            self.dav2_backbone.load_state_dict(torch.load(path, map_location="cpu"))
        elif getattr(config, "random_init_teachers", False):
            with torch.no_grad():
                for p_ in self.dav2_backbone.parameters():
                    p_.normal_(0.0, 0.02)
                for blk in self.dav2_backbone.pretrained.blocks:
                    blk.norm1.weight.fill_(1.0), blk.norm2.weight.fill_(1.0)
                self.dav2_backbone.pretrained.norm.weight.fill_(1.0)
        else:
            raise FileNotFoundError(f"depth teacher weights not found: {path!r} (the reference downloads "
                                    "depth_anything_v2_vitl.pth; no network here)")
        self.dav2_backbone.requires_grad_(False)

    def _init_gen_teacher(self, config):
        """base_ola_vlm.py:61-67: `self.pipe` = the unCLIP pipeline, of which only feature_extractor
        and image_encoder are used in training (:323-333).  Here `pipe` holds exactly those two: HF's
        CLIPImageProcessor (host-side PIL preprocessing) and the on-GPU encoder (model/gen_teacher.py).
        Weights: <config.image_generator>/image_encoder/{model.safetensors | pytorch_model.bin}."""
        import os
        from .gen_teacher import UNCLIP_VIT_H, CLIPVisionModelWithProjection

        enc = CLIPVisionModelWithProjection(UNCLIP_VIT_H, self._device)
        root = os.path.join(str(getattr(config, "image_generator", "")), "image_encoder")
        st, pt = os.path.join(root, "model.safetensors"), os.path.join(root, "pytorch_model.bin")
        if os.path.exists(st):
            from safetensors.torch import load_file
            enc.load_state_dict(load_file(st), strict=False)
        elif os.path.exists(pt):
            enc.load_state_dict(torch.load(pt, map_location="cpu"), strict=False)
        elif getattr(config, "random_init_teachers", False):
            with torch.no_grad():
                for n_, p_ in enc.named_parameters():
                    p_.fill_(1.0) if ("norm" in n_ and n_.endswith("weight")) else p_.normal_(0.0, 0.02)
        else:
            raise FileNotFoundError(f"gen teacher weights not found under {root!r} (the reference downloads "
                                    "stabilityai/stable-diffusion-2-1-unclip; no network here)")
        try:
            from transformers import CLIPImageProcessor
            fe = CLIPImageProcessor()  # the unCLIP feature_extractor's settings are the class defaults
        except Exception:  # pragma: no cover
            fe = None
        self.pipe = SimpleNamespace(image_encoder=enc, feature_extractor=fe)

    def _init_seg_teacher(self, config):
        """base_ola_vlm.py:84-94: `self.oneformer` (OneFormerHead — only its Swin-L backbone runs in
        training, model/seg_teacher.py) and `self.oneformer_processor` (HF OneFormerProcessor, host-side
        PIL preprocessing; needs the checkpoint directory's preprocessor_config.json).
        Weights: <config.image_segmentor>/{model.safetensors | pytorch_model.bin}, backbone keys only."""
        import os
        from .seg_teacher import OneFormerHead

        net = OneFormerHead(None, self._device)
        root = str(getattr(config, "image_segmentor", ""))
        st, pt = os.path.join(root, "model.safetensors"), os.path.join(root, "pytorch_model.bin")
        sd = None
        if os.path.exists(st):
            from safetensors.torch import load_file
            sd = load_file(st)
        elif os.path.exists(pt):
            sd = torch.load(pt, map_location="cpu")
        if sd is not None:
            sd = {(k[6:] if k.startswith("model.") else k): v for k, v in sd.items()}
            want = set(net.state_dict().keys())
            net.load_state_dict({k: v for k, v in sd.items() if k in want}, strict=True)
        elif getattr(config, "random_init_teachers", False):
            with torch.no_grad():
                for n_, p_ in net.named_parameters():
                    p_.fill_(1.0) if ("norm" in n_ and n_.endswith("weight")) else p_.normal_(0.0, 0.02)
        else:
            raise FileNotFoundError(f"seg teacher weights not found under {root!r} (the reference downloads "
                                    "oneformer/oneformer_coco_swin_large; no network here)")
        self.oneformer = net
        self.oneformer_processor = None
        if os.path.exists(os.path.join(root, "preprocessor_config.json")):
            from transformers import OneFormerProcessor
            self.oneformer_processor = OneFormerProcessor.from_pretrained(root)

    def _layer_loss_weight(self, cfgd, prefix):
        idx = [int(i) - 1 for i in cfgd[f"{prefix}_layer_indices"].split("-")]  # base_ola_vlm.py:97-102
        return idx, cfgd[f"{prefix}_loss_weight"]

    def init_heads(self, config):
        """base_ola_vlm.py:104-168 (task-token heads only: num_task_tokens > 0)."""
        dev = self._device
        self.mode = getattr(config, "aux_mode", "gen-depth-seg")
        self.pass_text_to_aux_head = getattr(config, "pass_text_to_aux", True)
        self.contrastive_loss_weight = config.contrastive_loss_weight
        assert config.num_task_tokens > 0, "only the TaskToken* heads are on the shipped training path"
        D = config.hidden_size
        use_con = getattr(config, "use_contrastive", True)
        if "gen" in self.mode:
            self.img_layer_indices, self.img_gen_loss_weight = self._layer_loss_weight(config.image_gen, "img")
            self.gen_logit_scale = nn.Parameter(torch.tensor(2.0, device=dev)) if use_con else None
            self.image_gen_heads = nn.ModuleList([M.TaskTokenGenHead(config.image_gen, D, dev)
                                                  for _ in self.img_layer_indices])
        if "depth" in self.mode:
            self.depth_layer_indices, self.img_depth_loss_weight = self._layer_loss_weight(config.image_depth, "depth")
            self.depth_logit_scale = nn.Parameter(torch.tensor(2.0, device=dev)) if use_con else None
            self.use_intermediate_depth = config.image_depth.get("use_intermediate_depth", True)
            self.image_depth_heads = nn.ModuleList([
                M.TaskTokenDepthHead(config.image_depth, D, self.use_intermediate_depth, dev)
                for _ in self.depth_layer_indices])
            # frozen DPT decoder behind `depth_preds` (base_ola_vlm.py:141-148 loads it from
            # config.depth_estimator); hard-wired to 1024-channel features like the reference's
            if config.image_depth.get("output_dim") == 1024 and getattr(config, "depth_preds", True):
                from .dpt import DAv2_Head
                self.da_v2_head = DAv2_Head(dev)
                path = getattr(config, "depth_estimator", None)
                if path and __import__("os").path.exists(str(path)):   # base_ola_vlm.py:148 (strict=False)
                    sd = torch.load(path, map_location="cpu")
                    own = self.da_v2_head.state_dict()
                    self.da_v2_head.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=False)
                    self._da_v2_head_loaded = True
                self.da_v2_head.requires_grad_(False)
        if "seg" in self.mode:
            self.seg_layer_indices, self.img_seg_loss_weight = self._layer_loss_weight(config.image_seg, "seg")
            self.seg_logit_scale = nn.Parameter(torch.tensor(2.0, device=dev)) if use_con else None
            self.image_seg_heads = nn.ModuleList([M.OneFormerTaskTokenSegHead(config.image_seg, D, dev)
                                                  for _ in self.seg_layer_indices])

    # ---- weights -----------------------------------------------------------------------------
    @torch.no_grad()
    def init_weights(self, seed_fn=None, std=0.02, seed=0, only=None):
        """Random init (HF std=0.02 style) or deterministic by-name init (tests: seed_fn(name, shape)).
        only: substrings selecting the parameters to touch — `only=self.NEW_MODULE_KEYS` initialises the
        projector / task tokens / heads added on top of a loaded LLM and leaves the loaded weights alone."""
        g = torch.Generator(device="cpu").manual_seed(seed)
        for name, p in self.named_parameters():
            if only is not None and not any(t in name for t in only):
                continue
            if seed_fn is not None:
                p.copy_(seed_fn(name, tuple(p.shape)).to(p.dtype))
                continue
            last = name.split(".")[-1]
            if name.endswith("logit_scale"):
                p.fill_(2.0)
            elif last == "bias":
                p.zero_()
            elif last == "weight" and p.dim() == 1:
                p.fill_(1.0)
            elif "special_" in name:
                p.normal_(0.0, 1.0)
            else:
                p.normal_(0.0, std)
        return self

    # ---- the multimodal splice ---------------------------------------------------------------
    def encode_images(self, images):
        """ola_arch.py:187-190 — frozen tower (no grad) then the trainable mlp2x_gelu projector."""
        feats = self.get_vision_tower()(images)
        if self._pre_trainable_hook is not None:   # ZeRO-2: the updated projector must have arrived by now
            self._pre_trainable_hook(0)
        pj = self.model.mm_projector
        h = A.linear(feats, pj[0].weight, pj[0].bias, ACT_GELU)
        return A.linear(h, pj[2].weight, pj[2].bias, ACT_NONE)

    def _task_rows(self):
        """append_special_tokens: rows appended after each image, in token_order.
        Distillation classes (ola_arch.py:224-254): depth / seg parameters [576, D] pooled to num_task_tokens rows.
        NTP-only classes (llava_arch.py:251-293) — reached when the VPT / IFT stages load a distilled checkpoint,
        whose config still carries the task-token keys: the same pooling for task_token_format "expand_emb", the
        RAW parameters (576 rows each) for "emb", exactly as published."""
        nt = self.num_task_tokens
        if not nt:
            return None
        fmt = getattr(self.model, "task_token_format", "emb")
        if not self.distill and fmt not in ("emb", "expand_emb"):
            raise NotImplementedError(f"task_token_format {fmt!r} (token-id task tokens) is not on the built path")
        pool = self.distill or fmt == "expand_emb"
        rows = []
        for task in self.token_order:
            tok = {"depth": self.depth_tokens, "seg": self.seg_tokens, "gen": self.gen_tokens}[task]
            if tok is None or task not in self.model.aux_tokens:
                continue
            if task == "gen" or not pool:
                rows.append(tok)
            else:
                rows.append(A.GroupMeanFn.apply(tok, nt, tok.shape[0] // nt))
        return torch.cat(rows, 0) if rows else None

    def prepare_inputs_labels_for_multimodal(self, input_ids, position_ids, attention_mask,
                                             past_key_values, labels, images, image_sizes=None):
        """Same signature / return tuple as ola_arch.py:256-259,444."""
        if self.get_vision_tower() is None or images is None or input_ids.shape[1] == 1:
            return input_ids, position_ids, attention_mask, past_key_values, None, labels
        if isinstance(images, (list, tuple)) or images.dim() == 5:
            images = torch.cat([im.unsqueeze(0) if im.dim() == 3 else im for im in images], 0)
        dev = self.device
        image_features = self.encode_images(images.to(dev, non_blocking=True))  # [n_img*576, D]
        n_tok = self.get_vision_tower().num_patches
        if self._pre_trainable_hook is not None:
            self._pre_trainable_hook(1)   # embedding table and task tokens
        task_rows = self._task_rows()
        cpu = lambda t: None if t is None else (t.cpu() if t.is_cuda else t)
        plan = SplicePlan(cpu(input_ids), cpu(labels), cpu(attention_mask), n_tok,
                          0 if task_rows is None else task_rows.shape[0],
                          getattr(self.config, "tokenizer_model_max_length", None),
                          getattr(self.config, "tokenizer_padding_side", "right")).to(dev)
        if task_rows is None:
            task_rows_in = image_features[:0]
        else:
            task_rows_in = task_rows
        embeds = A.SpliceFn.apply(self.model.embed_tokens.weight, image_features, task_rows_in, plan)
        B, T = plan.B, plan.T
        self._last_plan = plan
        new_labels = plan.labels if labels is not None else None
        new_mask = plan.mask.to(attention_mask.dtype) if attention_mask is not None else None
        return None, None if position_ids is None else position_ids, new_mask, past_key_values, \
            embeds.view(B, T, -1), new_labels

    # ---- decoder -----------------------------------------------------------------------------
    def _decoder(self, inputs_embeds):
        cfg = self.config
        B, T, D = inputs_embeds.shape
        H, KVH = cfg.num_attention_heads, cfg.num_key_value_heads
        hd = D // H
        hook = self._pre_trainable_hook
        cos, sin = ops.rope_tables(max(cfg.max_position_embeddings, T), hd, cfg.rope_theta, inputs_embeds.device)
        sw = getattr(cfg, "sliding_window", None)
        meta = SimpleNamespace(B=B, T=T, H=H, KVH=KVH, hd=hd, eps=cfg.rms_norm_eps, cos=cos, sin=sin,
                               pos_ids=None, window=int(sw) if (sw and T > sw + 1) else 0)
        x = inputs_embeds.reshape(B * T, D)
        if x.dtype != BF16:
            x = x.to(BF16)
        states = [x]
        for li, layer in enumerate(self.model.layers):
            if hook is not None:
                hook(2 + li)        # ZeRO-2: wait only for the all-gathers up to this layer's weights
            x = layer.run(x, meta)
            states.append(x)
        if hook is not None:
            hook(None)              # final norm, lm_head, heads: everything
        states[-1] = A.RMSNormFn.apply(x, self.model.norm.weight, cfg.rms_norm_eps)
        return states

    # ---- teachers (hooks kept for drop-in monkeypatching; OUT OF SCOPE by default) --------------
    def _get_dav2_feats(self, pil_images, device, decode=True):
        """base_ola_vlm.py:348-366, batched: one DINOv2 pass over all images instead of a batch-1
        Python loop.  pil_images: PIL images (resized to 336x336 like :351) or uint8 [B,336,336,3].
        Returns ([(targets [B,576,1024], None)], depth_gts [B,336,336] min-max normalised or None
        when the DPT decoder is not loaded)."""
        teacher = getattr(self, "dav2_backbone", None)
        if teacher is None:
            raise NotImplementedError("depth teacher not initialised: call init_target_models(config) or "
                                      "pass distill_targets=")
        if torch.is_tensor(pil_images):
            raw = pil_images
        else:
            import numpy as np
            raw = torch.from_numpy(np.stack([np.array(im.resize((336, 336))) for im in pil_images]))
        B = raw.shape[0]
        ft = teacher.dsg_targets(raw, 336)
        head = getattr(self, "da_v2_head", None)
        # depth_gts only feed the reference's wandb depth logging; the training step skips the decode
        gts = head.normalized([ft] * 4) if (decode and head is not None) else None
        return [(ft.view(B, -1, ft.shape[-1]), None)], gts

    def _seg_pixel_values(self, pil_images):
        if torch.is_tensor(pil_images):
            return pil_images
        proc = getattr(self, "oneformer_processor", None)
        if proc is None:
            raise NotImplementedError("no OneFormerProcessor loaded: pass preprocessed pixel_values")
        return torch.cat([proc(im.resize((768, 768)), ["panoptic"], return_tensors="pt")["pixel_values"]
                          for im in pil_images], 0)            # base_ola_vlm.py:383-386

    def _get_seg_targets(self, pil_images, seg_preds):
        """base_ola_vlm.py:382-397, batched: the Swin-L backbone's last feature map resized to 24x24 →
        [B,1536,24,24].  pil_images: PIL images (through oneformer_processor, as the reference) or an
        already preprocessed pixel_values tensor [B,3,H,W]."""
        net = getattr(self, "oneformer", None)
        if net is None:
            raise NotImplementedError("seg teacher not initialised: call init_target_models(config) or "
                                      "pass distill_targets=")
        return net.forward_features(self._seg_pixel_values(pil_images))

    def _get_gen_feats(self, pil_images, device):
        """base_ola_vlm.py:323-333, batched: image_embeds of the unCLIP image encoder → [B,1,1024].
        pil_images: PIL images (through pipe.feature_extractor, as the reference) or an already
        preprocessed pixel_values tensor [B,3,224,224]."""
        pipe = getattr(self, "pipe", None)
        if pipe is None:
            raise NotImplementedError("gen teacher not initialised: call init_target_models(config) or "
                                      "pass distill_targets=")
        if torch.is_tensor(pil_images):
            px = pil_images
        else:
            px = pipe.feature_extractor(images=list(pil_images), return_tensors="pt").pixel_values
        return pipe.image_encoder.image_embeds(px)[:, None]

    def _targets(self, task, pil_images, distill_targets, device):
        """Caller-supplied targets win; otherwise the task's on-GPU teacher when it is loaded and the
        batch carries images for it (PIL images, or a dict {task: preprocessed tensor}); otherwise the
        reference's hooks."""
        if distill_targets is not None and task in distill_targets:
            return distill_targets[task]
        images = pil_images.get(task) if isinstance(pil_images, dict) else pil_images
        have_images = torch.is_tensor(images) or (images is not None and len(images) > 0
                                                  and images[0] is not None)
        loaded = {"depth": getattr(self, "dav2_backbone", None) is not None,
                  "gen": getattr(self, "pipe", None) is not None,
                  "seg": getattr(self, "oneformer", None) is not None}[task]
        if not (have_images and loaded) and (distill_targets is not None or images is None):
            return None
        if task == "depth":
            return self._get_dav2_feats(images, device, decode=False)[0][0][0]
        if task == "seg":
            net = getattr(self, "oneformer", None)
            if net is None:
                return self._get_seg_targets(images, None)
            rows = net.seg_target_rows(self._seg_pixel_values(images))   # token-major: no NCHW round trip
            return rows.view(-1, 576, rows.shape[-1])
        return self._get_gen_feats(images, device)

    def _gather_targets(self, tgt_flat):
        """dist_collect (ola_utils.py:96-106): targets carry no grad → plain NCCL all-gather, once
        per task per step."""
        return gather_targets(tgt_flat)

    def _emb_loss(self, pred_flat, mask, tgt_all, off, logit_scale):
        """base_ola_vlm.py:289-320 → (loss, sl1, contrastive) as a 3-vector."""
        tau = logit_scale.float().reshape(1)
        m = None if mask is None else mask.float().contiguous()
        return A.DistillLossFn.apply(pred_flat, tgt_all, tau, m, off, float(self.contrastive_loss_weight))

    def _distill(self, states, B, T, pil_images, masks, distill_targets):
        cfg = self.config
        dev = states[0].device
        layer_states = states[1:]
        S, nt = self.NUM_SYS_TOKENS, self.num_task_tokens
        out = dict(depth_embs=[], seg_embs=[], image_embs=[], depth_preds=[], terms={})
        total = None
        spec = [("depth", "depth", "image_depth_heads", "depth_layer_indices", "img_depth_loss_weight",
                 "depth_logit_scale", self.depth_tokens, "depth_embs"),
                ("seg", "seg", "image_seg_heads", "seg_layer_indices", "img_seg_loss_weight",
                 "seg_logit_scale", self.seg_tokens, "seg_embs"),
                ("gen", "gen", "image_gen_heads", "img_layer_indices", "img_gen_loss_weight",
                 "gen_logit_scale", self.gen_tokens, "image_embs")]
        for task, key, heads_name, idx_name, w_name, scale_name, special, out_name in spec:
            if task not in self.mode or T <= S:
                continue
            heads = getattr(self, heads_name)
            weight = getattr(self, w_name)
            tgt = self._targets(task, pil_images, distill_targets, dev)
            tgt_all = off = tgt_ev = None
            if tgt is not None:
                tgt, masks[task] = tile_targets(tgt, masks.get(task), B)
                tgt = tgt.to(dev, BF16)
                if task == "seg" and tgt.dim() == 4:  # [B,C,24,24] → token-major [B,576,C] (head layout)
                    Bc, C = tgt.shape[0], tgt.shape[1]
                    tflat = torch.empty((Bc, 576 * C), dtype=BF16, device=dev)
                    tgt = tgt.contiguous()
                    for b in range(Bc):
                        ops.transpose(tgt[b].view(C, 576), out=tflat[b].view(576, C))
                    tgt_flat = tflat
                else:
                    tgt_flat = tgt.reshape(tgt.shape[0], -1).contiguous()
                tgt_all, off, tgt_ev = self._gather_targets(tgt_flat)
            mask = masks.get(task)
            if tgt is not None and tgt.shape[0] != B:
                raise ValueError(f"{task} targets: {tgt.shape[0]} rows cannot be tiled to a batch of {B}")
            for i, idx in enumerate(getattr(self, idx_name)):
                head = heads[i]
                n_lat = 1 if task == "gen" else special.shape[0]
                plan = HeadPlan.get(B, T, S, nt, self.token_order, task, self.pass_text_to_aux_head,
                                    n_lat, head.projector.heads, dev)
                e = head.projector.run(layer_states[idx], None if task == "gen" else special, plan)
                if task == "depth":
                    feats = [(head.mlp(k, e).view(B, plan.nq, -1), None) for k in (1, 2, 3)] \
                        if head.use_intermediate_depth else []
                    feats.append((e.view(B, plan.nq, -1), None))
                    out[out_name].append(feats)
                    pred = feats[0][0]  # base_ola_vlm.py:369 supervises features[0] only
                    if getattr(self, "da_v2_head", None) is not None and getattr(cfg, "depth_preds", True):
                        # base_ola_vlm.py:462-470 (no_grad): DPT decoder + per-image min-max
                        lv = [f[0] for f in feats] if head.use_intermediate_depth else [feats[0][0]] * 4
                        out["depth_preds"].append(self.da_v2_head.normalized(
                            [f.detach().reshape(B * plan.nq, -1) for f in lv]))
                elif task == "seg":
                    ev = e.view(B, plan.nq, -1)
                    side = int(math.sqrt(plan.nq))
                    out[out_name].append(ev.permute(0, 2, 1).unflatten(2, (side, side)))
                    pred = ev
                else:
                    pred = e.view(B, plan.nq, -1)
                    out[out_name].append(pred)
                if mask is not None and cfg.zero_masks_like_reference:
                    mask.zero_()  # base_ola_vlm.py:472-473, 498-499, 525-526
                if tgt_all is not None:
                    if tgt_ev is not None:      # the side-stream all-gather of the targets must have landed
                        torch.cuda.current_stream().wait_event(tgt_ev)
                        tgt_ev = None
                    l3 = self._emb_loss(pred.reshape(B, -1), mask, tgt_all, off, getattr(self, scale_name))
                    out["terms"].setdefault(task, []).append(l3)
                    total = l3[0] * weight if total is None else total + l3[0] * weight
        out["total"] = total
        return out

    # ---- forward -----------------------------------------------------------------------------
    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None,
                inputs_embeds=None, labels=None, use_cache=None, output_attentions=None,
                output_hidden_states=None, images=None, image_sizes=None, return_dict=None,
                pil_images=None, gen_mask=None, seg_mask=None, depth_mask=None, **kwargs):
        """Keyword surface of ola_llama.py:190-209 (llava_llama.py:73-91 for the NTP-only classes)."""
        distill_targets = kwargs.pop("distill_targets", None)
        if images is None and self._pre_trainable_hook is not None:
            self._pre_trainable_hook(1)  # no frozen tower to hide the parameter all-gather behind
        if inputs_embeds is None:
            (input_ids, position_ids, attention_mask, past_key_values, inputs_embeds,
             labels) = self.prepare_inputs_labels_for_multimodal(
                input_ids, position_ids, attention_mask, past_key_values, labels, images, image_sizes)
        if inputs_embeds is None:
            inputs_embeds = ops.gather_rows(input_ids.reshape(-1).to(torch.int32),
                                            [self.model.embed_tokens.weight], self.config.hidden_size
                                            ).view(*input_ids.shape, -1)
        B, T, D = inputs_embeds.shape
        states = self._decoder(inputs_embeds)
        hidden = states[-1]
        loss = text_loss = logits = None
        if labels is not None:
            if labels.device != hidden.device or not labels.is_contiguous():
                labels = labels.to(hidden.device).contiguous()
            V = self.config.vocab_size
            chunk = max(1024, min(B * T, (1 << 31) // V))
            if not hasattr(self, "_lm_head_t"):
                self._lm_head_t = M.FrozenTranspose()
            wt = (self._lm_head_t.get(self.lm_head.weight)
                  if (torch.is_grad_enabled() and getattr(self.config, "frozen_transposes", False)) else None)
            plan = getattr(self, "_last_plan", None)
            compact = (plan.ce_compaction() if (plan is not None and labels is plan.labels
                                                and getattr(self.config, "skip_ignored_ce_rows", True))
                       else None)
            text_loss = A.LMHeadCEFn.apply(hidden, self.lm_head.weight, labels, T, chunk, wt, compact)
        if labels is None or self.config.materialize_logits:
            with torch.no_grad():
                # bf16 GEMM, then upcast — the reference's `logits = self.lm_head(h); logits = logits.float()`
                # (ola_llama.py:121-122): fp32 [B,T,V] whenever the field is produced
                logits = A.lm_head_logits(hidden.detach(), self.lm_head.weight).view(B, T, -1).float()
        d = None
        if self.distill and hasattr(self, "mode"):
            d = self._distill(states, B, T, pil_images,
                              {"depth": depth_mask, "seg": seg_mask, "gen": gen_mask}, distill_targets)
        if text_loss is not None:
            loss = text_loss if (d is None or d["total"] is None) else text_loss + d["total"]
        hs = tuple(s.view(B, T, D) for s in states)
        if self.steps is not None:
            self.steps += 1
        out = OlaCausalLLMOutputWithPast(
            loss=loss, logits=logits, past_key_values=None, hidden_states=hs, attentions=None,
            image_embs=d["image_embs"] if d else None, seg_embs=d["seg_embs"] if d else None,
            depth_embs=d["depth_embs"] if d else None, depth_preds=d["depth_preds"] if d else None,
            text_loss=text_loss, loss_terms=d["terms"] if d else None)
        if return_dict is False:
            return tuple(v for v in (loss, logits, hs) if v is not None)
        return out


class OlaLlavaLlamaForCausalLM(VisperForCausalLM):
    family, distill, config_class = "llama", True, OlaLlavaLlamaConfig


class OlaLlavaPhi3ForCausalLM(VisperForCausalLM):
    family, distill, config_class = "phi3", True, OlaLlavaPhi3Config


class LlavaLlamaForCausalLM(VisperForCausalLM):
    family, distill, config_class = "llama", False, LlavaConfig


class LlavaPhi3ForCausalLM(VisperForCausalLM):
    family, distill, config_class = "phi3", False, LlavaPhi3Config


def _register_with_transformers():
    """AutoConfig.register / AutoModelForCausalLM.register as the reference does at import time
    (ola_llama.py:246-247, ola_phi3.py, llava_llama.py:174-175, llava_phi3.py), so `model_type` strings in a saved
    config.json resolve to these classes through the Auto* factories."""
    try:
        from transformers import AutoConfig, AutoModelForCausalLM
    except Exception:  # transformers absent: the classes work without the factories
        return
    for cfg, cls in ((OlaLlavaLlamaConfig, OlaLlavaLlamaForCausalLM), (OlaLlavaPhi3Config, OlaLlavaPhi3ForCausalLM),
                     (LlavaConfig, LlavaLlamaForCausalLM), (LlavaPhi3Config, LlavaPhi3ForCausalLM)):
        try:
            AutoConfig.register(cfg.model_type, cfg, exist_ok=True)
            AutoModelForCausalLM.register(cfg, cls, exist_ok=True)
        except Exception:  # a different class already owns the name in this process (the shimmed reference in tests)
            pass


_register_with_transformers()
