"""TEST / BENCH INFRASTRUCTURE ONLY — imports the UNMODIFIED reference classes from /root/reference
or, where that is not mounted (the GPU box), from the byte-identical copy oracle/build_ref.py placed under
the git-ignored oracle/_ref/ (hashes in oracle/ref_manifest.json).

Used (a) to validate oracle/restate.py, (b) to generate the golden vectors under tests/golden/
(oracle/make_golden.py) and (c) by bench.py's reference / cpu_baseline legs (oracle/ref_run.py).
Nothing in the product path imports it.

The reference cannot be imported as-is here: its package __init__ eagerly imports
open_clip / diffusers / diffdist / matplotlib (absent) and four private transformers names removed
in transformers 5.x.  The recipe below bypasses the eager __init__s with namespace packages, mocks
the absent third-party modules, and patches only *loading / teacher / logging* code — the hot-path
arithmetic (ola_llama.py, ola_arch.py, base_ola_vlm.py, resampler.py, aux heads, ola_utils.py,
clip_encoder.py) runs exactly as published.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from unittest.mock import MagicMock

def _ref_root() -> str:
    env = os.environ.get("VISPER_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/ola_vlm"):
        return "/root/reference"
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


REF_ROOT = _ref_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "ola_vlm"))


_loaded = None


def load():
    """Returns a namespace with the reference classes. Idempotent."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found under {REF_ROOT}")
    import torch
    import torch.distributed as dist
    import transformers  # noqa: F401  (must come first)
    import transformers.modeling_utils as mu

    def ns(name, rel):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF_ROOT, rel)]
        sys.modules[name] = m
        return m

    ns("ola_vlm", "ola_vlm")
    ns("ola_vlm.model", "ola_vlm/model")
    ns("ola_vlm.model.language_model", "ola_vlm/model/language_model")
    ns("ola_vlm.model.multimodal_encoder", "ola_vlm/model/multimodal_encoder")
    ns("ola_vlm.model.multimodal_projector", "ola_vlm/model/multimodal_projector")
    aux = ns("ola_vlm.model.aux_heads", "ola_vlm/model/aux_heads")
    ns("ola_vlm.train", "ola_vlm/train")

    for name in ["open_clip", "open_clip.model", "open_clip.coca_model", "open_clip.openai",
                 "open_clip.pretrained", "open_clip.transform", "open_clip.factory", "timm",
                 "timm.models", "timm.models.convnext", "timm.layers", "matplotlib",
                 "matplotlib.pyplot", "matplotlib.cm", "diffdist", "diffdist.functional",
                 "diffusers", "icecream", "wandb", "cv2", "s2wrapper"]:
        if name not in sys.modules:
            sys.modules[name] = MagicMock()
    for missing in ["set_initialized_submodules", "_load_state_dict_into_model",
                    "_load_state_dict_into_meta_model", "get_disk_only_shard_files"]:
        if not hasattr(mu, missing):
            setattr(mu, missing, lambda *a, **k: None)

    # differentiable all_gather standing in for diffdist.functional.all_gather (ola_utils.py:105)
    def _all_gather(out_list, x):
        if dist.is_initialized() and dist.get_world_size() > 1:
            import torch.distributed.nn.functional as dnf

            return list(dnf.all_gather(x))
        return [x]

    sys.modules["diffdist.functional"].all_gather = _all_gather
    sys.modules["diffdist"].functional = sys.modules["diffdist.functional"]

    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29531")
        dist.init_process_group("gloo", rank=0, world_size=1)

    gen_head = importlib.import_module("ola_vlm.model.aux_heads.gen_head")
    da_head = importlib.import_module("ola_vlm.model.aux_heads.da_v2_head")
    of_head = importlib.import_module("ola_vlm.model.aux_heads.oneformer_head")
    for mod in (gen_head, da_head, of_head):
        for k, v in vars(mod).items():
            if not k.startswith("_"):
                setattr(aux, k, v)

    ola_llama = importlib.import_module("ola_vlm.model.language_model.ola_llama")
    ola_phi3 = importlib.import_module("ola_vlm.model.language_model.ola_phi3")
    llava_llama = importlib.import_module("ola_vlm.model.language_model.llava_llama")
    llava_phi3 = importlib.import_module("ola_vlm.model.language_model.llava_phi3")
    base = importlib.import_module("ola_vlm.model.language_model.base_ola_vlm")
    clip_enc = importlib.import_module("ola_vlm.model.multimodal_encoder.clip_encoder")
    resampler = importlib.import_module("ola_vlm.model.multimodal_projector.resampler")
    ola_utils = importlib.import_module("ola_vlm.ola_utils")

    out = types.SimpleNamespace(
        torch=torch, ola_llama=ola_llama, ola_phi3=ola_phi3, llava_llama=llava_llama,
        llava_phi3=llava_phi3, base=base, clip_encoder=clip_enc, resampler=resampler,
        ola_utils=ola_utils, da_head=da_head, gen_head=gen_head, of_head=of_head)
    _loaded = out
    return out


def build_reference_model(cfg: dict, family: str = "llama", distill: bool = True, seed_fn=None,
                          ntp_task_token_format=None, attn_implementation: str = "eager", device=None,
                          dtype=None, fast_init: bool = False, train_mode: bool = False):
    """Instantiate the reference's own OlaLlava*/Llava* class on a (tiny or full) config.

    cfg keys: see oracle.configs.  Weights are then overwritten *by name* from `seed_fn(name, shape)`
    so the product can reproduce them through the state-dict ABI.

    Timing runs (oracle/ref_run.py) pass `fast_init=True` (skip HF's per-module normal_ init of billions of
    parameters; a cheap uniform fill of the same scale instead — values do not affect timing), `device` /
    `dtype` (construction happens on that device) and `attn_implementation` ("sdpa" on CPU,
    "flash_attention_2" on GPU as ola_vlm/train/ola_vlm_train_mem.py:5 selects it).
    """
    R = load()
    import tempfile

    import torch
    from transformers import CLIPVisionConfig, CLIPVisionModel

    vis_cfg = CLIPVisionConfig(
        hidden_size=cfg["vis_hidden"], intermediate_size=cfg["vis_inter"],
        num_hidden_layers=cfg["vis_layers"], num_attention_heads=cfg["vis_heads"],
        image_size=cfg["image_size"], patch_size=cfg["patch_size"], hidden_act="quick_gelu",
        layer_norm_eps=1e-5, projection_dim=cfg["vis_hidden"])

    def _load_model(self, device_map=None):
        tower = CLIPVisionModel(vis_cfg)
        tower.requires_grad_(False)

        # explicit layer loop: the hook-based hidden-state capture of transformers 5.x double-registers
        # on the nested tower after the first full-model forward
        def explicit_forward(pixel_values, output_hidden_states=False, **kw):
            vm = tower.vision_model
            hs = vm.embeddings(pixel_values)
            hs = vm.pre_layrnorm(hs)
            states = [hs]
            for layer in vm.encoder.layers:
                out = layer(hs, None, None) if _layer_takes_causal(layer) else layer(hs, None)
                hs = out[0] if isinstance(out, tuple) else out
                states.append(hs)
            return types.SimpleNamespace(hidden_states=tuple(states), last_hidden_state=hs)

        tower.forward = explicit_forward
        self.vision_tower = tower
        self.image_processor = None
        self.is_loaded = True

    R.clip_encoder.CLIPVisionTower.load_model = _load_model

    if family == "llama":
        mod = R.ola_llama if distill else R.llava_llama
        Cfg = mod.OlaLlavaLlamaConfig if distill else mod.LlavaConfig
        Cls = mod.OlaLlavaLlamaForCausalLM if distill else mod.LlavaLlamaForCausalLM
        lm_kwargs = dict(rope_theta=cfg["rope_theta"], num_key_value_heads=cfg["kv_heads"])
    else:
        mod = R.ola_phi3 if distill else R.llava_phi3
        Cfg = mod.OlaLlavaPhi3Config if distill else mod.LlavaPhi3Config
        Cls = mod.OlaLlavaPhi3ForCausalLM if distill else mod.LlavaPhi3ForCausalLM
        lm_kwargs = dict(rope_theta=cfg["rope_theta"], num_key_value_heads=cfg["kv_heads"],
                         # transformers 5.5 (this image) shows a query W keys including itself
                         # (i-j < W); the reference's pin 4.41.1 / flash_attn shows W+1 (i-j <= W).
                         # The shim therefore asks 5.5 for W+1 to reproduce the pinned behaviour.
                         sliding_window=(cfg["sliding_window"] + 1) if cfg.get("sliding_window") else None,
                         resid_pdrop=0.0, embd_pdrop=0.0, attention_dropout=0.0,
                         pad_token_id=0, bos_token_id=1, eos_token_id=2,
                         original_max_position_embeddings=cfg["max_pos"])
    config = Cfg(
        vocab_size=cfg["vocab"], hidden_size=cfg["hidden"], intermediate_size=cfg["inter"],
        num_hidden_layers=cfg["layers"], num_attention_heads=cfg["heads"],
        max_position_embeddings=cfg["max_pos"], rms_norm_eps=1e-5, tie_word_embeddings=False,
        attn_implementation=attn_implementation, **lm_kwargs)
    config.mm_vision_tower = "synthetic-clip"
    config.mm_vision_select_layer = -2
    config.mm_vision_select_feature = "patch"
    config.mm_projector_type = "mlp2x_gelu"
    config.mm_hidden_size = cfg["vis_hidden"]
    config.mm_patch_merge_type = "flat"
    config.tokenizer_model_max_length = cfg["max_pos"]
    config.tokenizer_padding_side = "right"
    config.use_cache = False
    if distill:
        # mirrors the config injection of ola_vlm/train/ola_vlm_train.py:1149-1229
        config.aux_mode = cfg.get("aux_mode", "gen-depth-seg")
        config.contrastive_loss_weight = 0.3
        config.num_task_tokens = 8
        config.task_token_format = "emb"
        config.pass_text_to_aux = True
        config.use_contrastive = True
        config.use_ce = False
        config.sample_tokens = False

        def head(prefix, nt, od, li):
            return {"depth": 1, "dim_head": 32, "num_heads": 4, "num_tokens": nt, "output_dim": od,
                    "ff_mult": 1, f"{prefix}_layer_indices": li, f"{prefix}_loss_weight": 0.5}

        config.image_gen = head("img", 1, cfg["gen_dim"], cfg["gen_layers"])
        config.image_seg = head("seg", 576, cfg["seg_dim"], cfg["seg_layers"])
        config.image_seg["seg_teacher"] = "oneformer"
        config.image_depth = head("depth", 576, cfg["depth_dim"], cfg["depth_layers"])
        tmp = tempfile.NamedTemporaryFile(suffix=".pth", delete=False)
        torch.save(R.da_head.DAv2_Head().state_dict(), tmp.name)
        config.depth_estimator = tmp.name
        config.image_generator = "none"
        config.image_segmentor = "none"
    if not distill and ntp_task_token_format is not None:
        # the VPT / IFT stages (scripts/train/vpt.sh, finetune.sh → train.py) load a distilled PT checkpoint into
        # the NTP-only class: its config still carries the task-token keys, so LlavaMetaModel.__init__
        # (llava_arch.py:50-51) creates the special tokens and append_special_tokens (:251-293) splices them
        config.aux_mode = cfg.get("aux_mode", "gen-depth-seg")
        config.num_task_tokens = 8
        config.task_token_format = ntp_task_token_format
        config.sample_tokens = False
        config.image_seg = {"num_tokens": 576}
        config.image_depth = {"num_tokens": 576}
    torch.manual_seed(0)
    import contextlib

    ctx = contextlib.ExitStack()
    if fast_init:
        from transformers.initialization import no_init_weights

        ctx.enter_context(no_init_weights())
    if device is not None:
        ctx.enter_context(torch.device(device))
    with ctx:
        model = Cls(config)
    model.steps = 1  # skip the `steps % 1000 == 0` wandb image logging
    if distill:
        model.img_gen_loss_weight = 0.5
        model.img_seg_loss_weight = 0.5
        model.img_depth_loss_weight = 0.5
    if distill and cfg["depth_dim"] != 1024:
        # the frozen DPT decoder ("next" row) is hard-wired to 1024 channels; its
        # output only feeds the logging-only `depth_preds` field, so tiny configs stub it out.
        class _NoDPT(torch.nn.Module):
            def forward(self, feats):
                f = feats[0][0]
                return torch.zeros(f.shape[0], 336, 336, dtype=f.dtype, device=f.device)

        model.da_v2_head = _NoDPT()
    model = model.to(dtype or torch.float32)
    # no dropout anywhere on the path; eval() only disables HF's gradient-checkpointing hooks
    model = model.train() if train_mode else model.eval()
    if fast_init:
        with torch.no_grad():
            for name, p in model.named_parameters():
                if "da_v2_head" in name:
                    continue
                if p.dim() >= 2:
                    p.uniform_(-0.0346, 0.0346)  # std 0.02
                elif name.endswith("bias"):
                    p.zero_()
                elif "norm" in name and name.endswith("weight"):
                    p.fill_(1.0)
                elif name.endswith("logit_scale"):
                    p.fill_(2.0)
                else:
                    p.uniform_(-0.0346, 0.0346)
    if seed_fn is not None:
        with torch.no_grad():
            for name, p in list(model.named_parameters()) + list(model.named_buffers()):
                if "da_v2_head" in name or not p.dtype.is_floating_point:
                    continue
                if "rotary_emb" in name or "position_ids" in name:
                    continue
                p.copy_(seed_fn(name, tuple(p.shape)))
    return model


def _layer_takes_causal(layer) -> bool:
    import inspect

    params = inspect.signature(layer.forward).parameters
    return "causal_attention_mask" in params


def install_synthetic_teachers(model, targets: dict):
    """Monkeypatch the frozen teachers (OUT OF SCOPE, base_ola_vlm.py:323-397) to return fixed
    synthetic targets: targets = {"depth": [B,576,Dd], "seg": [B,Ds,24,24], "gen": [B,1,Dg]}."""
    import torch

    B = targets["depth"].shape[0]

    def _dav2(self, pil_images, device):
        return [(targets["depth"].to(device), None)], torch.zeros(B, 336, 336, device=device)

    def _seg(self, pil_images, seg_preds):
        return targets["seg"].to(seg_preds.device)

    def _gen(self, pil_images, device):
        return targets["gen"].to(device)

    cls = type(model)
    cls._get_dav2_feats = _dav2
    cls._get_seg_targets = _seg
    cls._get_gen_feats = _gen
