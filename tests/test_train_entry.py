"""CPU tests of visper_lm_b200.train.entry — the reference's train() (ola_vlm_train.py:977-1326) up to the
point where the first step would run: the launch script's own flags parse, the --layer_indices / --loss_weights
DSLs, and build_model's order of operations on a stock HF checkpoint (class choice, tower choice, tokenizer
pad handling, aux config, fresh-module init, requires_grad policy)."""
import pytest
import torch

# scripts/train/pretrain.sh:16-58 of the reference, verbatim flags
PRETRAIN_SH = """--deepspeed ./scripts/zero2.json --model_name_or_path meta-llama/Meta-Llama-3-8B-Instruct
 --version llava_llama_3 --mode gen-depth-seg --layer_indices d18-20_s10-18_g12-20 --num_task_tokens 8
 --loss_weights d0.5_s0.5_g0.5 --contrastive_loss_weight 0.3 --image_generator stabilityai/stable-diffusion-2-1-unclip
 --image_segmentor shi-labs/oneformer_coco_swin_large --depth_estimator depth_anything_v2_vitl.pth
 --data_path datasets/LLaVA-Pretrain/blip_laion_cc_sbu_558k.json --image_folder datasets/LLaVA-Pretrain/images
 --vision_tower openai/clip-vit-large-patch14-336 --mm_projector_type mlp2x_gelu --tune_mm_mlp_adapter True
 --mm_vision_select_layer -2 --mm_use_im_start_end False --mm_use_im_patch_token False --bf16 True
 --output_dir outputs/pretrain_dsg --num_train_epochs 1 --per_device_train_batch_size 32
 --per_device_eval_batch_size 4 --gradient_accumulation_steps 1 --evaluation_strategy no --save_strategy steps
 --save_steps 200 --save_total_limit 3 --learning_rate 1e-3 --weight_decay 0. --warmup_ratio 0.03
 --lr_scheduler_type cosine --logging_steps 1 --tf32 True --model_max_length 4096 --gradient_checkpointing True
 --dataloader_num_workers 4 --lazy_preprocess True --report_to wandb""".split()


def test_parses_the_reference_launch_flags():
    from visper_lm_b200.train.entry import parse_args

    m, d, t, ignored = parse_args(PRETRAIN_SH)
    assert (m.model_name_or_path, m.version, m.mode) == ("meta-llama/Meta-Llama-3-8B-Instruct", "llava_llama_3", "gen-depth-seg")
    assert (m.layer_indices, m.loss_weights, m.num_task_tokens, m.contrastive_loss_weight) == \
        ("d18-20_s10-18_g12-20", "d0.5_s0.5_g0.5", 8, 0.3)
    assert m.tune_mm_mlp_adapter is True and m.mm_vision_select_layer == -2 and m.mm_projector_type == "mlp2x_gelu"
    assert m.mm_use_im_start_end is False and m.mm_use_im_patch_token is False
    assert (m.task_token_format, m.use_ce, m.sample_tokens) == ("emb", False, False)
    assert d.data_path.endswith("blip_laion_cc_sbu_558k.json") and d.lazy_preprocess is True
    assert (t.per_device_train_batch_size, t.learning_rate, t.weight_decay, t.warmup_ratio) == (32, 1e-3, 0.0, 0.03)
    assert (t.save_steps, t.save_total_limit, t.model_max_length, t.dataloader_num_workers, t.bf16) == (200, 3, 4096, 4, True)
    assert set(ignored) == {"--deepspeed", "--per_device_eval_batch_size", "--evaluation_strategy", "--save_strategy",
                            "--tf32", "--gradient_checkpointing", "--report_to"}
    with pytest.raises(NotImplementedError):
        parse_args(["--deepspeed", "./scripts/zero3.json"])


def test_dsl_parsing_matches_the_reference_rules():
    """ola_vlm_train.py:1159-1194: unknown letters ignored, missing tasks default to layer '0' / weight 0.5."""
    from visper_lm_b200.train.entry import parse_dsl

    assert parse_dsl("d18-20_s10-18_g12-20", "d0.5_s0.5_g0.5") == ({"d": "18-20", "s": "10-18", "g": "12-20"},
                                                                   {"d": 0.5, "s": 0.5, "g": 0.5})
    assert parse_dsl("d8_g12-20", "s0.25_g1.0") == ({"d": "8", "s": "0", "g": "12-20"}, {"d": 0.5, "s": 0.25, "g": 1.0})
    li, lw = parse_dsl("x3_d4", "d1")          # 'x' is not a task; 'd1' has no decimal point → not a weight
    assert li == {"d": "4", "s": "0", "g": "0"} and lw == {"d": 0.5, "s": 0.5, "g": 0.5}


class _Tok:
    """The tokenizer surface train() touches (ola_vlm_train.py:1077-1092)."""
    padding_side = "right"
    model_max_length = 512
    unk_token = None

    def __init__(self, n, pad=None):
        self.n, self.pad_token, self.pad_token_id = n, pad, (n - 1 if pad else None)

    def add_special_tokens(self, d):
        self.pad_token, self.pad_token_id = d["pad_token"], self.n
        self.n += 1
        return 1

    def add_tokens(self, toks, special_tokens=False):
        self.n += len(toks)
        return len(toks)

    def __len__(self):
        return self.n


def _tiny_llm(path, layers=4):
    from transformers import LlamaConfig, LlamaForCausalLM

    torch.manual_seed(0)
    hf = LlamaForCausalLM(LlamaConfig(vocab_size=160, hidden_size=64, intermediate_size=128, num_hidden_layers=layers,
                                      num_attention_heads=4, num_key_value_heads=2, tie_word_embeddings=False))
    hf.save_pretrained(path, safe_serialization=True)
    return hf


def _args(path, **kw):
    from visper_lm_b200.train.data import DataArguments
    from visper_lm_b200.train.entry import ModelArguments
    from visper_lm_b200.train.trainer import TrainingArguments

    m = ModelArguments(model_name_or_path=str(path), version="llava_llama_3", vision_tower="openai/clip-vit-large-patch14-336",
                       mm_projector_type="mlp2x_gelu", mm_vision_select_layer=-2, mm_use_im_patch_token=False,
                       layer_indices="d2-3_s1-2_g2-3", loss_weights="d0.5_s0.25_g1.0", img_head_output_dim=32,
                       seg_head_output_dim=48, depth_head_output_dim=32, random_init_teachers=True)
    for k, v in kw.items():
        setattr(m, k, v)
    return m, DataArguments(image_aspect_ratio="pad"), TrainingArguments()


def test_build_model_pt_stage(tmp_path, monkeypatch):
    from visper_lm_b200.model import OlaLlavaLlamaForCausalLM
    from visper_lm_b200.train.entry import build_model

    hf = _tiny_llm(tmp_path)
    monkeypatch.setattr(OlaLlavaLlamaForCausalLM, "init_target_models", lambda self, cfg: setattr(self, "_teachers", cfg.aux_mode))
    m_args, d_args, t_args = _args(tmp_path, tune_mm_mlp_adapter=True)
    tok = _Tok(160)
    vis = dict(hidden_size=32, intermediate_size=64, num_hidden_layers=3, num_attention_heads=2, image_size=28, patch_size=14)
    model, trainable = build_model(m_args, d_args, t_args, tok, device="cpu", config_overrides={"vision": vis})
    assert type(model) is OlaLlavaLlamaForCausalLM and model._teachers == "gen-depth-seg"
    # tokenizer: no pad token → "<pad>" added, tables resized (to a multiple of 8), loaded rows untouched
    assert tok.pad_token == "<pad>" and tok.pad_token_id == 160 and len(tok) == 161
    emb = model.get_input_embeddings().weight
    assert emb.shape[0] == 168 and torch.equal(emb[:160].float(), hf.model.embed_tokens.weight.to(torch.bfloat16).float())
    assert torch.isfinite(emb.float()).all() and torch.isfinite(model.get_output_embeddings().weight.float()).all()
    cfg = model.config
    assert (cfg.tokenizer_padding_side, cfg.tokenizer_model_max_length, cfg.image_aspect_ratio) == ("right", 512, "pad")
    assert cfg.tune_mm_mlp_adapter is True and t_args.tune_mm_mlp_adapter is True and cfg.use_cache is False
    assert d_args.is_multimodal is True and d_args.version == "llava_llama_3"
    assert cfg.image_depth["depth_layer_indices"] == "2-3" and cfg.image_seg["seg_loss_weight"] == 0.25
    assert cfg.image_gen == {"depth": 1, "dim_head": 32, "num_heads": 4, "num_tokens": 1, "output_dim": 32, "ff_mult": 1,
                             "img_layer_indices": "2-3", "img_loss_weight": 1.0}
    assert (model.img_depth_loss_weight, model.img_seg_loss_weight, model.img_gen_loss_weight) == (0.5, 0.25, 1.0)
    assert model.depth_layer_indices == [1, 2] and model.seg_layer_indices == [0, 1]
    # PT stage: projector + everything created after the freeze trains, the LLM and the tower do not
    names = dict(model.named_parameters())
    assert set(trainable) == {n for n, p in names.items() if p.requires_grad}
    assert all(("mm_projector." in n) or ("_heads." in n) or ("model.special_" in n) or n.endswith("logit_scale") for n in trainable)
    assert {"model.mm_projector.0.weight", "model.special_gen_tokens", "depth_logit_scale",
            "image_seg_heads.1.projector.proj_in.weight"} <= set(trainable)
    assert not names["model.layers.0.mlp.down_proj.weight"].requires_grad and not names["lm_head.weight"].requires_grad
    assert all(torch.isfinite(p.float()).all() for n, p in names.items() if n in trainable)      # fresh modules initialised
    for k, v in hf.state_dict().items():                                                        # loaded LLM untouched
        if "embed_tokens" not in k and "lm_head" not in k:
            assert torch.equal(names[k].float(), v.to(torch.bfloat16).float()), k


def test_build_model_finetune_stage_and_convnext_tower(tmp_path, monkeypatch):
    """finetune.sh regime (tune_mm_mlp_adapter False → the LLM trains) behind the ConvNeXt tower
    (multimodal_encoder/builder.py:10-11 picks it from the name); a real pad token needs no resize."""
    from visper_lm_b200.model import OlaLlavaLlamaForCausalLM
    from visper_lm_b200.model.convnext import CLIPConvNextVisionTower, ProcessorWrapper
    from visper_lm_b200.train.entry import build_model

    _tiny_llm(tmp_path, layers=3)
    monkeypatch.setattr(OlaLlavaLlamaForCausalLM, "init_target_models", lambda self, cfg: None)
    m_args, d_args, t_args = _args(tmp_path, vision_tower="CLIP-convnext_xxlarge-res768", layer_indices="d2_s1_g3",
                                   freeze_task_token=True)
    vis = dict(depths=(1, 1, 1, 1), dims=(64, 64, 64, 128), eps=1e-5, image_size=768)
    model, trainable = build_model(m_args, d_args, t_args, _Tok(160, pad="<pad>"), device="cpu",
                                   config_overrides={"vision": vis})
    tower = model.get_vision_tower()
    assert isinstance(tower, CLIPConvNextVisionTower) and tower.num_patches == 576 and tower.hidden_size == 128
    assert isinstance(d_args.image_processor, ProcessorWrapper) and d_args.image_processor.crop_size["height"] == 768
    assert model.get_model().mm_projector[0].weight.shape == (64, 128)
    assert model.get_input_embeddings().weight.shape[0] == 160
    assert "model.layers.0.self_attn.q_proj.weight" in trainable and "lm_head.weight" in trainable
    assert not any("vision_tower" in n or "model.special_" in n for n in trainable)              # tower frozen, tokens frozen
    with pytest.raises(NotImplementedError):
        build_model(*_args(tmp_path, mm_projector_type="linear"), _Tok(160, pad="<pad>"), device="cpu")
    with pytest.raises(ValueError):
        build_model(*_args(tmp_path, vision_tower=None), _Tok(160, pad="<pad>"), device="cpu")


def test_builder_loads_what_the_trainer_saves(tmp_path):
    """ola_vlm.model.builder.load_pretrained_model (builder.py:26-191) on a directory in the trainer's save
    layout: class from config.json's model_type, the reference's 4-tuple, image-patch token added."""
    from parity_utils import build_product, configs

    from visper_lm_b200.model import OlaLlavaPhi3ForCausalLM
    # (the implementation behind ola_vlm.model.builder; imported directly because the oracle tests of the same
    # session may have pointed the `ola_vlm` namespace at the reference tree — oracle/ref_shim.py)
    from visper_lm_b200.model.loader import load_pretrained_model
    from visper_lm_b200.train import checkpoint as C

    torch.manual_seed(5)
    model = build_product(configs.TINY_PHI3, True, None)
    with torch.no_grad():
        for p in model.parameters():
            p.copy_(torch.randn(p.shape))
    C.save_config(model.config, str(tmp_path))
    C.save_pretrained_weights(model.state_dict(), str(tmp_path))
    V = model.config.vocab_size
    tok = _Tok(V, pad="<pad>")
    tokenizer, again, image_processor, context_len = load_pretrained_model(str(tmp_path), None, "whatever-name",
                                                                          device="cpu", tokenizer=tok)
    assert type(again) is OlaLlavaPhi3ForCausalLM and tokenizer is tok and context_len == 4096
    assert image_processor is again.get_vision_tower().image_processor
    assert len(tok) == V + 1                                           # <im_patch> (mm_use_im_patch_token defaults True)
    rows = again.get_input_embeddings().weight.shape[0]
    assert rows >= V + 1 and rows % 8 == 0
    a, b = model.state_dict(), again.state_dict()
    for k in a:
        if "embed_tokens" in k or "lm_head" in k:
            assert torch.equal(a[k], b[k][:V]), k
        else:
            assert torch.equal(a[k], b[k]), k
    with pytest.raises(NotImplementedError):
        load_pretrained_model(str(tmp_path), "some/base", "lora-x", tokenizer=tok)


def test_freeze_backbone_leaves_lm_head_trainable():
    """--freeze_backbone = `model.model.requires_grad_(False)` (ola_vlm_train.py:1043-1044): decoder frozen,
    lm_head and the later-created projector still train."""
    from parity_utils import configs

    from visper_lm_b200.model import LlavaLlamaForCausalLM, presets
    from visper_lm_b200.train.policy import apply_freeze_policy

    model = LlavaLlamaForCausalLM(presets.from_dict(configs.TINY_LLAMA, distill=False), device="meta")
    names = apply_freeze_policy(model, freeze_backbone=True)
    assert "lm_head.weight" in names and "model.mm_projector.0.weight" in names
    assert not any(n.startswith("model.layers.") or n.startswith("model.embed_tokens") or n == "model.norm.weight" for n in names)
    assert not any("vision_tower" in n for n in names)


def test_ntp_class_loads_a_distilled_checkpoint_and_keeps_the_task_tokens(tmp_path):
    """vpt.sh / finetune.sh: train.py loads the PT output into LlavaLlamaForCausalLM — heads are dropped like
    HF drops unexpected keys, the task tokens stay (llava_arch.py:50-51) and, for task_token_format "emb", are
    appended RAW (576 + 576 + 8 rows, llava_arch.py:259-260)."""
    from parity_utils import build_product, configs

    from visper_lm_b200.model import LlavaLlamaForCausalLM
    from visper_lm_b200.train import checkpoint as C

    torch.manual_seed(7)
    pt = build_product(configs.TINY_LLAMA, True, None)
    with torch.no_grad():
        for p in pt.parameters():
            p.copy_(torch.randn(p.shape))
    C.save_config(pt.config, str(tmp_path))
    C.save_pretrained_weights(pt.state_dict(), str(tmp_path))
    m = LlavaLlamaForCausalLM.from_pretrained(str(tmp_path))
    assert not hasattr(m, "image_depth_heads") and m.num_task_tokens == 8
    assert torch.equal(m.model.special_seg_tokens, pt.model.special_seg_tokens)
    assert torch.equal(m.model.layers[1].mlp.down_proj.weight, pt.model.layers[1].mlp.down_proj.weight)
    rows = m._task_rows()                                   # "emb" → raw parameters, in token_order gen-depth-seg
    assert rows.shape == (8 + 576 + 576, configs.TINY_LLAMA["hidden"])
    assert torch.equal(rows[:8], m.model.special_gen_tokens) and torch.equal(rows[8:584], m.model.special_depth_tokens)
    m.model.task_token_format = "text"
    with pytest.raises(NotImplementedError):
        m._task_rows()


def test_stage_chaining_keeps_trained_projector_and_pt_output_is_loadable(tmp_path, monkeypatch):
    """(a) the PT stage's end-of-run save also writes the FULL model (the reference falls through to
    trainer.save_model under DeepSpeed, ola_vlm_train.py:249-251), so the next stage can from_pretrained() it;
    (b) loading such a multimodal checkpoint must keep its projector (llava_arch.py:126-127 builds one only if
    missing) instead of re-initialising it; (c) a tower that is neither a local directory nor in the checkpoint
    raises instead of training on uninitialised memory."""
    import types

    from visper_lm_b200.model import LlavaLlamaForCausalLM, OlaLlavaLlamaForCausalLM
    from visper_lm_b200.train.checkpoint import safe_save_model_for_hf_trainer
    from visper_lm_b200.train.entry import build_model

    llm = tmp_path / "llm"
    _tiny_llm(llm)
    monkeypatch.setattr(OlaLlavaLlamaForCausalLM, "init_target_models", lambda self, cfg: None)
    vis = dict(hidden_size=32, intermediate_size=64, num_hidden_layers=3, num_attention_heads=2, image_size=28, patch_size=14)
    m_args, d_args, t_args = _args(llm, tune_mm_mlp_adapter=True)
    model, _ = build_model(m_args, d_args, t_args, _Tok(160, pad="<pad>"), device="cpu", config_overrides={"vision": vis})
    with torch.no_grad():   # "training": make the projector and a task token recognisable
        model.model.mm_projector[0].weight.fill_(0.125)
        model.model.special_gen_tokens.fill_(-0.5)
    out = tmp_path / "pretrain_out"
    tr = types.SimpleNamespace(args=t_args, model=model, rank=0, optimizer=None)
    from visper_lm_b200.train.trainer import LLaVATrainer
    tr._save = lambda d, state_dict=None: LLaVATrainer._save(tr, d, state_dict)
    safe_save_model_for_hf_trainer(trainer=tr, output_dir=str(out))
    assert (out / "mm_projector.bin").exists() and (out / "config.json").exists()
    assert (out / "model.safetensors").exists() or (out / "model.safetensors.index.json").exists()
    # next stage (finetune.sh → train.py): NTP-only class on the PT output, no --pretrain_mm_mlp_adapter
    m2, d2, t2 = _args(out, tune_mm_mlp_adapter=False, random_init_teachers=False)
    nxt, _ = build_model(m2, d2, t2, _Tok(160, pad="<pad>"), device="cpu", distill=False, config_overrides={"vision": vis})
    assert type(nxt) is LlavaLlamaForCausalLM
    assert torch.equal(nxt.model.mm_projector[0].weight, torch.full_like(nxt.model.mm_projector[0].weight, 0.125)), \
        "the trained projector was re-initialised"
    assert torch.equal(nxt.model.special_gen_tokens, torch.full_like(nxt.model.special_gen_tokens, -0.5))
    # (c) base LLM + hub id for the tower, no random-init flag → loud failure
    m3, d3, t3 = _args(llm, random_init_teachers=False)
    with pytest.raises(FileNotFoundError, match="vision tower"):
        build_model(m3, d3, t3, _Tok(160, pad="<pad>"), device="cpu", config_overrides={"vision": vis})
