"""CPU test: the C-ABI shared library builds, loads, and exports every symbol include/visper_b200.h
declares (no compute without a GPU)."""
import ctypes

from visper_lm_b200 import build, lib


def test_library_builds_and_exports_header_symbols():
    path = build.build()
    assert path.exists()
    decls = lib.parse_header()
    assert len(decls) >= 35
    handle = ctypes.CDLL(str(path))
    missing = [n for n in decls if not hasattr(handle, n)]
    assert not missing, missing
    L = lib.load()
    assert L.vpb_abi_version() == 1
    assert lib.launch_count() >= 0


def test_errors_are_reported_not_swallowed():
    L = lib.load()
    # an invalid problem must be rejected before any launch: negative status + message
    rc = L.vpb_gemm_bf16(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0)
    assert rc != 0
    assert b"gemm" in L.vpb_last_error()


def test_product_never_imports_oracle():
    import pathlib
    import re

    root = pathlib.Path(lib.__file__).resolve().parent
    for f in root.rglob("*.py"):
        assert not re.search(r"^\s*(from|import)\s+oracle\b", f.read_text(), re.M), f


def test_reference_module_paths_resolve():
    """ola_vlm.model / ola_vlm.train.llava_trainer export the names the reference's train() imports."""
    import importlib
    import sys

    for k in [k for k in sys.modules if k == "ola_vlm" or k.startswith("ola_vlm.")]:
        del sys.modules[k]  # the oracle shim may have registered namespace stubs in this process
    m = importlib.import_module("ola_vlm.model")
    for name in ("OlaLlavaLlamaForCausalLM", "OlaLlavaPhi3ForCausalLM", "LlavaLlamaForCausalLM",
                 "LlavaPhi3ForCausalLM", "OlaLlavaLlamaConfig", "LlavaConfig"):
        assert hasattr(m, name)
    assert m.OlaLlavaLlamaConfig.model_type == "ola_llama" and m.LlavaConfig.model_type == "llava_llama"
    t = importlib.import_module("ola_vlm.train.llava_trainer")
    assert hasattr(t, "LLaVATrainer")
    a = importlib.import_module("ola_vlm.model.aux_heads")       # base_ola_vlm.py:13-15 imports
    for name in ("DAv2_Head", "TaskTokenGenHead", "TaskTokenDepthHead", "OneFormerTaskTokenSegHead", "OneFormerHead"):
        assert hasattr(a, name)
    assert hasattr(importlib.import_module("ola_vlm.model.aux_heads.depth_anything_v2.dpt"), "DepthAnythingV2")
    tr = importlib.import_module("ola_vlm.train.ola_vlm_train")                # what pretrain.sh / finetune.sh launch
    assert callable(tr.train) and hasattr(tr, "ModelArguments") and hasattr(tr, "TrainingArguments")
    assert callable(importlib.import_module("ola_vlm.train.ola_vlm_train_mem").train)
    assert callable(importlib.import_module("ola_vlm.train.train_mem").train)
    assert callable(importlib.import_module("ola_vlm.model.builder").load_pretrained_model)
    b = importlib.import_module("ola_vlm.model.multimodal_encoder.builder")   # ola_arch.py:12 import
    from types import SimpleNamespace

    t = b.build_vision_tower(SimpleNamespace(mm_vision_tower="CLIP-convnext_xxlarge-res768", mm_vision_select_layer=-2),
                             device="meta")
    assert type(t).__name__ == "CLIPConvNextVisionTower" and t.hidden_size == 3072 and t.num_patches == 576
    t = b.build_vision_tower(SimpleNamespace(mm_vision_tower="openai/clip-vit-large-patch14-336", mm_vision_select_layer=-2),
                             device="meta")
    assert type(t).__name__ == "CLIPVisionTower" and t.hidden_size == 1024 and t.num_patches == 576
    import pytest

    with pytest.raises(ValueError):
        b.build_vision_tower(SimpleNamespace(mm_vision_tower="sam-vit"))
    for k in [k for k in sys.modules if k == "ola_vlm" or k.startswith("ola_vlm.")]:
        del sys.modules[k]


def test_auto_factories_know_the_model_types():
    """ola_llama.py:246-247 / llava_llama.py:174-175: the four model_type strings resolve through transformers'
    Auto* factories to this package's classes."""
    from transformers import AutoConfig, AutoModelForCausalLM

    from visper_lm_b200.model import vlm

    for mt, cfg, cls in (("ola_llama", vlm.OlaLlavaLlamaConfig, vlm.OlaLlavaLlamaForCausalLM),
                         ("ola_phi3", vlm.OlaLlavaPhi3Config, vlm.OlaLlavaPhi3ForCausalLM),
                         ("llava_llama", vlm.LlavaConfig, vlm.LlavaLlamaForCausalLM),
                         ("llava_phi3", vlm.LlavaPhi3Config, vlm.LlavaPhi3ForCausalLM)):
        c = AutoConfig.for_model(mt, hidden_size=128, num_hidden_layers=2)
        assert type(c) is cfg and c.model_type == mt and c.hidden_size == 128
        assert AutoModelForCausalLM._model_mapping[cfg] is cls
