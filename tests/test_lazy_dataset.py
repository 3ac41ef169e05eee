"""LazySupervisedDataset / make_supervised_data_module against the reference's own
class (ola_vlm_train.py:774-878, extracted from the source) on a temporary LLaVA-style json + image
folder: item schema, pad-to-square, black image + zero masks for text-only samples, length properties."""
import copy
import json
import types

import numpy as np
import pytest
import torch
from PIL import Image

from parity_utils import ROOT  # noqa: F401
from oracle import ref_functions, ref_shim
from test_data_pipeline import MarkerTokenizer
from visper_lm_b200.train import data as D


class ToyProcessor:
    """Stands in for CLIPImageProcessor: crop_size / image_mean attributes and preprocess()."""
    crop_size = {"height": 12, "width": 12}
    image_mean = [0.48145466, 0.4578275, 0.40821073]

    def preprocess(self, image, return_tensors="pt"):
        a = np.asarray(image.resize((12, 12)), dtype=np.float32) / 255.0
        return {"pixel_values": [torch.from_numpy(a).permute(2, 0, 1)]}


def _write(tmp_path, jsonl):
    rng = np.random.default_rng(0)
    for name, (w, h) in (("wide.png", (30, 18)), ("tall.png", (14, 26)), ("sq.png", (20, 20))):
        Image.fromarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8)).save(tmp_path / name)
    conv = lambda q, a: [{"from": "human", "value": q}, {"from": "gpt", "value": a}]  # noqa: E731
    data = [{"image": "wide.png", "conversations": conv("What is shown <image> here ?", "A wide picture of noise.")},
            {"conversations": conv("Say hello .", "Hello !")},
            {"image": "tall.png", "conversations": conv("<image>\nDescribe it .", "Tall noise.") + conv("More ?", "No.")},
            {"image": "sq.png", "conversations": conv("Square ? <image>", "Yes , square.")}]
    path = tmp_path / ("d.jsonl" if jsonl else "d.json")
    with open(path, "w") as fh:
        if jsonl:
            fh.write("\n".join(json.dumps(d) for d in data))
        else:
            json.dump(data, fh)
    return str(path), data


@pytest.mark.parametrize("version,aspect,jsonl", [("llava_llama_3", "pad", False), ("llava_phi_3", "square", True)])
def test_items_equal_reference_class(tmp_path, version, aspect, jsonl):
    path, data = _write(tmp_path, jsonl)
    tok = MarkerTokenizer()
    args = D.DataArguments(data_path=path, is_multimodal=True, image_folder=str(tmp_path), image_aspect_ratio=aspect,
                           image_processor=ToyProcessor(), version=version)
    mod = D.make_supervised_data_module(tok, args)
    ds = mod["train_dataset"]
    assert len(ds) == 4 and mod["eval_dataset"] is None
    assert ds.modality_lengths[1] < 0 < ds.modality_lengths[0] and ds.lengths[0] == ds.modality_lengths[0] + 128
    items = [ds[i] for i in range(4)]
    for it, d in zip(items, data):
        assert it["image"].shape == (3, 12, 12) and it["pil_image"].mode == "RGB"
        assert it["seg_mask"] == it["depth_mask"] == it["gen_mask"] == int("image" in d)
        assert (it["input_ids"] == D.IMAGE_TOKEN_INDEX).sum().item() == int("image" in d)
    assert float(items[1]["image"].abs().max()) == 0.0 and items[1]["pil_image"].size == (12, 12)
    batch = mod["data_collator"](items)
    assert batch["images"].shape == (4, 3, 12, 12) and batch["seg_mask"].tolist() == [1, 0, 1, 1]
    if not ref_shim.available():
        pytest.skip("/root/reference not mounted")
    v = "llama3" if "llama" in version else "phi3"
    fns = ref_functions.preprocess_fns(v)
    lib = ref_functions.conversation_lib(v)

    def preprocess(sources, tokenizer, has_image=False):   # the dispatcher at ola_vlm_train.py:717-732
        return fns["preprocess_llama_3" if v == "llama3" else "preprocess_phi_3"](sources, tokenizer, has_image=has_image)

    import os
    g = {"Dataset": torch.utils.data.Dataset, "json": json, "os": os, "copy": copy, "Image": Image,
         "rank0_print": lambda *a: None, "read_jsonl": D.read_jsonl, "preprocess": preprocess,
         "preprocess_multimodal": fns["preprocess_multimodal"], "DataArguments": object, "conversation_lib": lib}
    Ref = ref_functions.extract("ola_vlm/train/ola_vlm_train.py", ["LazySupervisedDataset"], g)["LazySupervisedDataset"]
    rargs = types.SimpleNamespace(is_multimodal=True, image_folder=str(tmp_path), image_aspect_ratio=aspect,
                                  image_processor=ToyProcessor(), mm_use_im_start_end=False)
    ref = Ref(path, tok, rargs)
    assert ref.lengths == ds.lengths and ref.modality_lengths == ds.modality_lengths
    for i in range(4):
        a, b = ref[i], items[i]
        assert sorted(a) == sorted(b)
        assert torch.equal(a["input_ids"], b["input_ids"]) and torch.equal(a["labels"], b["labels"])
        assert torch.equal(a["image"], b["image"])
        assert a["pil_image"].size == b["pil_image"].size and a["pil_image"].tobytes() == b["pil_image"].tobytes()
        assert (a["seg_mask"], a["depth_mask"], a["gen_mask"]) == (b["seg_mask"], b["depth_mask"], b["gen_mask"])


def test_expand2square():
    im = Image.new("RGB", (6, 2), (9, 9, 9))
    out = D.expand2square(im, (1, 2, 3))
    assert out.size == (6, 6) and out.getpixel((0, 0)) == (1, 2, 3) and out.getpixel((0, 2)) == (9, 9, 9)
    assert D.expand2square(Image.new("RGB", (4, 4)), (0, 0, 0)).size == (4, 4)


def test_dataset_with_the_convnext_processor(tmp_path):
    """The ConvNeXt tower's image_processor (ProcessorWrapper, base_encoder.py:8-40) through the dataset:
    pad-to-square with its image_mean, 64x64 pixel tensors, black image of crop_size for the text-only sample."""
    from visper_lm_b200.model.convnext import OpenClipEvalTransform, ProcessorWrapper

    path, data = _write(tmp_path, False)
    proc = ProcessorWrapper(OpenClipEvalTransform(64), height=64, width=64)
    args = D.DataArguments(data_path=path, is_multimodal=True, image_folder=str(tmp_path), image_aspect_ratio="pad",
                           image_processor=proc, version="llava_llama_3")
    ds = D.make_supervised_data_module(MarkerTokenizer(), args)["train_dataset"]
    items = [ds[i] for i in range(4)]
    assert all(it["image"].shape == (3, 64, 64) and it["image"].dtype == torch.float32 for it in items)
    assert float(items[1]["image"].abs().max()) == 0.0 and items[1]["seg_mask"] == 0       # text-only: black image
    # padded-to-square wide image: top rows are the CLIP mean colour → ~0 after normalisation
    assert float(items[0]["image"][:, :6].abs().max()) < 0.02 and float(items[0]["image"].abs().max()) > 0.5
