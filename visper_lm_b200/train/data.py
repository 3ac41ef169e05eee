"""Host-side input pipeline of the training path, index logic only:

  * the (modality-)length-grouped sampler behind `--group_by_modality_length`
    (ola_vlm/train/llava_trainer.py:122-215) — same index streams for the same torch generator;
  * `tokenizer_image_token` (ola_vlm/mm_utils.py:336-355) — `<image>` → IMAGE_TOKEN_INDEX splice;
  * the supervised collator (ola_vlm/train/ola_vlm_train.py:881-925) — the batch schema the model
    boundary consumes;
  * `LazySupervisedDataset` / `make_supervised_data_module` (:774-878, 928-937): LLaVA json / jsonl +
    image folder → items, equal to the reference class on the same files (tests/test_lazy_dataset.py);
  * a synthetic dataset in the same item schema for runs without the json / image folders.
Prompt templating and label masking live in train/prompts.py.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch

IGNORE_INDEX = -100
IMAGE_TOKEN_INDEX = -200


# ------------------------------------------------------------------------------------------------ sampler
def split_to_even_chunks(indices: Sequence[int], lengths: Sequence[int], num_chunks: int) -> List[List[int]]:
    """`num_chunks` lists with (nearly) equal total length.  Ragged input falls back to a strided
    deal; otherwise each index goes to the currently lightest chunk that still has room."""
    n = len(indices)
    if n % num_chunks:
        return [list(indices[c::num_chunks]) for c in range(num_chunks)]
    cap = n // num_chunks
    bins: List[List[int]] = [[] for _ in range(num_chunks)]
    load = [0.0] * num_chunks
    for idx in indices:
        c = min(range(num_chunks), key=load.__getitem__)  # first minimum, like list.index(min(...))
        bins[c].append(idx)
        load[c] += lengths[idx]
        if len(bins[c]) == cap:
            load[c] = float("inf")
    return bins


def get_length_grouped_indices(lengths, batch_size, world_size, generator=None) -> List[int]:
    """Random permutation cut into mega-batches of world_size*batch_size, each sorted by length
    (longest first, stable) and dealt into `world_size` balanced chunks."""
    perm = torch.randperm(len(lengths), generator=generator).tolist()
    mega = world_size * batch_size
    out: List[int] = []
    for s in range(0, len(perm), mega):
        block = sorted(perm[s:s + mega], key=lambda i: lengths[i], reverse=True)
        for chunk in split_to_even_chunks(block, lengths, world_size):
            out.extend(chunk)
    return out


def get_modality_length_grouped_indices(lengths, batch_size, world_size, generator=None) -> List[int]:
    """Lengths > 0 are multimodal samples, < 0 language-only.  Each modality is length-grouped on
    its own (with the GLOBAL torch RNG, as the reference does), the full mega-batches of both are
    shuffled together with `generator`, and the two ragged tails form one last sorted mega-batch."""
    if any(l == 0 for l in lengths):
        raise AssertionError("Should not have zero length.")
    if all(l > 0 for l in lengths) or all(l < 0 for l in lengths):
        return get_length_grouped_indices(lengths, batch_size, world_size, generator=generator)
    mm = [(i, l) for i, l in enumerate(lengths) if l > 0]
    lang = [(i, -l) for i, l in enumerate(lengths) if l < 0]
    mega = world_size * batch_size

    def grouped(pairs):
        ids, lens = [p[0] for p in pairs], [p[1] for p in pairs]
        order = [ids[j] for j in get_length_grouped_indices(lens, batch_size, world_size, generator=None)]
        return [order[s:s + mega] for s in range(0, len(order), mega)]

    mm_blocks, lang_blocks = grouped(mm), grouped(lang)
    tail = mm_blocks[-1] + lang_blocks[-1]
    full = mm_blocks[:-1] + lang_blocks[:-1]
    shuffled = [full[j] for j in torch.randperm(len(full), generator=generator).tolist()]
    if tail:
        shuffled.append(sorted(tail))
    return [i for block in shuffled for i in block]


class LengthGroupedSampler(torch.utils.data.Sampler):
    """llava_trainer.py:170-215: yields the whole (global) index order; the trainer deals
    consecutive batches to the ranks."""

    def __init__(self, batch_size: int, world_size: int, lengths: Optional[List[int]] = None,
                 generator=None, group_by_modality: bool = False):
        if lengths is None:
            raise ValueError("Lengths must be provided.")
        self.batch_size, self.world_size, self.lengths = batch_size, world_size, lengths
        self.generator, self.group_by_modality = generator, group_by_modality

    def __len__(self):
        return len(self.lengths)

    def __iter__(self):
        fn = get_modality_length_grouped_indices if self.group_by_modality else get_length_grouped_indices
        return iter(fn(self.lengths, self.batch_size, self.world_size, generator=self.generator))


# ------------------------------------------------------------------------------------------------ tokens
def tokenizer_image_token(prompt: str, tokenizer, image_token_index: int = IMAGE_TOKEN_INDEX,
                          return_tensors: Optional[str] = None):
    """Tokenise the text around every `<image>` and put ONE image_token_index between the pieces,
    keeping a single leading BOS."""
    pieces = [tokenizer(chunk).input_ids for chunk in prompt.split("<image>")]
    has_bos = bool(pieces) and bool(pieces[0]) and pieces[0][0] == tokenizer.bos_token_id
    skip = 1 if has_bos else 0
    ids: List[int] = [pieces[0][0]] if has_bos else []
    for k, piece in enumerate(pieces):
        if k:
            ids.append(image_token_index)
        ids.extend(piece[skip:])
    if return_tensors is None:
        return ids
    if return_tensors == "pt":
        return torch.tensor(ids, dtype=torch.long)
    raise ValueError(f"Unsupported tensor type: {return_tensors}")


# ------------------------------------------------------------------------------------------------ collator
@dataclass
class DataCollatorForSupervisedDataset:
    """Right-pads ids (pad_token_id) / labels (-100), truncates to tokenizer.model_max_length,
    derives the attention mask from the pad id, stacks same-shaped images, and forwards the
    PIL images plus the per-sample int64 distillation masks."""

    tokenizer: object

    def __call__(self, instances: Sequence[Dict]) -> Dict[str, torch.Tensor]:
        pad, limit = self.tokenizer.pad_token_id, self.tokenizer.model_max_length
        T = max(int(x["input_ids"].shape[0]) for x in instances)
        ids = torch.full((len(instances), T), pad, dtype=instances[0]["input_ids"].dtype)
        lab = torch.full((len(instances), T), IGNORE_INDEX, dtype=instances[0]["labels"].dtype)
        for r, x in enumerate(instances):
            n = x["input_ids"].shape[0]
            ids[r, :n] = x["input_ids"]
            lab[r, :x["labels"].shape[0]] = x["labels"]
        ids, lab = ids[:, :limit], lab[:, :limit]
        batch = {"input_ids": ids, "labels": lab, "attention_mask": ids.ne(pad)}
        if "image" in instances[0]:
            imgs = [x["image"] for x in instances]
            same = all(i is not None and i.shape == imgs[0].shape for i in imgs)
            batch["images"] = torch.stack(imgs) if same else imgs
        if "pil_image" in instances[0]:
            batch["pil_images"] = [x["pil_image"] for x in instances]
            for key in ("seg_mask", "depth_mask", "gen_mask"):
                batch[key] = torch.tensor([x[key] for x in instances])
        return batch


# ------------------------------------------------------------------------------------------------ dataset
@dataclass
class DataArguments:
    """ola_vlm_train.py:111-118 (+ the fields train() attaches: image_processor, mm_use_im_start_end)."""
    data_path: Optional[str] = None
    lazy_preprocess: bool = False
    is_multimodal: bool = False
    image_folder: Optional[str] = None
    image_aspect_ratio: str = "square"
    image_processor: object = None
    mm_use_im_start_end: bool = False
    version: str = "llava_llama_3"   # --version: selects the chat template (conversation_lib.default_conversation)


def read_jsonl(path):
    import json

    with open(path, "r") as fh:
        return [json.loads(line) for line in fh]


def expand2square(pil_img, background_color):
    """Pad to a square on the mean colour (ola_vlm_train.py:826-838, image_aspect_ratio == 'pad')."""
    from PIL import Image

    w, h = pil_img.size
    if w == h:
        return pil_img
    side = max(w, h)
    out = Image.new(pil_img.mode, (side, side), background_color)
    out.paste(pil_img, (0, (w - h) // 2) if w > h else ((h - w) // 2, 0))
    return out


class LazySupervisedDataset(torch.utils.data.Dataset):
    """ola_vlm_train.py:774-878: JSON / JSONL conversations, images opened on access, preprocessed with
    the tower's image processor, prompts rendered + label-masked by the llama3 / phi3 template; text-only
    samples of a multimodal run get a black image and zero distillation masks."""

    def __init__(self, data_path: str, tokenizer, data_args):
        import json

        super().__init__()
        self.list_data_dict = read_jsonl(data_path) if "jsonl" in data_path else json.load(open(data_path, "r"))
        self.tokenizer = tokenizer
        self.data_args = data_args

    def __len__(self):
        return len(self.list_data_dict)

    @property
    def lengths(self):
        return [sum(len(c["value"].split()) for c in s["conversations"]) + (128 if "image" in s else 0)
                for s in self.list_data_dict]

    @property
    def modality_lengths(self):
        out = []
        for s in self.list_data_dict:
            n = sum(len(c["value"].split()) for c in s["conversations"])
            out.append(n if "image" in s else -n)
        return out

    def _crop_size(self):
        proc = self.data_args.image_processor
        size = getattr(proc, "crop_size", None) or proc.size
        return size

    def _preprocess(self, sources, has_image):
        from . import prompts as P

        version = getattr(self.data_args, "version", "llava_llama_3")
        if "phi" in version:
            return P.preprocess_phi_3(sources, self.tokenizer, has_image=has_image)
        if "llama_3" in version or "llama3" in version:
            return P.preprocess_llama_3(sources, self.tokenizer, has_image=has_image)
        raise NotImplementedError(f"chat template {version!r}: only the llama3 / phi3 templates are on the shipped path")

    def __getitem__(self, i) -> Dict[str, torch.Tensor]:
        import copy
        import os

        from PIL import Image

        from . import prompts as P

        sample = self.list_data_dict[i]
        has_image = "image" in sample
        if has_image:
            path = os.path.join(self.data_args.image_folder, sample["image"])
            proc = self.data_args.image_processor
            image = Image.open(path).convert("RGB")
            pil_image = Image.open(path).convert("RGB")
            if self.data_args.image_aspect_ratio == "pad":
                image = expand2square(image, tuple(int(x * 255) for x in proc.image_mean))
            image = proc.preprocess(image, return_tensors="pt")["pixel_values"][0]
            sources = P.preprocess_multimodal(copy.deepcopy([sample["conversations"]]),
                                              is_multimodal=self.data_args.is_multimodal,
                                              mm_use_im_start_end=self.data_args.mm_use_im_start_end)
        else:
            sources = copy.deepcopy([sample["conversations"]])
        d = self._preprocess(sources, has_image)
        out = dict(input_ids=d["input_ids"][0], labels=d["labels"][0])
        if has_image:
            out.update(image=image, pil_image=pil_image, seg_mask=1, depth_mask=1, gen_mask=1)
        elif self.data_args.is_multimodal:
            cs = self._crop_size()
            out.update(image=torch.zeros(3, cs["height"], cs["width"]),
                       pil_image=Image.new("RGB", (cs["width"], cs["height"]), color="black"),
                       seg_mask=0, depth_mask=0, gen_mask=0)
        return out


def make_supervised_data_module(tokenizer, data_args) -> Dict:
    """ola_vlm_train.py:928-937."""
    return dict(train_dataset=LazySupervisedDataset(tokenizer=tokenizer, data_path=data_args.data_path,
                                                    data_args=data_args),
                eval_dataset=None, data_collator=DataCollatorForSupervisedDataset(tokenizer=tokenizer))


class SyntheticSupervisedDataset(torch.utils.data.Dataset):
    """Items in LazySupervisedDataset.__getitem__'s schema with seeded synthetic content
    (SURVEY.md §8d): one image token at `n_sys`, labels masked up to n_sys+8, N(0,1) image."""

    def __init__(self, n, vocab, n_sys, min_text=48, max_text=192, image_size=336, distill=True,
                 text_only_every=0, seed=0):
        g = torch.Generator().manual_seed(seed)
        self.lens = torch.randint(min_text, max_text + 1, (n,), generator=g).tolist()
        self.vocab, self.n_sys, self.image_size, self.distill = vocab, n_sys, image_size, distill
        self.text_only = [bool(text_only_every) and (i % text_only_every == text_only_every - 1) for i in range(n)]
        self.seed = seed

    def __len__(self):
        return len(self.lens)

    @property
    def lengths(self):
        return [l + (0 if t else 128) for l, t in zip(self.lens, self.text_only)]

    @property
    def modality_lengths(self):  # >0 multimodal, <0 language-only (ola_vlm_train.py:801-809)
        return [-l if t else l for l, t in zip(self.lens, self.text_only)]

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed * 1_000_003 + i)
        n = self.lens[i]
        ids = torch.randint(0, self.vocab - 1, (n,), generator=g)
        labels = ids.clone()
        labels[: self.n_sys + 8] = IGNORE_INDEX
        item = {"input_ids": ids, "labels": labels}
        if not self.text_only[i]:
            ids[self.n_sys] = IMAGE_TOKEN_INDEX
            labels[self.n_sys] = IGNORE_INDEX
            item["image"] = torch.randn(3, self.image_size, self.image_size, generator=g)
        else:  # the reference feeds a zero image for text-only samples of a multimodal model (:872-875)
            item["image"] = torch.zeros(3, self.image_size, self.image_size)
        if self.distill:
            item.update(pil_image=None, seg_mask=1, depth_mask=1, gen_mask=1)
        return item
