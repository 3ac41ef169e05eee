"""Host-side input pipeline: sampler, <image> token splice and collator against
the reference's own functions (where /root/reference is mounted) and committed golden index lists."""
import json
from pathlib import Path

import pytest
import torch

from parity_utils import ROOT  # noqa: F401  (sys.path)
from oracle import ref_functions, ref_shim
from visper_lm_b200.train import data as D

GOLDEN = Path(__file__).parent / "golden" / "sampler_indices.json"


def _lengths(n, seed, mixed):
    g = torch.Generator().manual_seed(seed)
    l = torch.randint(5, 400, (n,), generator=g).tolist()
    if mixed:
        l = [-x if i % 3 == 2 else x for i, x in enumerate(l)]
    return l


CASES = [(64, 4, 2, False), (100, 4, 4, False), (96, 8, 2, True), (131, 4, 8, True), (7, 2, 2, True)]


def _run(mod, n, bs, ws, mixed):
    lengths = _lengths(n, 11 * n + bs, mixed)
    torch.manual_seed(1234)  # the modality path draws from the global RNG
    g = torch.Generator().manual_seed(77)
    fn = mod["get_modality_length_grouped_indices"] if mixed else mod["get_length_grouped_indices"]
    return fn(lengths, bs, ws, generator=g)


MINE = {"split_to_even_chunks": D.split_to_even_chunks, "get_length_grouped_indices": D.get_length_grouped_indices,
        "get_modality_length_grouped_indices": D.get_modality_length_grouped_indices}


@pytest.mark.parametrize("n,bs,ws,mixed", CASES)
def test_sampler_matches_golden(n, bs, ws, mixed):
    want = json.loads(GOLDEN.read_text())[f"{n}-{bs}-{ws}-{int(mixed)}"]
    got = _run(MINE, n, bs, ws, mixed)
    assert got == want
    assert sorted(got) == list(range(n))


@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not mounted")
@pytest.mark.parametrize("n,bs,ws,mixed", CASES)
def test_sampler_matches_reference(n, bs, ws, mixed):
    ref = ref_functions.extract("ola_vlm/train/llava_trainer.py", list(MINE))
    assert _run(MINE, n, bs, ws, mixed) == _run(ref, n, bs, ws, mixed)
    lengths = _lengths(n, 5, False)
    idx = list(range(n))
    assert D.split_to_even_chunks(idx, lengths, ws) == ref["split_to_even_chunks"](idx, lengths, ws)


class ToyTokenizer:
    """Whitespace tokenizer with a BOS, enough for the `<image>` splice and the collator."""
    bos_token_id, pad_token_id, model_max_length = 1, 0, 24

    def __call__(self, text):
        ids = [self.bos_token_id] + [2 + (sum(map(ord, w)) % 97) for w in text.split()]
        return type("Enc", (), {"input_ids": ids})


PROMPTS = ["<image>\ndescribe the picture", "look at <image> and then <image> again", "no image here", "",
           "<image>"]


@pytest.mark.parametrize("prompt", PROMPTS)
def test_tokenizer_image_token(prompt):
    tok = ToyTokenizer()
    ids = D.tokenizer_image_token(prompt, tok)
    assert ids.count(D.IMAGE_TOKEN_INDEX) == prompt.count("<image>")
    assert ids.count(tok.bos_token_id) == 1
    assert torch.equal(D.tokenizer_image_token(prompt, tok, return_tensors="pt"), torch.tensor(ids, dtype=torch.long))
    if ref_shim.available():
        ref = ref_functions.extract("ola_vlm/mm_utils.py", ["tokenizer_image_token"])["tokenizer_image_token"]
        assert ids == ref(prompt, tok)


def _instances(distill):
    ds = D.SyntheticSupervisedDataset(6, vocab=300, n_sys=13, min_text=20, max_text=40, image_size=28,
                                      distill=distill, text_only_every=3, seed=3)
    return [ds[i] for i in range(6)], ds


@pytest.mark.parametrize("distill", [True, False])
def test_collator_schema_and_reference(distill):
    items, ds = _instances(distill)
    tok = ToyTokenizer()
    batch = D.DataCollatorForSupervisedDataset(tok)(items)
    B, T = batch["input_ids"].shape
    assert B == 6 and T == min(max(ds.lens), tok.model_max_length)
    assert batch["labels"].shape == (B, T) and batch["attention_mask"].dtype == torch.bool
    assert torch.equal(batch["attention_mask"], batch["input_ids"].ne(tok.pad_token_id))
    assert batch["images"].shape == (6, 3, 28, 28)
    if distill:
        assert batch["seg_mask"].dtype == torch.int64 and batch["depth_mask"].tolist() == [1] * 6
        assert batch["pil_images"] == [None] * 6
    if ref_shim.available():
        ref = ref_functions.extract("ola_vlm/train/ola_vlm_train.py", ["DataCollatorForSupervisedDataset"])
        want = ref["DataCollatorForSupervisedDataset"](tok)(items)
        assert set(want) == set(batch)
        for k, v in want.items():
            assert torch.equal(v, batch[k]) if isinstance(v, torch.Tensor) else v == batch[k], k


def test_modality_lengths_sign_convention():
    _, ds = _instances(True)
    assert [l < 0 for l in ds.modality_lengths] == ds.text_only
    order = list(D.LengthGroupedSampler(2, 2, lengths=ds.modality_lengths, group_by_modality=True,
                                        generator=torch.Generator().manual_seed(0)))
    assert sorted(order) == list(range(6))


# ------------------------------------------------------------------------------------------------ prompts
import copy  # noqa: E402
import re  # noqa: E402
import types  # noqa: E402

from visper_lm_b200.train import prompts as P  # noqa: E402


class MarkerTokenizer:
    """Context-free toy tokenizer: chat markers (<|...|>), words and punctuation are single tokens,
    whitespace is dropped — so tokenising a conversation piecewise or whole gives the same tokens."""
    bos_token_id, pad_token_id, model_max_length = 1, 0, 512
    _re = re.compile(r"<\|[^|]+\|>|<image>|\w+|[^\w\s]")

    def _ids(self, text):
        return [self.bos_token_id] + [2 + (sum(map(ord, t)) * 31 + len(t)) % 9973 for t in self._re.findall(text)]

    def __call__(self, text, return_tensors=None, padding=None, max_length=None, truncation=None):
        if isinstance(text, str):
            return types.SimpleNamespace(input_ids=self._ids(text))
        rows = [self._ids(t)[:max_length] for t in text]
        n = max(map(len, rows))
        return types.SimpleNamespace(input_ids=torch.tensor([r + [self.pad_token_id] * (n - len(r)) for r in rows]))


def _sources(multi_round):
    one = [{"from": "human", "value": "What is in the <image> picture ?"},
           {"from": "gpt", "value": "A cat sitting on a mat."}]
    two = one + [{"from": "human", "value": "Which colour ?"}, {"from": "gpt", "value": "It is black , with white paws."}]
    lead = [{"from": "gpt", "value": "ignored leading turn"}] + one
    return [copy.deepcopy(two if multi_round else one), copy.deepcopy(lead)]


@pytest.mark.parametrize("version", ["llama3", "phi3"])
@pytest.mark.parametrize("multi_round", [False, True])
@pytest.mark.parametrize("has_image", [True, False])
def test_preprocess_label_masking(version, multi_round, has_image):
    tok = MarkerTokenizer()
    mine_fn = P.preprocess_llama_3 if version == "llama3" else P.preprocess_phi_3
    src = P.preprocess_multimodal(_sources(multi_round))
    assert src[0][0]["value"].startswith("<image>\n")
    if multi_round or not has_image:
        src = [src[0]] if has_image is False else src   # stacked tensors need equal lengths with images
    if has_image and len(src) == 2 and len(tok(P.render([src[0]], P.conv_templates["llava_" + ("llama_3" if version == "llama3" else "phi_3")])[0]).input_ids) != \
            len(tok(P.render([src[1]], P.conv_templates["llava_" + ("llama_3" if version == "llama3" else "phi_3")])[0]).input_ids):
        src = [src[0]]
    out = mine_fn(copy.deepcopy(src), tok, has_image=has_image)
    ids, lab = out["input_ids"], out["labels"]
    assert ids.shape == lab.shape
    if has_image:
        assert (ids == D.IMAGE_TOKEN_INDEX).sum().item() == len(src)
    scored = lab != D.IGNORE_INDEX
    assert torch.equal(lab[scored], ids[scored])
    if version == "llama3" or not multi_round:
        assert out["mismatches"] == 0 and scored.any(), "assistant answers must be scored"
        assert not scored[:, 0].any()  # BOS never scored
    if ref_shim.available():
        ref = ref_functions.preprocess_fns(version)
        ref_src = ref["preprocess_multimodal"](copy.deepcopy(_sources(multi_round))[: len(src)],
                                               types.SimpleNamespace(is_multimodal=True, mm_use_im_start_end=False))
        assert ref_src == src
        want = ref["preprocess_llama_3" if version == "llama3" else "preprocess_phi_3"](copy.deepcopy(src), tok,
                                                                                       has_image=has_image)
        assert torch.equal(want["input_ids"], ids)
        assert torch.equal(want["labels"], lab)


def test_templates_equal_reference():
    if not ref_shim.available():
        pytest.skip("/root/reference not mounted")
    for version, mine in (("llama3", P.LLAMA3), ("phi3", P.PHI3)):
        lib = ref_functions.conversation_lib(version)
        ref = lib.default_conversation
        assert (ref.system, tuple(ref.roles), ref.sep, ref.version) == (mine.system, mine.roles, mine.sep, mine.version)
        c, m = ref.copy(), mine.copy()
        for role_i, text in ((0, "hello <image>"), (1, "hi"), (0, "and ?"), (1, None)):
            c.append_message(ref.roles[role_i], text)
            m.append_message(mine.roles[role_i], text)
        assert c.get_prompt() == m.get_prompt()
