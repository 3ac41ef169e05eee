"""initialize_vision_tokenizer / resize_token_embeddings (SURVEY.md §8b accessors) against the reference's
own method (ola_arch.py:446-488, extracted from the source and run on a plain-torch stand-in)."""
from types import SimpleNamespace

import pytest
import torch
import torch.nn as nn

from parity_utils import build_product, configs


class _Tok:
    def __init__(self, n):
        self.n = n
        self.added = []

    def add_tokens(self, toks, special_tokens=False):
        new = [t for t in toks if t not in self.added]
        self.added += new
        self.n += len(new)
        return len(new)

    def __len__(self):
        return self.n


class _RefModel:
    """What the extracted reference method needs from `self` (HF PreTrainedModel behaviour)."""

    def __init__(self, emb, head):
        self.emb, self.head = nn.Embedding.from_pretrained(emb.clone(), freeze=False), nn.Linear(head.shape[1], head.shape[0], bias=False)
        self.head.weight.data.copy_(head)

    def get_input_embeddings(self):
        return self.emb

    def get_output_embeddings(self):
        return self.head

    def resize_token_embeddings(self, n):
        for mod, attr in ((self.emb, "weight"), (self.head, "weight")):
            old = getattr(mod, attr).data
            new = torch.zeros(n, old.shape[1])
            new[:old.shape[0]] = old[:n]
            getattr(mod, attr).data = new
        self.emb.num_embeddings = n


def _args(**kw):
    base = dict(mm_use_im_patch_token=False, mm_use_im_start_end=False, tune_mm_mlp_adapter=False,
                pretrain_mm_mlp_adapter=None)
    base.update(kw)
    return SimpleNamespace(**base)


@pytest.mark.parametrize("patch,startend,tune", [(False, False, False), (True, False, True), (False, True, True),
                                                   (True, True, False)])
def test_matches_reference_method(patch, startend, tune):
    from oracle import ref_functions, ref_shim

    if not ref_shim.available():
        pytest.skip("/root/reference not mounted")
    ref_fn = ref_functions.extract_method(
        "ola_vlm/model/ola_arch.py", "OlaLlavaMetaForCausalLM", "initialize_vision_tokenizer",
        {"DEFAULT_IMAGE_PATCH_TOKEN": "<im_patch>", "DEFAULT_IM_START_TOKEN": "<im_start>", "DEFAULT_IM_END_TOKEN": "<im_end>"})
    cfg = configs.TINY_LLAMA
    model = build_product(cfg, False, None)
    V = model.config.vocab_size
    with torch.no_grad():
        model.model.embed_tokens.weight.copy_(torch.randn(V, cfg["hidden"]))
        model.lm_head.weight.copy_(torch.randn(V, cfg["hidden"]))
    ref = _RefModel(model.model.embed_tokens.weight.float(), model.lm_head.weight.float())
    args = _args(mm_use_im_patch_token=patch, mm_use_im_start_end=startend, tune_mm_mlp_adapter=tune)
    t_ref, t_mine = _Tok(V), _Tok(V)
    ref_fn(ref, args, t_ref)
    model.initialize_vision_tokenizer(args, t_mine)
    n = len(t_ref)
    assert len(t_mine) == n == V + int(patch) + 2 * int(startend)
    assert model.config.vocab_size % 8 == 0 and model.config.vocab_size >= n
    assert model.model.embed_tokens.weight.shape[0] == model.lm_head.weight.shape[0] == model.config.vocab_size
    ours_in, ours_out = model.model.embed_tokens.weight.float(), model.lm_head.weight.float()
    assert torch.equal(ours_in[:V], ref.emb.weight[:V].to(torch.bfloat16).float())          # old rows untouched
    if startend:   # the two new rows = mean of everything before them (which includes <im_patch> if added)
        lo = n - 2
        if not patch:   # with <im_patch> the mean includes that row's random init, which differs by construction
            assert torch.allclose(ours_in[lo:n], ref.emb.weight[lo:n], atol=2e-2)
            assert torch.allclose(ours_out[lo:n], ref.head.weight[lo:n], atol=2e-2)
        assert torch.allclose(ours_in[lo:n], ours_in[:lo].mean(0, keepdim=True), atol=1e-2)
    if startend or patch:
        assert model.model.embed_tokens.weight.requires_grad == ref.emb.weight.requires_grad or not tune
        if tune:
            assert model.model.embed_tokens.weight.requires_grad == ref.emb.weight.requires_grad
            assert model.lm_head.weight.requires_grad == ref.head.weight.requires_grad


def test_pretrained_adapter_embeddings_are_handed_over(tmp_path):
    cfg = configs.TINY_LLAMA
    model = build_product(cfg, False, None)
    V, D = model.config.vocab_size, cfg["hidden"]
    new_rows = torch.randn(2, D)
    path = tmp_path / "mm_projector.bin"
    torch.save({"model.embed_tokens.weight": new_rows}, path)
    tok = _Tok(V)
    model.initialize_vision_tokenizer(_args(mm_use_im_start_end=True, pretrain_mm_mlp_adapter=str(path)), tok)
    got = model.model.embed_tokens.weight[V:V + 2].float()
    assert torch.equal(got, new_rows.to(torch.bfloat16).float())
    with pytest.raises(ValueError):
        torch.save({"model.embed_tokens.weight": torch.randn(5, D + 1)}, path)
        model.initialize_vision_tokenizer(_args(mm_use_im_start_end=True, pretrain_mm_mlp_adapter=str(path)), _Tok(V))
