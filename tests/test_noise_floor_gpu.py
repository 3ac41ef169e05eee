"""How close is "close"?  The reference trains in bf16 (flash-attn + cuBLAS), so it deviates from its own fp32
arithmetic; SURVEY.md §7 asks that parity tolerances be read against THAT gap.  This test measures it on the GPU at
full size — Llama-3-8B (32 layers) + CLIP-ViT-L/14-336 + the six dsg heads, B = 2, T = 663 — with three forwards on the
same bf16-rounded weights and inputs:

    R32  the unmodified reference classes (oracle/ref_shim) in fp32, sdpa attention        = ground truth
    R16  the same classes in bf16 with attn_implementation="flash_attention_2"             = the reference as it trains
    V    this repo's CUDA path (through the C ABI, bf16 storage)

and reports text loss, logits on the label rows, all 33 hidden states, the head embeddings and the reference's own
`_emb_loss(preds, ones, targets, logit_scale)` (base_ola_vlm.py:289-320; called directly because the published
forward zeroes its masks) for R16 and V against R32.  Asserted: V is no farther from R32 than twice the
reference's own bf16 path (plus small absolute floors), and the text loss is within 1e-3 relative."""
import json
import math
import os
import zlib

import pytest
import torch

from parity_utils import configs, pt_freeze, rel_err, round_batch, run_product

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _seeded_init(model):
    with torch.no_grad():
        for name, p in model.named_parameters():
            g = torch.Generator(device=p.device).manual_seed(zlib.crc32(name.encode()) & 0x7FFFFFFF)
            last = name.split(".")[-1]
            is_norm = any(k in name for k in ("norm", "layrnorm", "layer_norm"))
            if name.endswith("logit_scale"):
                p.fill_(2.0)
                continue
            r = torch.randn(p.shape, generator=g, device=p.device, dtype=torch.float32)
            if is_norm and last == "weight":
                r = 1.0 + 0.1 * r
            elif last == "bias":
                r = 0.05 * r
            elif p.dim() >= 2 and not any(k in name for k in ("special_", "embed_tokens", "position_embedding")):
                r = r / math.sqrt(p[0].numel())
            else:
                r = 0.5 * r
            p.copy_(r.to(p.dtype))


def _reference(cfg, ours_sd, dtype, attn):
    from oracle import ref_shim

    m = ref_shim.build_reference_model(cfg, "llama", True, attn_implementation=attn, device=DEV, dtype=dtype,
                                       fast_init=True)
    with torch.no_grad():
        own = dict(m.named_parameters())
        missing = [n for n in own if n not in ours_sd]
        assert not missing, missing[:5]
        for n, p in own.items():
            p.copy_(ours_sd[n].to(p.dtype))
    return m


def _ref_forward(m, batch, dtype):
    from oracle import ref_shim

    B = batch["input_ids"].shape[0]
    tg = {k: v.to(DEV, dtype) for k, v in batch["targets"].items()}
    ref_shim.install_synthetic_teachers(m, tg)
    ones = lambda: torch.ones(B, dtype=torch.long, device=DEV)
    with torch.no_grad():
        out = m(input_ids=batch["input_ids"].to(DEV), labels=batch["labels"].to(DEV),
                attention_mask=batch["attention_mask"].to(DEV), images=batch["images"].to(DEV, dtype),
                pil_images=[None] * B, depth_mask=ones(), seg_mask=ones(), gen_mask=ones())
        terms = {}
        for task, embs, scale in (("depth", [e[0][0] for e in out.depth_embs], m.depth_logit_scale),
                                  ("seg", out.seg_embs, m.seg_logit_scale), ("gen", out.image_embs, m.gen_logit_scale)):
            terms[task] = []
            for e in embs:
                l, s1, c = m._emb_loss(e, ones(), tg[task], scale)
                terms[task].append((float(l), float(s1), float(c)))
    return out, terms


def test_cuda_path_vs_the_references_own_bf16_noise_floor():
    from oracle import ref_shim
    from parity_utils import product_classes
    from visper_lm_b200.model import presets

    if not ref_shim.available():
        pytest.skip("neither /root/reference nor oracle/_ref present")
    cfg = dict(configs.LLAMA3_8B)
    ours = product_classes()[("llama", True)](presets.from_dict(cfg, distill=True), device=DEV)
    _seeded_init(ours)
    ours.config.materialize_logits = True
    pt_freeze(ours)
    batch = round_batch(configs.synthetic_batch(cfg, 2, 64, seed=4242))
    with torch.no_grad():
        v = run_product(ours, batch, True, DEV)
    torch.cuda.synchronize()
    sd = {n: p.detach() for n, p in ours.named_parameters()}
    v_out = dict(text=float(v.text_loss), hidden=[h.float() for h in v.hidden_states], logits=v.logits.float(),
                 embs=dict(depth=[e[0][0].float() for e in v.depth_embs], seg=[e.float() for e in v.seg_embs],
                           gen=[e.float() for e in v.image_embs]),
                 terms={k: [tuple(t.tolist()) for t in ts] for k, ts in v.loss_terms.items()})

    r32m = _reference(cfg, sd, torch.float32, "sdpa")
    r32, t32 = _ref_forward(r32m, batch, torch.float32)
    r32 = dict(text=float(r32.loss), hidden=[h.float() for h in r32.hidden_states], logits=r32.logits.float(),
               embs=dict(depth=[e[0][0].float() for e in r32.depth_embs], seg=[e.float() for e in r32.seg_embs],
                         gen=[e.float() for e in r32.image_embs]), terms=t32)
    del r32m
    torch.cuda.empty_cache()
    attn = "flash_attention_2"
    try:
        r16m = _reference(cfg, sd, torch.bfloat16, attn)
        r16, t16 = _ref_forward(r16m, batch, torch.bfloat16)
    except Exception as ex:  # flash_attn unusable on this box: the library's other fused path
        attn = f"sdpa ({type(ex).__name__})"
        r16m = _reference(cfg, sd, torch.bfloat16, "sdpa")
        r16, t16 = _ref_forward(r16m, batch, torch.bfloat16)
    r16 = dict(text=float(r16.loss), hidden=[h.float() for h in r16.hidden_states], logits=r16.logits.float(),
               embs=dict(depth=[e[0][0].float() for e in r16.depth_embs], seg=[e.float() for e in r16.seg_embs],
                         gen=[e.float() for e in r16.image_embs]), terms=t16)
    del r16m

    labels = batch["labels"]
    T = r32["logits"].shape[1]
    # label rows of the embedded sequence: the reference returns its spliced labels only inside the loss; the text
    # rows after the image + task tokens are the last (n_text - S - 1) positions
    rows = torch.arange(T - (labels.shape[1] - cfg["num_sys_tokens"] - 9), T - 1, device=DEV)

    def report(x):
        d = {"text_loss": x["text"], "text_loss_rel": abs(x["text"] - r32["text"]) / abs(r32["text"]),
             "hidden_rel_worst": max(rel_err(a, b) for a, b in zip(x["hidden"], r32["hidden"])),
             "hidden_rel_last": rel_err(x["hidden"][-1], r32["hidden"][-1]),
             "logits_rel": rel_err(x["logits"][:, rows], r32["logits"][:, rows]),
             "logits_max_abs": float((x["logits"][:, rows] - r32["logits"][:, rows]).abs().max())}
        for task in ("depth", "seg", "gen"):
            d[f"{task}_emb_rel_worst"] = max(rel_err(a.reshape(b.shape), b) for a, b in zip(x["embs"][task], r32["embs"][task]))
            d[f"{task}_sl1_rel_worst"] = max(abs(a[1] - b[1]) / abs(b[1]) for a, b in zip(x["terms"][task], r32["terms"][task]))
            d[f"{task}_infonce_abs_worst"] = max(abs(a[2] - b[2]) for a, b in zip(x["terms"][task], r32["terms"][task]))
            d[f"{task}_total_rel_worst"] = max(abs(a[0] - b[0]) / abs(b[0]) for a, b in zip(x["terms"][task], r32["terms"][task]))
        return d

    res = {"config": "Llama-3-8B (32 layers) + CLIP-ViT-L/14-336 + six dsg heads, B=2, T=%d, identical bf16-rounded weights" % T,
           "ground_truth": "unmodified reference classes, fp32, sdpa", "reference_bf16_attention": attn,
           "reference_bf16_vs_fp32": report(r16), "this_repo_vs_fp32": report(v_out)}
    line = json.dumps(res)
    print("[noise-floor] " + line)
    try:
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/r02_parity_noise_floor.json", "w") as fh:
            fh.write(line + "\n")
    except OSError:
        pass
    a, b = res["this_repo_vs_fp32"], res["reference_bf16_vs_fp32"]
    assert a["text_loss_rel"] <= 1e-3, a
    for k in a:
        if k.endswith("_rel") or k.endswith("_rel_worst") or k.endswith("_abs_worst") or k == "logits_max_abs":
            assert a[k] <= 2.0 * b[k] + 2e-3, f"{k}: this repo {a[k]:.3e} vs the reference's own bf16 path {b[k]:.3e}"
