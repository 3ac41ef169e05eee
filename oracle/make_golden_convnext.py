"""TEST INFRASTRUCTURE ONLY — golden vectors for the CLIP-ConvNeXt tower.
The reference's CLIPConvNextVisionTower._forward (multimodal_encoder/clip_convnext_encoder.py:150-174) runs
the stem and the four stages of timm's ConvNeXt (timm==1.0.8 / open_clip: third-party, absent from
/root/reference AND from this image).  The closest executable third-party build here is transformers'
ConvNextModel — the same published architecture under other parameter names, with LayerNorm eps fixed at
1e-6.  This script runs it on seeded weights (HF names mapped 1:1 onto timm's) and stores sub-sampled stage
outputs; oracle/restate.convnext_stage_features must reproduce them.  Run:
    python -m oracle.make_golden_convnext
Cases: a miniature (dims 64..256, 96 px → 3x3 grid, so every depthwise halo is clipped) stored under
tests/golden/, and the ConvNeXt-XXL geometry (depths 3-4-30-3, dims 384..3072) at 128 px checked here only
(846 M parameters; the printed difference is recorded in DESIGN.md)."""
from __future__ import annotations

import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import restate  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"
MINI = dict(depths=(2, 2, 3, 2), dims=(64, 128, 192, 256), eps=1e-6, image_size=96)
PREFIX = "model.vision_tower.vision_tower."


def hf_to_timm(name):
    """transformers ConvNextModel parameter name → timm ConvNeXt parameter name (None = not in the trunk)."""
    if name.startswith("layernorm."):
        return None  # pooler norm: timm's head.norm, which the reference never runs (global_pool='')
    name = name.replace("embeddings.patch_embeddings.", "stem.0.").replace("embeddings.layernorm.", "stem.1.")
    name = name.replace("encoder.stages.", "stages.").replace(".downsampling_layer.", ".downsample.")
    name = name.replace(".layers.", ".blocks.").replace(".dwconv.", ".conv_dw.").replace(".layernorm.", ".norm.")
    name = name.replace(".pwconv1.", ".mlp.fc1.").replace(".pwconv2.", ".mlp.fc2.")
    return name.replace(".layer_scale_parameter", ".gamma")


def pixels(B, size, seed):
    return 1.1 * torch.randn(B, 3, size, size, generator=torch.Generator().manual_seed(seed))


def hf_model(cfg, sd):
    from transformers import ConvNextConfig, ConvNextModel

    net = ConvNextModel(ConvNextConfig(num_channels=3, patch_size=4, num_stages=4, hidden_sizes=list(cfg["dims"]),
                                       depths=list(cfg["depths"]), layer_scale_init_value=1e-6,
                                       drop_path_rate=0.1)).float().eval()
    seen = set()
    with torch.no_grad():
        for n, p in net.named_parameters():
            t = hf_to_timm(n)
            if t is None:
                continue
            p.copy_(sd[PREFIX + t].view(p.shape))
            seen.add(PREFIX + t)
    assert seen == set(sd), sorted(set(sd) ^ seen)[:5]
    return net


def main():
    for name, cfg, size, B, seed, store in (("convnext_mini_96", MINI, 96, 2, 771, True),
                                            ("convnext_xxl_128", dict(restate.CONVNEXT_XXL, eps=1e-6), 128, 1, 772, False)):
        sd = restate.convnext_seeded_state(cfg, PREFIX)
        net = hf_model(cfg, sd)
        px = pixels(B, size, seed)
        with torch.no_grad():
            hs = net(px, output_hidden_states=True).hidden_states[1:]   # [0] is the stem output
            mine = restate.convnext_stage_features(sd, px, cfg, PREFIX)
        diffs = [float((a - b).abs().max()) for a, b in zip(mine, hs)]
        print(name, [tuple(h.shape) for h in hs], "std", [round(float(h.std()), 3) for h in hs],
              "restatement max abs diff per stage", diffs)
        if store:
            fx = {"config": dict(cfg), "size": size, "B": B, "seed": seed,
                  "stages_sub": [h[:, ::4, ::2, ::2].clone() for h in hs[:3]] + [hs[3].clone()],
                  "stage_std": [float(h.std()) for h in hs]}
            torch.save(fx, GOLDEN / f"{name}.pt")
            print("  wrote", GOLDEN / f"{name}.pt", (GOLDEN / f"{name}.pt").stat().st_size, "bytes")


if __name__ == "__main__":
    main()
