"""TEST INFRASTRUCTURE ONLY — golden vectors for the frozen generation teacher.
The reference's `self.pipe.image_encoder` (base_ola_vlm.py:323-333) is transformers'
CLIPVisionModelWithProjection (third-party, absent from /root/reference); this script runs the
transformers build installed in this image on seeded weights / pixel values, image by image like the
reference's loop, and stores the image_embeds.  Run:  python -m oracle.make_golden_gen_teacher
Cases: the unCLIP ViT-H/14 geometry (stabilityai/stable-diffusion-2-1-unclip image_encoder config)
and a 3-layer miniature with the same head_dim 80."""
from __future__ import annotations

import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import restate  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"
VIT_H = dict(hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=16,
             patch_size=14, image_size=224, projection_dim=1024, hidden_act="gelu", layer_norm_eps=1e-5)
MINI = dict(hidden_size=320, intermediate_size=640, num_hidden_layers=3, num_attention_heads=4,
            patch_size=14, image_size=112, projection_dim=64, hidden_act="gelu", layer_norm_eps=1e-5)
CASES = [("gen_teacher_mini", MINI, 3, 881), ("gen_teacher_vith_224", VIT_H, 2, 882)]


def gen_pixels(B, size, seed):
    g = torch.Generator().manual_seed(seed)
    return 1.2 * torch.randn(B, 3, size, size, generator=g)


def main():
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection

    for name, cfg, B, seed in CASES:
        net = CLIPVisionModelWithProjection(CLIPVisionConfig(**cfg)).float().eval()
        spec = {}
        with torch.no_grad():
            for n, p in net.named_parameters():
                p.copy_(restate.seeded_param("image_encoder." + n, tuple(p.shape)))
                spec["image_encoder." + n] = tuple(p.shape)
        px = gen_pixels(B, cfg["image_size"], seed)
        with torch.no_grad():
            emb = torch.stack([net(px[b:b + 1]).image_embeds for b in range(B)], 0)   # [B,1,P]
        fx = {"config": cfg, "B": B, "seed": seed, "state_spec": spec, "image_embeds": emb.clone()}
        torch.save(fx, GOLDEN / f"{name}.pt")
        sd = {n: restate.seeded_param(n, s) for n, s in spec.items()}
        with torch.no_grad():
            mine = restate.gen_teacher_targets(sd, px, cfg["num_attention_heads"], cfg["hidden_act"], "image_encoder.")
        print(name, tuple(emb.shape), "std", float(emb.std()), "restatement max abs diff",
              float((mine - emb).abs().max()), (GOLDEN / f"{name}.pt").stat().st_size)


if __name__ == "__main__":
    main()
