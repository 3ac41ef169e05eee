"""TEST INFRASTRUCTURE ONLY — golden vectors for the frozen segmentation teacher.
The reference's `self.oneformer.forward_features` (aux_heads/oneformer_head.py:42-69) returns the last
feature map of transformers' SwinBackbone (third-party, absent from /root/reference) resized to 24x24;
this script runs the transformers build installed in this image on seeded weights / pixel values, image
by image like the reference's loop (base_ola_vlm.py:382-397).  Run:
    python -m oracle.make_golden_seg_teacher
Cases: the OneFormer Swin-L geometry at the processor's 800x800 and a miniature at 120x120 whose grid
needs window padding (30→32, 15→16), shifted-window masks and an odd patch merge (15→8)."""
from __future__ import annotations

import sys
from pathlib import Path

import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import restate  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"
MINI = dict(embed_dim=32, depths=(2, 2, 2, 2), num_heads=(1, 2, 4, 8), window_size=4, patch_size=4)
CASES = [("seg_teacher_mini_120", MINI, 128, 120, 2, 991), ("seg_teacher_swinl_800", restate.SWIN_L, 384, 800, 1, 992)]
PREFIX = "oneformer.pixel_level_module.encoder."


def seg_pixels(B, size, seed):
    g = torch.Generator().manual_seed(seed)
    return 1.1 * torch.randn(B, 3, size, size, generator=g)


def seg_param(name, shape):
    r = restate.seeded_param(name, shape)
    return 2.0 * r if name.endswith("relative_position_bias_table") else r   # O(1) biases, so they matter


def main():
    from transformers import SwinBackbone, SwinConfig

    for name, cfg, init_size, size, B, seed in CASES:
        hf = SwinConfig(image_size=init_size, patch_size=cfg["patch_size"], embed_dim=cfg["embed_dim"],
                        depths=list(cfg["depths"]), num_heads=list(cfg["num_heads"]), window_size=cfg["window_size"],
                        drop_path_rate=0.3, out_features=["stage1", "stage2", "stage3", "stage4"])
        net = SwinBackbone(hf).float().eval()
        spec = {}
        with torch.no_grad():
            for n, p in net.named_parameters():
                p.copy_(seg_param(PREFIX + n, tuple(p.shape)))
                spec[PREFIX + n] = tuple(p.shape)
        px = seg_pixels(B, size, seed)
        with torch.no_grad():
            maps = [net(px[b:b + 1]).feature_maps for b in range(B)]
            tgt = torch.cat([F.interpolate(m[-1], size=(24, 24), mode="bilinear", align_corners=False) for m in maps])
        fx = {"config": dict(cfg), "size": size, "B": B, "seed": seed, "state_spec": spec,
              "targets_sub": tgt[:, ::8, ::2, ::2].clone(), "tgt_mean": float(tgt.mean()), "tgt_std": float(tgt.std()),
              "stage_shapes": [tuple(m.shape[1:]) for m in maps[0]],
              "stage1_sub": torch.cat([m[0] for m in maps])[:, ::8, ::5, ::5].clone()}
        torch.save(fx, GOLDEN / f"{name}.pt")
        sd = {n: seg_param(n, s) for n, s in spec.items()}
        with torch.no_grad():
            mine = restate.seg_teacher_targets(sd, px, cfg, PREFIX)
            st = restate.swin_stage_features(sd, px, cfg, PREFIX)
        print(name, tuple(tgt.shape), fx["stage_shapes"], "std", fx["tgt_std"], "restatement max abs diff",
              float((mine - tgt).abs().max()), float((st[0] - torch.cat([m[0] for m in maps])).abs().max()),
              (GOLDEN / f"{name}.pt").stat().st_size)


if __name__ == "__main__":
    main()
