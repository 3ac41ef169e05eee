"""TEST INFRASTRUCTURE ONLY — golden vectors for the frozen depth teacher: runs the
UNMODIFIED reference DepthAnythingV2.infer_image(img, is_dsg=True) (aux_heads/depth_anything_v2/
dpt.py:183-221 → dinov2.py get_intermediate_layers) image by image, exactly as _get_dav2_feats does
(base_ola_vlm.py:348-366), on seeded weights and seeded uint8 images, and stores sub-sampled targets.
Run here, where /root/reference exists (needs the real cv2, which image2tensor calls):
    python -m oracle.make_golden_dinov2
Weights: teacher_param(name) below; images: torch.randint(0, 256) from manual_seed(seed)."""
from __future__ import annotations

import importlib
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import restate  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"
CASES = [("dinov2_vits_224", "vits", 224, 2, 777, False), ("dav2_teacher_vitl_336", "vitl", 336, 2, 778, True)]


def teacher_param(name, shape):
    """seeded_param with O(1) positional / class embeddings and LayerScale gains (their real scale)."""
    r = restate.seeded_param(name, shape)
    if name.endswith("pos_embed") or name.endswith("cls_token"):
        fan = 1
        for s in tuple(shape)[1:]:
            fan *= s
        return r * (fan ** 0.5) * 0.5
    return r


def teacher_images(B, size, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (B, size, size, 3), generator=g, dtype=torch.uint8)


def main():
    import cv2  # noqa: F401  — before the shim, which would otherwise stub it
    from oracle import ref_shim

    R = ref_shim.load()
    dpt = importlib.import_module("ola_vlm.model.aux_heads.depth_anything_v2.dpt")
    for name, enc, size, B, seed, with_head in CASES:
        net = dpt.DepthAnythingV2(encoder=enc, features=256, out_channels=[256, 512, 1024, 1024]).float().eval()
        spec = {}
        with torch.no_grad():
            for n, p in net.pretrained.named_parameters():
                p.copy_(teacher_param("dav2_backbone.pretrained." + n, tuple(p.shape)))
                spec["dav2_backbone.pretrained." + n] = tuple(p.shape)
        head = hspec = None
        if with_head:
            head = R.da_head.DAv2_Head().float().eval()
            hspec = {}
            with torch.no_grad():
                for n, p in head.named_parameters():
                    p.copy_(restate.seeded_param("da_v2_head." + n, tuple(p.shape)))
                    hspec["da_v2_head." + n] = tuple(p.shape)
        raw = teacher_images(B, size, seed)
        fts, gts, cls = [], [], []
        for b in range(B):  # the reference's per-image loop
            with torch.no_grad():
                feat = net.infer_image(raw[b].numpy(), input_size=size, is_dsg=True)
                ft = (feat[0][0] + feat[1][0] + feat[2][0] + feat[3][0]) / 4
                fts.append(ft)
                cls.append(torch.stack([f[1] for f in feat], 1))
                if head is not None:
                    d = head([(ft, None)] * 4)
                    gts.append(restate.depth_pred_normalized(d))
        ft = torch.cat(fts)
        fx = {"encoder": enc, "size": size, "B": B, "seed": seed, "state_spec": spec, "head_spec": hspec,
              "ft_sub": ft[:, ::7, ::16].clone(), "ft_mean": float(ft.mean()), "ft_std": float(ft.std()),
              "cls_sub": torch.cat(cls)[:, :, ::16].clone()}
        if gts:
            fx["depth_gts_sub"] = torch.cat(gts)[:, ::7, ::7].clone()
        torch.save(fx, GOLDEN / f"{name}.pt")
        sd = {n: teacher_param(n, s) for n, s in spec.items()}
        hsd = None if hspec is None else {n: restate.seeded_param(n, s) for n, s in hspec.items()}
        with torch.no_grad():
            mine_ft, mine_gt = restate.dav2_depth_teacher(sd, hsd, raw, enc, prefix="dav2_backbone.pretrained.")
        print(name, tuple(ft.shape), "ft mean/std", fx["ft_mean"], fx["ft_std"], "restatement max abs diff",
              float((mine_ft - ft).abs().max()),
              "" if mine_gt is None else float((mine_gt - torch.cat(gts)).abs().max()),
              (GOLDEN / f"{name}.pt").stat().st_size)


if __name__ == "__main__":
    main()
