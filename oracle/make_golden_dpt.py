"""TEST INFRASTRUCTURE ONLY — golden vector for the frozen DPT decoder: runs the
UNMODIFIED reference DAv2_Head (aux_heads/da_v2_head.py:296-321) on seeded weights / inputs and
stores a sub-sampled depth map.  Run here, where /root/reference exists:
    python -m oracle.make_golden_dpt
Weights: restate.seeded_param("da_v2_head." + name); inputs: four N(0,1)·0.5 feature levels from
torch.Generator().manual_seed(4321)."""
from __future__ import annotations

import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import ref_shim, restate  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"


def dpt_inputs(B=2, seed=4321):
    g = torch.Generator().manual_seed(seed)
    return [0.5 * torch.randn(B, 576, 1024, generator=g) for _ in range(4)]


def main():
    R = ref_shim.load()
    head = R.da_head.DAv2_Head().float().eval()
    spec = {}
    with torch.no_grad():
        for n, p in head.named_parameters():
            p.copy_(restate.seeded_param("da_v2_head." + n, tuple(p.shape)))
            spec["da_v2_head." + n] = tuple(p.shape)
    feats = dpt_inputs()
    with torch.no_grad():
        depth = head([(f, None) for f in feats])
    norm = restate.depth_pred_normalized(depth)
    fx = {"state_spec": spec, "B": 2, "seed": 4321, "depth_sub": depth[:, ::7, ::7].clone(),
          "depth_norm_sub": norm[:, ::7, ::7].clone(), "depth_mean": float(depth.mean()),
          "depth_max": float(depth.max())}
    torch.save(fx, GOLDEN / "dpt_head.pt")
    print("dpt_head", tuple(depth.shape), fx["depth_mean"], fx["depth_max"],
          (GOLDEN / "dpt_head.pt").stat().st_size)
    # the restatement against the live reference, full resolution
    sd = {n: restate.seeded_param(n, s) for n, s in spec.items()}
    mine = restate.dav2_head(sd, feats)
    print("restatement max abs diff", float((mine - depth).abs().max()))


if __name__ == "__main__":
    main()
