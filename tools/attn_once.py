"""One attention forward + backward at the decoder's shape (B=8, H=32, KVH=8, S=2048, hd=128, causal) after one
warm-up pair — the target of the per-kernel captures under profiles/ (r02_ncu_final_attn_*.csv):
    ncu --set full --import-source on --clock-control none -k regex:attn_bwd_dkdv_tc2 -s 1 -c 1 -o rep python tools/attn_once.py
    ncu -i rep.ncu-rep --page raw --csv ; ncu -i rep.ncu-rep --page source --csv"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from visper_lm_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, H, KVH, S, hd = 8, 32, 8, 2048, 128
qkv = torch.randn(B * S, (H + 2 * KVH) * hd, device=dev).to(torch.bfloat16)
q, k, v = qkv[:, :H * hd], qkv[:, H * hd:(H + KVH) * hd], qkv[:, (H + KVH) * hd:]
o, lse = ops.attn_fwd(q, k, v, B, H, KVH, S, S, hd, hd ** -0.5, True)
do = torch.randn_like(o)
dqkv = torch.empty_like(qkv)
dq, dk, dv = dqkv[:, :H * hd], dqkv[:, H * hd:(H + KVH) * hd], dqkv[:, (H + KVH) * hd:]
for _ in range(2):
    ops.attn_fwd(q, k, v, B, H, KVH, S, S, hd, hd ** -0.5, True, out=o)
    ops.attn_bwd(q, k, v, o, do, lse, dq, dk, dv, B, H, KVH, S, S, hd, hd ** -0.5, True)
torch.cuda.synchronize()
