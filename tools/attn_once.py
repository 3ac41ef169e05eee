"""One attention forward + backward at the decoder's shape (B=8, H=32, KVH=8, S=2048, hd=128, causal) after one
warm-up pair — the target of the per-kernel `ncu --set full --import-source on -k regex:<kernel> -s 1 -c 1` captures
under profiles/ (tools/gpu_validation.sh shows the command)."""
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
