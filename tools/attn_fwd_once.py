"""A few attention forwards at the Llama-3-8B step shape with the v2 (two-tile) kernel (for ncu)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from visper_lm_b200 import ops
dev = torch.device("cuda:0")
B, H, KVH, S, hd = 8, 32, 8, 2048, 128
qkv = torch.randn(B * S, (H + 2 * KVH) * hd, device=dev).to(torch.bfloat16)
q, k, v = qkv[:, :H * hd], qkv[:, H * hd:(H + KVH) * hd], qkv[:, (H + KVH) * hd:]
ops.set_option(ops.OPT_ATTN_FWD_V2, int(sys.argv[1]) if len(sys.argv) > 1 else 1)
for _ in range(4):
    o, lse = ops.attn_fwd(q, k, v, B, H, KVH, S, S, hd, hd ** -0.5, True)
torch.cuda.synchronize()
