"""Timeline of CTA 0 of the attention dK/dV kernel (clock64 stamps, vpb_set_trace_buffer)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from visper_lm_b200 import lib, ops
dev = torch.device("cuda:0")
B, H, KVH, S, hd = 8, 32, 8, 2048, 128
qkv = torch.randn(B * S, (H + 2 * KVH) * hd, device=dev).to(torch.bfloat16)
q, k, v = qkv[:, :H * hd], qkv[:, H * hd:(H + KVH) * hd], qkv[:, (H + KVH) * hd:]
o, lse = ops.attn_fwd(q, k, v, B, H, KVH, S, S, hd, hd ** -0.5, True)
do = torch.randn_like(o)
dqkv = torch.empty_like(qkv)
dq, dk, dv = dqkv[:, :H * hd], dqkv[:, H * hd:(H + KVH) * hd], dqkv[:, (H + KVH) * hd:]
for _ in range(2):
    ops.attn_bwd(q, k, v, o, do, lse, dq, dk, dv, B, H, KVH, S, S, hd, hd ** -0.5, True)
buf = torch.zeros(16 * 512, dtype=torch.int64, device=dev)
lib.load().vpb_set_trace_buffer(buf.data_ptr())
ops.attn_bwd(q, k, v, o, do, lse, dq, dk, dv, B, H, KVH, S, S, hd, hd ** -0.5, True)
torch.cuda.synchronize()
lib.load().vpb_set_trace_buffer(0)
t = buf.cpu().view(16, 512)
names = ["load_issue", "sd_issue", "dvk_issue", "sm_ready", "sm_sd_seen", "sm_regs", "sm_stored", "sm_arrived"]
n = int((t[1] > 0).sum())
t0 = int(t[1, 0])
print("iterations traced", n)
print("it   " + " ".join(f"{x:>10}" for x in names))
for it in list(range(0, 12)) + list(range(60, 72)):
    print(f"{it:3d}  " + " ".join(f"{int(t[s, it]) - t0:10d}" for s in range(8)))
import statistics
def d(a, b, lo=8, hi=None):
    hi = hi or n - 2
    return statistics.mean(int(t[b, i]) - int(t[a, i]) for i in range(lo, hi))
print("mean sd_issue(it+1)-sd_issue(it):", statistics.mean(int(t[1, i + 1]) - int(t[1, i]) for i in range(8, n - 2)))
print("mean sd_issue -> sm_sd_seen:", d(1, 4))
print("mean sm_sd_seen -> sm_regs:", d(4, 5))
print("mean sm_regs -> sm_stored:", d(5, 6))
print("mean sm_stored -> sm_arrived:", d(6, 7))
print("mean sm_arrived -> dvk_issue:", d(7, 2))
print("mean dvk_issue(it) -> sd_issue(it+2):", statistics.mean(int(t[1, i + 2]) - int(t[2, i]) for i in range(8, n - 3)))
print("mean load_issue -> sd_issue:", d(0, 1))
print("mean sm_ready -> sm_sd_seen (idle wait):", d(3, 4))
