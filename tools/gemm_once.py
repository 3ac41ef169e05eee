"""A few launches of the dominant decoder GEMM shapes (for ncu --set full captures: dram bytes,
tensor pipe activity of the CTA-pair kernel)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from visper_lm_b200 import ops
dev = torch.device("cuda:0")
M = 16384
shapes = [("down dgrad", 14336, 4096, 1), ("gate_up fwd", 28672, 4096, 0), ("qkv fwd", 6144, 4096, 0), ("down fwd", 4096, 14336, 0)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, N, K, bl in shapes:
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    b = torch.randn((N, K) if bl == 0 else (K, N), device=dev).to(torch.bfloat16)
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    for _ in range(2):
        flush.zero_()
        ops.gemm(a, b, b_layout=bl, out=out)
torch.cuda.synchronize()
