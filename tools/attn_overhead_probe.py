"""Separates the per-CTA fixed cost F from the per-key-tile cost t of the tcgen05 attention forward: non-causal
runs at several sequence lengths (n = S/128 key tiles per CTA), cycles per CTA = time x SMs / CTAs ~ F + n*t."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from visper_lm_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=8, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


for variant in (1, 0):
    ops.set_option(ops.OPT_ATTN_FWD_NS2, variant)
    pts = []
    for B, S in ((16, 256), (8, 512), (8, 1024), (8, 2048), (4, 4096), (2, 8192)):
        H, KVH, hd = 32, 8, 128
        qkv = torch.randn(B * S, (H + 2 * KVH) * hd, device=dev).to(torch.bfloat16)
        q, k, v = qkv[:, :H * hd], qkv[:, H * hd:(H + KVH) * hd], qkv[:, (H + KVH) * hd:]
        o, _ = ops.attn_fwd(q, k, v, B, H, KVH, S, S, hd, hd ** -0.5, False)
        ms = timeit(lambda: ops.attn_fwd(q, k, v, B, H, KVH, S, S, hd, hd ** -0.5, False, out=o))
        nct = B * H * (S // 128)
        n = S // 128
        us_per_cta = ms * 1e3 * 148 / nct
        pts.append((n, us_per_cta))
        print(json.dumps({"variant": variant, "B": B, "S": S, "ctas": nct, "tiles_per_cta": n, "ms": round(ms, 4),
                          "us_per_cta_slot": round(us_per_cta, 3), "us_per_tile": round(us_per_cta / n, 3),
                          "tflops": round(4.0 * B * H * S * S * hd / ms / 1e9, 1)}), flush=True)
    # least-squares fit of us_per_cta = F + n t
    import numpy as np
    A = np.array([[1.0, n] for n, _ in pts]); y = np.array([u for _, u in pts])
    (F, t), *_ = np.linalg.lstsq(A, y, rcond=None)
    print(json.dumps({"variant": variant, "fit_us_fixed_per_cta": round(float(F), 3), "fit_us_per_tile": round(float(t), 3)}), flush=True)
