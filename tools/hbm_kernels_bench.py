"""The HBM-bound kernels of the step at their production shapes (Llama-3-8B, B=8, T=2048 → M=16384 rows):
CUDA-event timing with the L2 flushed between iterations → achieved GB/s on the ALGORITHMIC bytes, next to
the measured HBM peak (MEASURED_PEAKS.json).  `--once` launches every kernel exactly twice (warm-up + one)
so that `ncu --set full -k regex:...` over this script stays small:

    python tools/hbm_kernels_bench.py            # prints one JSON line per kernel
    ncu --set full --clock-control none -k regex:'norm|swiglu_bwd|ce_fwd_bwd|distill|gather_rows|adamw|rope|colsum|sumsq' \
        -o gpurun_out/r02_hbm python tools/hbm_kernels_bench.py --once
"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from visper_lm_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
BF = torch.bfloat16
ONCE = "--once" in sys.argv
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
HBM = peaks.get("hbm_gbs", 6650.0)


def timeit(fn, iters=10, warmup=3):
    if ONCE:
        fn()
        flush.zero_()
        fn()
        torch.cuda.synchronize()
        return float("nan")
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


def report(name, ms, nbytes, **kw):
    gbs = nbytes / ms / 1e6 if ms == ms else None
    print(json.dumps({"kernel": name, **kw, "ms": None if ms != ms else round(ms, 4), "algorithmic_MB": round(nbytes / 1e6, 1),
                      "GBps": None if gbs is None else round(gbs, 1), "frac_of_measured_hbm": None if gbs is None else round(gbs / HBM, 3),
                      "hbm_peak_GBps": HBM}), flush=True)


def main():
    M, D, F, V = 16384, 4096, 14336, 128256
    x = torch.randn(M, D, device=dev).to(BF)
    w = (1 + 0.1 * torch.randn(D, device=dev)).to(BF)
    dy = torch.randn(M, D, device=dev).to(BF)
    dres = torch.randn(M, D, device=dev).to(BF)
    y = torch.empty_like(x)
    _, rstd = ops.rmsnorm_fwd(x, w, 1e-5, out=y)
    report("rmsnorm_fwd", timeit(lambda: ops.rmsnorm_fwd(x, w, 1e-5, out=y)), 4.0 * M * D, M=M, D=D)
    report("rmsnorm_bwd(+dres)", timeit(lambda: ops.rmsnorm_bwd(dy, x, w, rstd, dres, out=y)), 8.0 * M * D, M=M, D=D)
    ops.set_option(ops.OPT_NORM_R1, 1)
    report("rmsnorm_fwd[CTA-per-row, round 1]", timeit(lambda: ops.rmsnorm_fwd(x, w, 1e-5, out=y)), 4.0 * M * D, M=M, D=D)
    report("rmsnorm_bwd(+dres)[CTA-per-row, round 1]", timeit(lambda: ops.rmsnorm_bwd(dy, x, w, rstd, dres, out=y)),
           8.0 * M * D, M=M, D=D)
    ops.set_option(ops.OPT_NORM_R1, 0)
    # context for the numbers above: what a plain copy of the SAME size reaches (the 6.55 TB/s peak in
    # MEASURED_PEAKS.json is a 4 GiB copy; a 134 MB -> 134 MB pass lasts ~60 us, ramp and tail included)
    report("torch copy_, same size as rmsnorm_fwd", timeit(lambda: y.copy_(x)), 4.0 * M * D, M=M, D=D)
    big_a = torch.empty(1 << 30, dtype=BF, device=dev)
    big_b = torch.empty_like(big_a)
    report("torch copy_, 2 GiB -> 2 GiB", timeit(lambda: big_b.copy_(big_a), iters=5), 4.0 * big_a.numel(), elems=big_a.numel())
    del big_a, big_b
    report("colsum(dy*xhat) [rmsnorm dw]", timeit(lambda: ops.colsum(dy, x, None, rstd)), 4.0 * M * D, M=M, D=D)
    # SwiGLU backward: g|u [M,2F] + dh [M,F] in, d_gate|d_up [M,2F] out
    gu = torch.randn(M, 2 * F, device=dev).to(BF)
    dh = torch.randn(M, F, device=dev).to(BF)
    report("swiglu_bwd", timeit(lambda: ops.swiglu_bwd(gu, dh)), 10.0 * M * F, M=M, F=F)
    del gu, dh
    # RoPE in place on the packed q|k|v rows (inverse rotation of dq/dk in the backward): 40 heads of 128
    qkv = torch.randn(M, 6144, device=dev).to(BF)
    cos, sin = ops.rope_tables(4096, 128, 500000.0, dev)
    report("rope_inplace(inverse, dq|dk)", timeit(lambda: ops.rope_(qkv, 2048, cos, sin, 40, 128, inverse=True)),
           4.0 * M * 40 * 128, M=M, heads=40)
    del qkv
    # NTP cross-entropy on one chunk of label rows: 2 reads + 1 write of the bf16 logits
    R = 4096
    logits = torch.randn(R, V, device=dev).to(BF)
    labels = torch.randint(0, V, (R,), device=dev)
    row_loss = torch.empty(R, dtype=torch.float32, device=dev)
    count = ops.ce_count(labels, 1, shift=False)
    report("ce_fwd_bwd", timeit(lambda: ops.ce_fwd_bwd_(logits, labels, 0, 1, row_loss, count, 1.0, True, shift=False)),
           6.0 * R * V, rows=R, V=V)
    del logits
    # distillation loss (depth head: n = 576*1024 per sample; seg: 576*1536), B = 8 local, 8 gathered targets
    for name, n in (("depth", 576 * 1024), ("seg", 576 * 1536)):
        B = 8
        pred = torch.randn(B, n, device=dev).to(BF)
        tgt = torch.randn(B, n, device=dev).to(BF)
        tau = torch.full((1,), 2.0, device=dev)
        out4, coef, _ = ops.distill_loss_fwd(pred, tgt, 0, tau, None, 0.3)
        g0 = torch.ones(1, device=dev)
        report(f"distill_loss_fwd[{name}]", timeit(lambda: ops.distill_loss_fwd(pred, tgt, 0, tau, None, 0.3)),
               2.0 * (B + B) * n, B=B, n=n)
        report(f"distill_loss_bwd[{name}]", timeit(lambda: ops.distill_loss_bwd(pred, tgt, 0, coef, g0)),
               2.0 * (3 * B) * n, B=B, n=n)
    # splice gather: 16384 output rows of 4096 from the embedding table / image features / task tokens
    emb = torch.randn(V, D, device=dev).to(BF)
    idx = torch.randint(0, V, (M,), device=dev, dtype=torch.int32)
    out = torch.empty(M, D, dtype=BF, device=dev)
    report("gather_rows[splice]", timeit(lambda: ops.gather_rows(idx, [emb], D, out=out)), 4.0 * M * D, rows=M, D=D)
    del emb
    # AdamW on the PT-stage shard (196 M parameters at world 1): fp32 master/m/v read+write, bf16 grad in, bf16 param out
    n = 196 * 1024 * 1024
    master = torch.randn(n, device=dev)
    m = torch.zeros(n, device=dev)
    v = torch.zeros(n, device=dev)
    g = torch.randn(n, device=dev).to(BF)
    p = torch.empty(n, dtype=BF, device=dev)
    coef1 = torch.ones(1, device=dev)
    report("adamw_step", timeit(lambda: ops.adamw_step_(master, m, v, g, p, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, grad_scale=coef1)),
           28.0 * n, params=n)
    ss = torch.zeros(1, device=dev)
    report("grad_sumsq", timeit(lambda: ops.grad_sumsq(g, out=ss)), 2.0 * n, params=n)


if __name__ == "__main__":
    main()
