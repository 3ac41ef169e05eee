"""Micro-benchmarks of the hot kernels (CUDA events, L2 flushed between iterations).
Prints one JSON line per case; used to fill profiles/ and DESIGN.md."""
import json
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from visper_lm_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
BF = torch.bfloat16
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10, warmup=3):
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


def gemm_case(M, N, K, al, bl, name):
    a = torch.randn((M, K) if al == 0 else (K, M), device=dev).to(BF)
    b = torch.randn((N, K) if bl == 0 else (K, N), device=dev).to(BF)
    out = torch.empty(M, N, dtype=BF, device=dev)
    ms = timeit(lambda: ops.gemm(a, b, a_layout=al, b_layout=bl, out=out))
    at = a if al == 0 else a.t()
    bt = b.t() if bl == 0 else b
    ms_ref = timeit(lambda: torch.matmul(at, bt, out=out))
    fl = 2.0 * M * N * K
    print(json.dumps({"kernel": "gemm", "name": name, "M": M, "N": N, "K": K, "a": al, "b": bl,
                      "ms": round(ms, 4), "tflops": round(fl / ms / 1e9, 1),
                      "cublas_ms": round(ms_ref, 4), "cublas_tflops": round(fl / ms_ref / 1e9, 1)}), flush=True)


def attn_case(B, H, KVH, S, hd, causal):
    qkv = torch.randn(B * S, (H + 2 * KVH) * hd, device=dev).to(BF)
    q, k, v = qkv[:, :H * hd], qkv[:, H * hd:(H + KVH) * hd], qkv[:, (H + KVH) * hd:]
    scale = hd ** -0.5
    o, lse = ops.attn_fwd(q, k, v, B, H, KVH, S, S, hd, scale, causal)
    ms_f = timeit(lambda: ops.attn_fwd(q, k, v, B, H, KVH, S, S, hd, scale, causal, out=o))
    do = torch.randn_like(o)
    dqkv = torch.empty_like(qkv)
    dq, dk, dv = dqkv[:, :H * hd], dqkv[:, H * hd:(H + KVH) * hd], dqkv[:, (H + KVH) * hd:]
    ms_b = timeit(lambda: ops.attn_bwd(q, k, v, o, do, lse, dq, dk, dv, B, H, KVH, S, S, hd, scale, causal))
    fl = 4.0 * B * H * S * S * hd * (0.5 if causal else 1.0)
    print(json.dumps({"kernel": "attention", "B": B, "H": H, "KVH": KVH, "S": S, "hd": hd, "causal": causal,
                      "fwd_ms": round(ms_f, 4), "fwd_tflops": round(fl / ms_f / 1e9, 1),
                      "bwd_ms": round(ms_b, 4), "bwd_tflops": round(2.5 * fl / ms_b / 1e9, 1)}), flush=True)


def swiglu_case(M=16384, D=4096, F=14336):
    x = (torch.randn(M, D, device=dev) * 0.5).to(BF)
    wgu = (torch.randn(2 * F, D, device=dev) * 0.02).to(BF)
    wd = (torch.randn(D, F, device=dev) * 0.02).to(BF)
    dy = torch.randn(M, D, device=dev).to(BF)
    gu = ops.gemm(x, wgu)
    _, gut = ops.gemm_swiglu_fwd(x, wgu, tiled=True)
    res = {"kernel": "swiglu_mlp", "M": M, "D": D, "F": F}
    res["fwd_gemm_ms"] = round(timeit(lambda: ops.gemm(x, wgu, out=gu)), 4)
    res["fwd_swiglu_ms"] = round(timeit(lambda: ops.swiglu_fwd(gu)), 4)
    res["fwd_fused_rowmajor_ms"] = round(timeit(lambda: ops.gemm_swiglu_fwd(x, wgu)), 4)
    res["fwd_fused_tiled_ms"] = round(timeit(lambda: ops.gemm_swiglu_fwd(x, wgu, tiled=True)), 4)
    res["fwd_fused_nogu_ms"] = round(timeit(lambda: ops.gemm_swiglu_fwd(x, wgu, want_gu=False)), 4)
    dh = ops.gemm(dy, wd, b_layout=1)
    res["bwd_gemm_ms"] = round(timeit(lambda: ops.gemm(dy, wd, b_layout=1, out=dh)), 4)
    res["bwd_swiglu_ms"] = round(timeit(lambda: ops.swiglu_bwd(gu, dh)), 4)
    res["bwd_fused_rowmajor_ms"] = round(timeit(lambda: ops.gemm_swiglu_bwd(dy, wd, gu, b_layout=1)), 4)
    res["bwd_fused_tiled_ms"] = round(timeit(lambda: ops.gemm_swiglu_bwd(dy, wd, gut, b_layout=1, tiled=True, F=F)), 4)
    print(json.dumps(res), flush=True)


def norm_case(M, D):
    x = torch.randn(M, D, device=dev).to(BF)
    w = torch.ones(D, device=dev, dtype=BF)
    y = torch.empty_like(x)
    ms = timeit(lambda: ops.rmsnorm_fwd(x, w, 1e-5, out=y))
    by = 2.0 * M * D * 2
    print(json.dumps({"kernel": "rmsnorm_fwd", "M": M, "D": D, "ms": round(ms, 4),
                      "GBps": round(by / ms / 1e6, 1)}), flush=True)


def attn_profile(B=8, H=32, KVH=8, S=2048, hd=128):
    """Per-kernel device times of one attention fwd+bwd via torch.profiler (CUPTI)."""
    from torch.profiler import ProfilerActivity, profile
    qkv = torch.randn(B * S, (H + 2 * KVH) * hd, device=dev).to(BF)
    q, k, v = qkv[:, :H * hd], qkv[:, H * hd:(H + KVH) * hd], qkv[:, (H + KVH) * hd:]
    scale = hd ** -0.5
    o, lse = ops.attn_fwd(q, k, v, B, H, KVH, S, S, hd, scale, True)
    do = torch.randn_like(o)
    dqkv = torch.empty_like(qkv)
    dq, dk, dv = dqkv[:, :H * hd], dqkv[:, H * hd:(H + KVH) * hd], dqkv[:, (H + KVH) * hd:]
    for _ in range(2):
        ops.attn_bwd(q, k, v, o, do, lse, dq, dk, dv, B, H, KVH, S, S, hd, scale, True)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            ops.attn_fwd(q, k, v, B, H, KVH, S, S, hd, scale, True, out=o)
            ops.attn_bwd(q, k, v, o, do, lse, dq, dk, dv, B, H, KVH, S, S, hd, scale, True)
        torch.cuda.synchronize()
    for e in prof.key_averages():
        if e.device_time_total > 0:
            print(json.dumps({"kernel": e.key[:60], "calls": e.count,
                              "avg_us": round(e.device_time_total / e.count, 1)}), flush=True)


def rope_bwd_case(B=8, H=32, KVH=8, S=2048, hd=128, rounds=4):
    """attention backward + inverse-RoPE kernel vs the backward with the rotation in its epilogues
    (alternating, medians over `rounds`, same box)."""
    qkv = torch.randn(B * S, (H + 2 * KVH) * hd, device=dev).to(BF)
    qw, kw = H * hd, KVH * hd
    q, k, v = qkv[:, :qw], qkv[:, qw:qw + kw], qkv[:, qw + kw:]
    scale = hd ** -0.5
    cos, sin = ops.rope_tables(S, hd, 500000.0, dev)
    o, lse = ops.attn_fwd(q, k, v, B, H, KVH, S, S, hd, scale, True)
    do = torch.randn_like(o)
    dqkv = torch.empty_like(qkv)
    dq, dk, dv = dqkv[:, :qw], dqkv[:, qw:qw + kw], dqkv[:, qw + kw:]

    def unfused():
        ops.attn_bwd(q, k, v, o, do, lse, dq, dk, dv, B, H, KVH, S, S, hd, scale, True)
        ops.rope_(dqkv, S, cos, sin, H + KVH, hd, inverse=True)

    def plain():
        ops.attn_bwd(q, k, v, o, do, lse, dq, dk, dv, B, H, KVH, S, S, hd, scale, True)

    def fused():
        ops.attn_bwd_rope(q, k, v, o, do, lse, dq, dk, dv, B, H, KVH, S, hd, scale, True, cos, sin)

    res = {"unfused": [], "fused": [], "bwd_only": []}
    for _ in range(rounds):
        res["unfused"].append(timeit(unfused))
        res["fused"].append(timeit(fused))
        res["bwd_only"].append(timeit(plain))
    print(json.dumps({"kernel": "attn_bwd + inverse rope", "B": B, "H": H, "KVH": KVH, "S": S, "hd": hd,
                      **{n: round(sorted(v)[len(v) // 2], 4) for n, v in res.items()},
                      "all": {n: [round(x, 4) for x in v] for n, v in res.items()}}), flush=True)


def teacher_case(B=8, size=336):
    """Batched depth teacher (DINOv2-L taps -> mean target features, + DPT decode of the targets) for
    one training batch; the reference runs the same work as a batch-1 Python loop per image."""
    from visper_lm_b200.model.dinov2 import DepthAnythingV2
    from visper_lm_b200.model.dpt import DAv2_Head
    teacher = DepthAnythingV2("vitl", device=dev, with_depth_head=False)
    head = DAv2_Head(dev)
    with torch.no_grad():
        for m in (teacher, head):
            for p_ in m.parameters():
                p_.normal_(0.0, 0.02)
    raw = torch.randint(0, 256, (B, size, size, 3), dtype=torch.uint8, device=dev)
    from visper_lm_b200 import lib
    ms_t = timeit(lambda: teacher.dsg_targets(raw, size))
    lib.reset_launch_count()
    ft = teacher.dsg_targets(raw, size)
    launches = lib.launch_count()
    ms_d = timeit(lambda: head.normalized([ft] * 4))
    from visper_lm_b200.model.gen_teacher import UNCLIP_VIT_H, CLIPVisionModelWithProjection
    enc = CLIPVisionModelWithProjection(UNCLIP_VIT_H, dev)
    with torch.no_grad():
        for p_ in enc.parameters():
            p_.normal_(0.0, 0.02)
    px = torch.randn(B, 3, 224, 224, device=dev).to(BF)
    ms_g = timeit(lambda: enc.image_embeds(px))
    flg = B * 32 * (2.0 * 257 * 12 * 1280 * 1280 + 4.0 * 257 * 257 * 1280)
    print(json.dumps({"kernel": "gen teacher unCLIP ViT-H/14 image_embeds", "B": B, "ms": round(ms_g, 3),
                      "tflops": round(flg / ms_g / 1e9, 1), "images_per_s": round(B / ms_g * 1e3, 1)}), flush=True)
    from visper_lm_b200.model.seg_teacher import OneFormerHead
    seg = OneFormerHead(None, dev)
    with torch.no_grad():
        for n_, p_ in seg.named_parameters():
            p_.fill_(1.0) if ("norm" in n_ and n_.endswith("weight")) else p_.normal_(0.0, 0.02)
    spx = torch.randn(B, 3, 800, 800, device=dev).to(BF)
    ms_s = timeit(lambda: seg.seg_target_rows(spx), iters=5, warmup=2)
    lib.reset_launch_count()
    seg.seg_target_rows(spx)
    print(json.dumps({"kernel": "seg teacher OneFormer Swin-L backbone @800 -> 24x24", "B": B, "ms": round(ms_s, 3),
                      "launches": lib.launch_count(), "images_per_s": round(B / ms_s * 1e3, 1)}), flush=True)
    S = (size // 14) ** 2 + 1
    fl = B * 24 * (2.0 * S * 12 * 1024 * 1024 + 4.0 * S * S * 1024)
    print(json.dumps({"kernel": "depth teacher DINOv2-L (4 taps, mean)", "B": B, "size": size,
                      "ms": round(ms_t, 3), "tflops": round(fl / ms_t / 1e9, 1), "launches": launches,
                      "images_per_s": round(B / ms_t * 1e3, 1), "dpt_decode_ms": round(ms_d, 3)}), flush=True)


if __name__ == "__main__":
    M = 16384
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "gemm"):
        gemm_case(M, 6144, 4096, 0, 0, "llama qkv fwd")
        gemm_case(M, 4096, 4096, 0, 0, "llama o fwd")
        gemm_case(M, 28672, 4096, 0, 0, "llama gate_up fwd")
        gemm_case(M, 4096, 14336, 0, 0, "llama down fwd")
        gemm_case(M, 4096, 6144, 0, 1, "llama qkv dgrad")
        gemm_case(M, 14336, 4096, 0, 1, "llama down dgrad")
        gemm_case(M, 4096, 28672, 0, 1, "llama gate_up dgrad")
        gemm_case(4096, 4096, M, 1, 1, "wgrad 4096x4096")
        gemm_case(8 * 577, 3072, 1024, 0, 0, "clip qkv")
        gemm_case(8192, 8192, 8192, 0, 0, "square 8192")
    if which in ("all", "attn"):
        attn_case(8, 32, 8, 2048, 128, True)
        ops.set_option(ops.OPT_ATTN_LEGACY_FWD, 1)
        ops.set_option(ops.OPT_ATTN_LEGACY_BWD, 1)
        attn_case(8, 32, 8, 2048, 128, True)
        ops.set_option(ops.OPT_ATTN_LEGACY_FWD, 0)
        ops.set_option(ops.OPT_ATTN_LEGACY_BWD, 0)
        attn_case(8, 16, 16, 577, 64, False)
        attn_case(4, 32, 32, 2048, 96, True)
        ops.set_option(ops.OPT_ATTN_LEGACY_FWD, 1)
        ops.set_option(ops.OPT_ATTN_LEGACY_BWD, 1)
        attn_case(4, 32, 32, 2048, 96, True)
        ops.set_option(ops.OPT_ATTN_LEGACY_FWD, 0)
        ops.set_option(ops.OPT_ATTN_LEGACY_BWD, 0)
    if which == "teacher":
        teacher_case()
    if which == "ropebwd":
        rope_bwd_case()
    if which == "attnprof":
        print(json.dumps({"variant": "tc backward v2, column split, P/dS in TMEM (default)"}), flush=True)
        attn_profile()
        ops.set_option(ops.OPT_ATTN_BWD_PINGPONG, 1)
        print(json.dumps({"variant": "tc backward v2, ping-pong groups, P/dS in TMEM"}), flush=True)
        attn_profile()
        ops.set_option(ops.OPT_ATTN_BWD_PINGPONG, 0)
        ops.set_option(ops.OPT_ATTN_BWD_SS, 1)
        print(json.dumps({"variant": "tc backward v2, P/dS through shared memory (SS)"}), flush=True)
        attn_profile()
        ops.set_option(ops.OPT_ATTN_BWD_SS, 0)
        ops.set_option(ops.OPT_ATTN_TC_BWD_V1, 1)
        print(json.dumps({"variant": "tc backward v1"}), flush=True)
        attn_profile()
        ops.set_option(ops.OPT_ATTN_TC_BWD_V1, 0)
    if which in ("all", "swiglu"):
        swiglu_case()
    if which in ("all", "norm"):
        norm_case(16384, 4096)
