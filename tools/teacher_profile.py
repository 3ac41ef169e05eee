"""Per-kernel device times of the three frozen teachers (and, with `convnext`, of the ConvNeXt-XXL tower) on one
batch of 8 (torch profiler / CUPTI):  python tools/teacher_profile.py [seg|depth|gen|convnext ...]"""
import json
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from visper_lm_b200.model.dinov2 import DepthAnythingV2  # noqa: E402
from visper_lm_b200.model.gen_teacher import UNCLIP_VIT_H, CLIPVisionModelWithProjection  # noqa: E402
from visper_lm_b200.model.seg_teacher import OneFormerHead  # noqa: E402

dev = torch.device("cuda:0")
B = 8


def init(m):
    with torch.no_grad():
        for n, p in m.named_parameters():
            p.fill_(1.0) if ("norm" in n and n.endswith("weight")) else p.normal_(0.0, 0.02)
    return m


def _convnext():
    from types import SimpleNamespace

    from visper_lm_b200.model.convnext import CLIPConvNextVisionTower

    return CLIPConvNextVisionTower("CLIP-convnext_xxlarge-res768", args=SimpleNamespace(mm_vision_select_layer=-2), device=dev)


# built lazily: only the requested models are allocated ("convnext" = the ConvNeXt-XXL vision tower, not a teacher)
cases = {
    "seg": lambda: (init(OneFormerHead(None, dev)), torch.randn(B, 3, 800, 800, device=dev).to(torch.bfloat16),
                    lambda m, x: m.seg_target_rows(x)),
    "depth": lambda: (init(DepthAnythingV2("vitl", device=dev, with_depth_head=False)),
                      torch.randint(0, 256, (B, 336, 336, 3), dtype=torch.uint8, device=dev),
                      lambda m, x: m.dsg_targets(x, 336)),
    "gen": lambda: (init(CLIPVisionModelWithProjection(UNCLIP_VIT_H, dev)), torch.randn(B, 3, 224, 224, device=dev),
                    lambda m, x: m.image_embeds(x)),
    "convnext": lambda: (init(_convnext()), torch.randn(B, 3, 768, 768, device=dev).to(torch.bfloat16),
                         lambda m, x: m(x)),
}
which = sys.argv[1:] or ["seg", "depth", "gen"]
for name in which:
    m, x, fn = cases[name]()
    for _ in range(2):
        fn(m, x)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            fn(m, x)
        torch.cuda.synchronize()
    evs = [e for e in prof.key_averages() if e.device_time_total > 0]
    tot = sum(e.device_time_total for e in evs) / 3e3
    print(json.dumps({"teacher": name, "B": B, "gpu_busy_ms_per_batch": round(tot, 3)}), flush=True)
    for e in sorted(evs, key=lambda e: -e.device_time_total)[:12]:
        print(json.dumps({"teacher": name, "kernel": e.key[:80], "calls_per_batch": e.count // 3,
                          "ms_per_batch": round(e.device_time_total / 3e3, 3)}), flush=True)
