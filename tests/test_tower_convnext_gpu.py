"""GPU parity of the CLIP-ConvNeXt tower through the C ABI: the depthwise 7x7 kernel
against torch's conv2d, and CLIPConvNextVisionTower against the CPU oracle on identical bf16-rounded weights
and pixels — the miniature of tests/golden/convnext_mini_96.pt (also against the library's golden vectors)
and the ConvNeXt-XXL geometry (depths 3-4-30-3, dims 384..3072) at 128 px.

Tolerances: relative Frobenius error per stage.  A CPU emulation of the kernels' bf16 rounding points
(tests/test_convnext_cpu.py::_TorchOps) gives 0.4e-2 … 0.8e-2 on the miniature and up to 1.3e-2 at XXL depth,
so 2.5e-2 / 3e-2 leave 2-3x head-room while any layout error is O(1)."""
from types import SimpleNamespace

import pytest
import torch
import torch.nn.functional as F

from parity_utils import restate

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
GOLDEN = __import__("pathlib").Path(__file__).parent / "golden"


@pytest.mark.parametrize("B,H,W,C", [(1, 5, 9, 64), (2, 24, 24, 128), (1, 17, 35, 192), (3, 48, 48, 1536)])
def test_dwconv7x7_vs_torch(B, H, W, C):
    from visper_lm_b200 import ops

    g = torch.Generator().manual_seed(B * 1000 + C)
    x = torch.randn(B * H * W, C, generator=g).to(BF)
    w = (torch.randn(C, 1, 7, 7, generator=g) / 7).to(BF)
    b = (0.3 * torch.randn(C, generator=g)).to(BF)
    w49 = w.reshape(C, 49).t().contiguous()
    got = ops.dwconv7x7(x.cuda(), w49.cuda(), b.cuda(), B, H, W, C).float().cpu()
    want = F.conv2d(x.float().view(B, H, W, C).permute(0, 3, 1, 2), w.float(), b.float(), padding=3, groups=C)
    want = want.permute(0, 2, 3, 1).reshape(B * H * W, C)
    err = (got - want).abs()
    assert not (err > 1e-3 + 0.0079 * want.abs()).any(), err.max().item()   # 1 bf16 ulp of the result
    nb = ops.dwconv7x7(x.cuda(), w49.cuda(), None, B, H, W, C).float().cpu()
    assert not ((nb - (want - b.float())).abs() > 2e-2 + 0.0079 * want.abs()).any()


def test_dwconv7x7_rejects_bad_arguments():
    from visper_lm_b200 import lib, ops

    x = torch.zeros(4 * 4, 32, dtype=BF, device="cuda")
    w = torch.zeros(49, 32, dtype=BF, device="cuda")
    with pytest.raises(lib.KernelLibraryError):
        ops.dwconv7x7(x, w, None, 1, 4, 4, 32)                    # C % 64 != 0
    x = torch.zeros(4 * 4, 64, dtype=BF, device="cuda")
    w = torch.zeros(49, 64, dtype=BF, device="cuda")
    with pytest.raises(lib.KernelLibraryError):
        ops.dwconv7x7(x, w, None, 1, 4, 4, 64, out=x)             # in place


def _tower(cfg, size):
    from oracle.make_golden_convnext import PREFIX
    from visper_lm_b200.model.convnext import CLIPConvNextVisionTower

    m = CLIPConvNextVisionTower(f"CLIP-convnext_xxlarge-res{size}", args=SimpleNamespace(mm_vision_select_layer=-2),
                                device="cuda", cfg=dict(cfg, image_size=size))
    with torch.no_grad():
        for n, p in m.vision_tower.named_parameters():
            p.copy_(restate.convnext_seeded_state_one(PREFIX + n, tuple(p.shape)))
    return m, PREFIX


@pytest.mark.parametrize("case", ["mini_96", "xxl_128"])
def test_convnext_tower_vs_oracle(case):
    from oracle.make_golden_convnext import pixels
    from visper_lm_b200 import lib

    if case == "mini_96":
        fx = torch.load(GOLDEN / "convnext_mini_96.pt")
        cfg, size, B, seed, tol = fx["config"], fx["size"], fx["B"], fx["seed"], 2.5e-2
    else:
        fx = None
        cfg, size, B, seed, tol = dict(restate.CONVNEXT_XXL), 128, 2, 773, 3e-2
    m, prefix = _tower(cfg, size)
    px = pixels(B, size, seed)
    lib.reset_launch_count()
    rows, H, W, stages = m.vision_tower.forward_rows(px, return_stages=True)
    torch.cuda.synchronize()
    n_blocks = sum(cfg["depths"])
    assert lib.launch_count() >= 4 + 3 * 6 + 4 * n_blocks          # cast, im2col, GEMM, LN; 3 downsamples; 4 per block
    sd = {prefix + n: p.detach().float().cpu() for n, p in m.vision_tower.named_parameters()}
    with torch.no_grad():
        want = restate.convnext_stage_features(sd, px.to(BF).float(), cfg, prefix)
    for i, ((x, h, w), ref) in enumerate(zip(stages, want)):
        got = x.float().cpu().view(B, h, w, -1).permute(0, 3, 1, 2)
        assert got.shape == ref.shape and torch.isfinite(got).all()
        rel = ((got - ref).norm() / ref.norm()).item()
        print(f"convnext {case} stage {i}: rel Frobenius {rel:.3e}")
        assert rel < tol, (i, rel)
    out = m(px)                                                    # the tower's public forward: B-major rows
    assert out.shape == (B * (size // 32) ** 2, cfg["dims"][-1])
    ref_rows = restate.convnext_tower(sd, px.to(BF).float(), cfg, prefix).reshape(out.shape)
    assert ((out.float().cpu() - ref_rows).norm() / ref_rows.norm()).item() < tol
    if fx is not None:                                             # the third-party library's own outputs (fp32 weights)
        for i in range(4):
            got = stages[i][0].float().cpu().view(B, stages[i][1], stages[i][2], -1).permute(0, 3, 1, 2)
            sub = got[:, ::4, ::2, ::2] if i < 3 else got
            assert ((sub - fx["stages_sub"][i]).norm() / fx["stages_sub"][i].norm()).item() < 3.5e-2


def test_step_with_convnext_tower_vs_oracle():
    """BASELINE config 4 at test size: tiny Llama + dsg heads behind the ConvNeXt tower — loss against the CPU
    oracle on identical bf16-rounded weights / inputs (the north_star's 1e-3 rel is the expectation and is
    printed; the assertion is 2e-3 until this new case has been calibrated on hardware like the eight cases of
    test_parity_gpu.py), projector gradient (which sees the tower's features) aligned with the oracle's."""
    from parity_utils import (build_product, configs, cos_sim, oracle_state, pt_freeze, round_batch, run_product)

    cfg = configs.TINY_LLAMA_CONVNEXT
    model = build_product(cfg, True, "cuda:0")
    pt_freeze(model)
    batch = round_batch(configs.synthetic_batch(cfg, 2, 48, seed=4321))
    out = run_product(model, batch, True, "cuda:0")
    out.loss.backward()
    torch.cuda.synchronize()
    sd = oracle_state(model)
    req = {n: sd[n].clone().requires_grad_(True) for n in ("model.mm_projector.0.weight", "model.mm_projector.2.weight")}
    ref = restate.forward_step({**sd, **req}, cfg, batch, distill=True)
    ref["loss"].backward()
    got, want = out.loss.item(), ref["loss"].item()
    print(f"convnext step: loss {got:.6f} oracle {want:.6f} rel {abs(got - want) / abs(want):.2e}")
    assert abs(got - want) <= 2e-3 * abs(want), (got, want)
    for n, r in req.items():
        g = dict(model.named_parameters())[n].grad
        assert g is not None and cos_sim(g, r.grad) > 0.99, (n, cos_sim(g, r.grad))
