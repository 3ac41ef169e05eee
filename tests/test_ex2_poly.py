"""The FMA-pipe 2^x helper (common.cuh ex2_poly — for softmax loops that are MUFU-bound): the constants in the
CUDA header, evaluated here with the same fp32 / int32 steps in numpy, stay within 1e-4 relative of 2^x — far
below the bf16 rounding of P (3.9e-3)."""
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _consts():
    src = (ROOT / "visper_lm_b200" / "csrc" / "common.cuh").read_text()
    return [np.float32(re.search(rf"#define VPB_EX2_C{i} ([0-9.eE+-]+)f", src).group(1)) for i in range(4)]


def ex2_poly(x, c):
    x = np.asarray(x, np.float32)
    xc = np.maximum(x, np.float32(-126))
    t = (xc + np.float32(12582912.0)).astype(np.float32)
    f = (xc - (t - np.float32(12582912.0)).astype(np.float32)).astype(np.float32)
    p = (c[3] * f + c[2]).astype(np.float32)
    p = (p * f + c[1]).astype(np.float32)
    p = (p * f + c[0]).astype(np.float32)
    r = (p.view(np.int32) + (t.view(np.int32) << 23)).view(np.float32)
    return np.where(x < -126, np.float32(0), r)


def test_polynomial_exp2_accuracy_and_edges():
    c = _consts()
    x = np.float32(-np.random.default_rng(0).uniform(0, 60, 2_000_000))
    rel = np.abs(ex2_poly(x, c) / np.exp2(x.astype(np.float64)) - 1)
    assert rel.max() < 1e-4
    edge = ex2_poly(np.float32([0.0, -1.0, -0.5, -125.9, -126.5, -1e30, -np.inf]), c)
    assert abs(edge[0] - 1) < 1e-4 and abs(edge[1] - 0.5) < 1e-4 and abs(edge[2] - 2 ** -0.5) < 1e-4
    assert edge[3] > 0 and edge[4] == 0 and edge[5] == 0 and edge[6] == 0        # masked scores → exactly 0
    assert np.all(np.diff(ex2_poly(np.float32(np.linspace(-20, 0, 100001)), c)) >= 0)  # monotone across the splits
