"""CPU test: the C-ABI shared library builds, loads, and exports every symbol include/visper_b200.h
declares (no compute without a GPU)."""
import ctypes

from visper_lm_b200 import build, lib


def test_library_builds_and_exports_header_symbols():
    path = build.build()
    assert path.exists()
    decls = lib.parse_header()
    assert len(decls) >= 35
    handle = ctypes.CDLL(str(path))
    missing = [n for n in decls if not hasattr(handle, n)]
    assert not missing, missing
    L = lib.load()
    assert L.vpb_abi_version() == 1
    assert lib.launch_count() >= 0


def test_errors_are_reported_not_swallowed():
    L = lib.load()
    # an invalid problem must be rejected before any launch: negative status + message
    rc = L.vpb_gemm_bf16(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0)
    assert rc != 0
    assert b"gemm" in L.vpb_last_error()


def test_product_never_imports_oracle():
    import pathlib
    import re

    root = pathlib.Path(lib.__file__).resolve().parent
    for f in root.rglob("*.py"):
        assert not re.search(r"^\s*(from|import)\s+oracle\b", f.read_text(), re.M), f
