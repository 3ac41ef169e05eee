"""ctypes binding of libvisper_b200.so — the only way Python reaches the kernels.

The argument types are parsed from include/visper_b200.h, so the header is the single source of
truth for the C ABI (tests/test_abi.py checks every declared symbol is exported).  There is no CPU
fallback: if the library is missing or fails to load, importing the ops raises.
"""
from __future__ import annotations

import ctypes
import os
import re
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
HEADER = ROOT / "include" / "visper_b200.h"
LIB_PATH = PKG / "libvisper_b200.so"

_CTYPE = {
    "int": ctypes.c_int,
    "int64_t": ctypes.c_int64,
    "float": ctypes.c_float,
}


def parse_header(path: Path = HEADER):
    """Return {name: (restype, [argtypes])} for every function declared in the header."""
    text = path.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    text = "\n".join(l for l in text.splitlines() if not l.strip().startswith("#"))
    decls = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(vpb_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if "*" in ret:
            restype = ctypes.c_char_p if "char" in ret else ctypes.c_void_p
        elif ret == "void":
            restype = None
        else:
            restype = _CTYPE[ret.replace("const", "").strip()]
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    ty = a.replace("const", "").split()[0]
                    argtypes.append(_CTYPE[ty])
        decls[name] = (restype, argtypes)
    return decls


class KernelLibraryError(RuntimeError):
    pass


_lib = None


def load(build_if_missing: bool = True):
    """Load (building first if needed and possible) the kernel library. Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if not build_if_missing:
            raise KernelLibraryError(f"{LIB_PATH} not built; run `python -m visper_lm_b200.build`")
        from . import build as _build

        _build.build()
    try:
        lib = ctypes.CDLL(str(LIB_PATH))
    except OSError as e:  # pragma: no cover
        raise KernelLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (restype, argtypes) in parse_header().items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise KernelLibraryError(f"{LIB_PATH} does not export {name} (stale build?)") from e
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.vpb_abi_version() != 1:
        raise KernelLibraryError("ABI version mismatch")
    _lib = lib
    return lib


def check(status: int, what: str = ""):
    if status != 0:
        msg = load().vpb_last_error()
        raise KernelLibraryError(f"{what} failed ({status}): {msg.decode() if msg else '?'}")


def launch_count() -> int:
    return int(load().vpb_launch_count())


def reset_launch_count() -> None:
    load().vpb_reset_launch_count()
