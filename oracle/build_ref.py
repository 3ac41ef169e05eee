"""TEST / BENCH INFRASTRUCTURE ONLY — recipe that makes the UNMODIFIED reference travel to the GPU box.

The reference (SHI-Labs/VisPer-LM) is pure Python, so "building" it is a byte-for-byte copy of its
package sources from where they lie (/root/reference/ola_vlm/**/*.py) into the git-ignored directory
oracle/_ref/ola_vlm/ — the Python analogue of compiling a C reference into oracle/_ref/*.so.  Nothing is
edited; `oracle/ref_manifest.json` (committed: paths + sha256 only, no sources) lets anyone check on the
box that what runs there is the published code.  oracle/_ref/ is listed in .gitignore (sources never
enter the history) but not in .gpurunignore (it ships with the snapshot like the built .so files).

Consumers: oracle/ref_shim.py (falls back to oracle/_ref when /root/reference is not mounted), and through
it only tests/, __graft_entry__.smoke() and bench.py's reference / cpu_baseline legs.  The product package
never imports it (tests/test_abi.py).

    python -m oracle.build_ref            # run by __graft_entry__.build() when /root/reference exists
    python -m oracle.build_ref --verify   # compare oracle/_ref against the committed manifest
"""
from __future__ import annotations

import hashlib
import json
import shutil
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
SRC = Path("/root/reference")
DST = HERE / "_ref"
MANIFEST = HERE / "ref_manifest.json"
PACKAGE = "ola_vlm"


def _sha(path: Path) -> str:
    return hashlib.sha256(path.read_bytes()).hexdigest()


def reference_files(root: Path):
    return sorted(p for p in (root / PACKAGE).rglob("*.py") if "__pycache__" not in p.parts)


def build(verbose: bool = False) -> Path | None:
    """Copy the reference package into oracle/_ref/ and (re)write the manifest.  Returns the destination,
    or None when the reference tree is not mounted (the GPU box: it uses the copy that came with the
    snapshot)."""
    if not (SRC / PACKAGE).is_dir():
        return None
    files = reference_files(SRC)
    manifest = {}
    for src in files:
        rel = src.relative_to(SRC)
        dst = DST / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        if not dst.exists() or dst.read_bytes() != src.read_bytes():
            shutil.copyfile(src, dst)
        manifest[str(rel)] = _sha(src)
    # drop files that no longer exist upstream
    for old in reference_files(DST) if (DST / PACKAGE).is_dir() else []:
        if str(old.relative_to(DST)) not in manifest:
            old.unlink()
    text = json.dumps({"source": "SHI-Labs/VisPer-LM @ /root/reference", "files": manifest}, indent=1, sort_keys=True)
    if not MANIFEST.exists() or MANIFEST.read_text() != text:
        MANIFEST.write_text(text)
    if verbose:
        print(f"oracle/_ref: {len(manifest)} reference files", file=sys.stderr)
    return DST


def verify() -> list[str]:
    """Names of files under oracle/_ref that are missing or differ from the committed manifest."""
    want = json.loads(MANIFEST.read_text())["files"]
    bad = []
    for rel, sha in want.items():
        p = DST / rel
        if not p.exists() or _sha(p) != sha:
            bad.append(rel)
    return bad


def available() -> bool:
    return (DST / PACKAGE / "model" / "language_model" / "ola_llama.py").exists()


if __name__ == "__main__":
    if "--verify" in sys.argv:
        bad = verify()
        print("ok" if not bad else f"MISMATCH: {bad}")
        sys.exit(1 if bad else 0)
    print(build(verbose=True))
