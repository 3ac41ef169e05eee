"""Blackwell-native instruction counts per kernel of the shipped library (no GPU needed):
    python tools/sass_mnemonics.py > profiles/r02_sass_mnemonics.txt
UTCHMMA = tcgen05.mma (bf16), .2CTA = cta_group::2; UTMALDG = TMA load; LDTM / STTM = tcgen05.ld / st;
UTCBAR = tcgen05.commit; SYNCS = mbarrier ops; HMMA = mma.sync (the legacy path); FFMA2 = packed fp32 FMA."""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "visper_lm_b200" / "libvisper_b200.so"
WANT = ("UTCHMMA.2CTA", "UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "SYNCS", "HMMA", "MUFU.EX2", "FFMA2", "LDGSTS")


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = per.setdefault(re.sub(r"\(.*", "", name), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            for w in WANT:
                if op == w or op.startswith(w + "."):
                    if w == "UTCHMMA" and ".2CTA" in op:
                        w = "UTCHMMA.2CTA"
                    cur[w] += 1
                    break
    total = collections.Counter()
    for c in per.values():
        total.update(c)
    print(f"# cuobjdump -sass {LIB.name} (tools/sass_mnemonics.py): Blackwell-native instruction counts per kernel")
    print("# UTCHMMA = tcgen05.mma (bf16), .2CTA = cta_group::2; UTMALDG = TMA load; LDTM / STTM = tcgen05.ld / st; "
          "UTCBAR = tcgen05.commit; SYNCS = mbarrier; HMMA = mma.sync (legacy path)")
    print("TOTAL", dict(total))
    for name in sorted(per):
        if per[name]:
            print(name, "", dict(per[name]))


if __name__ == "__main__":
    sys.exit(main())
