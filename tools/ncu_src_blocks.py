"""Summarise an `ncu --page source --csv` export: samples per block of SASS instructions with the
notable opcodes in each block (which warp role / phase the time goes to)."""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
step = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
data = [r for r in rows[2:] if len(r) >= len(hdr) - 2 and r[0].startswith("0x")]
iS, isrc, iE = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
I = lambda x: int(x) if x.strip().lstrip("-").isdigit() else 0
# keep only the first copy if the listing is duplicated
n = len(data)
if n % 2 == 0 and all(data[i][isrc] == data[i + n // 2][isrc] for i in range(0, n // 2, 97)):
    data = data[: n // 2]
tot = sum(I(r[iS]) for r in data)
print("total samples", tot, "instructions", len(data))
print(sorted(((sum(I(r[i]) for r in data), h) for i, h in stall), reverse=True)[:8])
for s in range(0, len(data), step):
    blk = data[s:s + step]
    ops = []
    for r in blk:
        t = r[isrc].strip().split()
        if not t:
            continue
        op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
        if op.startswith(("UTC", "UTMA", "SYNCS", "LDTM", "STTM", "STS", "LDS", "LDG", "STG", "BAR", "MUFU", "FENCE", "UBLK")):
            ops.append(op.split(".")[0] + ("." + op.split(".")[1] if "." in op else ""))
    print(s, sum(I(r[iS]) for r in blk), max(I(r[iE]) for r in blk), dict(Counter(ops)))
