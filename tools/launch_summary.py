"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: launch_summary.py launches.csv out.csv [skip_first_n] [take_n]"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if r]
hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hdr_i]
iname, imet, ival = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
iunit = hdr.index("Metric Unit")
launches = []
for r in rows[hdr_i + 1:]:
    if len(r) <= ival or r[imet] != "gpu__time_duration.sum":
        continue
    v = float(r[ival].replace(",", ""))
    u = r[iunit]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v if u == "ms" else v * 1e3
    name = r[iname].split("(")[0]
    launches.append((name, ms))
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
take = int(sys.argv[4]) if len(sys.argv) > 4 else len(launches)
sel = launches[skip:skip + take]
agg = OrderedDict()
for n, ms in sel:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "launches", "total_ms", "share"])
    for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        w.writerow([n, c, f"{ms:.3f}", f"{ms / tot:.4f}"])
print(f"{len(launches)} launches in file, {len(sel)} summarised, {tot:.1f} ms")
