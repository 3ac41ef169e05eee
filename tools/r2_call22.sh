#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -n 4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --extras ntp 2> $O/r2c22_bench.err | grep '^{' > $O/r2c22_bench.json; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2c22_bench.json"))
print("DSG", round(d["value"],3), round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"],3), d["clocks"])
print({k:(round(v["ms_per_step"],2), round(v["frac"],3)) for k,v in d["roofline"]["kernels"].items()})
print("gemm", d["roofline"]["frac"], d["roofline"]["gemm_ms_per_step"], "NTP", d["ntp"]["value"], d["ntp"]["ms_per_step"])
PY
timeout 300 python bench.py --model phi3-mini --seq 4096 --batch 4 --workload ntp --extras none --no-cpu-baseline --steps 6 --warmup 3 2>/dev/null | grep '^{' > $O/r2c22_phi3_4096.json
timeout 300 python bench.py --model phi3-mini --seq 2048 --batch 8 --workload ntp --extras none --no-cpu-baseline --steps 6 --warmup 3 2>/dev/null | grep '^{' > $O/r2c22_phi3_2048.json
python - <<'PY'
import json
for f in ("r2c22_phi3_4096","r2c22_phi3_2048"):
    d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"],2), round(d["ms_per_step"],1), {k:(round(v["ms_per_step"],2), round(v["frac"],3)) for k,v in d["roofline"]["kernels"].items() if "attn" in k})
PY
