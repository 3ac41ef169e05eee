#!/bin/bash
mkdir -p gpurun_out
for v in 0 2 3 0; do
VPB_GEMM_L2_HINTS=$v timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_l2h${v}.json 2> gpurun_out/bench.err; python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_l2h${v}.json") if l.startswith("{")][-1])
print("l2_hints=$v", round(d["value"],3), round(d["ms_per_step"],1), d["clocks"]["sm_mhz"], round(d["roofline"]["achieved"]), round(d["roofline"]["gemm_ms_per_step"],1))
PY
done
