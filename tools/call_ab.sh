#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
for v in 0 1; do
VPB_ATTN_BWD_PINGPONG=$v timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_pp${v}_${rep}.json 2> gpurun_out/bench.err; python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_pp${v}_${rep}.json") if l.startswith("{")][-1])
print("pingpong=$v rep=$rep", round(d["value"],3), round(d["ms_per_step"],1), d["clocks"]["sm_mhz"], round(d["roofline"]["gemm_ms_per_step"],1), "non-gemm", round(d["ms_per_step"]-d["roofline"]["gemm_ms_per_step"],1))
PY
done
done
