#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
for v in 1 0; do
VPB_FUSE_SWIGLU_BWD=$v timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_fb${v}_${rep}.json 2> gpurun_out/bench.err; python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_fb${v}_${rep}.json") if l.startswith("{")][-1])
print("fuse_bwd=$v rep=$rep", round(d["value"],3), round(d["ms_per_step"],1), d["clocks"]["sm_mhz"], round(d["roofline"]["achieved"]), round(d["roofline"]["gemm_ms_per_step"],1))
PY
done
done
