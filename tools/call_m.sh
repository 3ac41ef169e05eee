#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout=120 -k "gemm" -p no:cacheprovider > gpurun_out/gemm_tests.log 2>&1; echo "gemm tests exit $?"; tail -n 3 gpurun_out/gemm_tests.log
for rep in 1 2; do
for v in 32 16 64; do
VPB_GEMM_PANEL_MB=$v timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_p${v}_${rep}.json 2> gpurun_out/bench.err; python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_p${v}_${rep}.json") if l.startswith("{")][-1])
print("panel_mb=$v rep=$rep", round(d["value"],3), round(d["ms_per_step"],1), d["clocks"]["sm_mhz"], round(d["roofline"]["achieved"]), round(d["roofline"]["gemm_ms_per_step"],1))
PY
done
done
