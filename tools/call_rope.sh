#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout=120 -k "gemm or rope" -p no:cacheprovider > gpurun_out/gemm_tests.log 2>&1; echo "gemm/rope tests exit $?"; tail -n 4 gpurun_out/gemm_tests.log
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x --timeout=300 -p no:cacheprovider > gpurun_out/parity.log 2>&1; echo "parity exit $?"; tail -n 3 gpurun_out/parity.log
for v in 1 0; do
VPB_FUSE_ROPE=$v timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_rope${v}.json 2> gpurun_out/bench.err; python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_rope${v}.json") if l.startswith("{")][-1])
print("fuse_rope=$v", round(d["value"],3), round(d["ms_per_step"],1), d["clocks"]["sm_mhz"], round(d["roofline"]["gemm_ms_per_step"],1), d["gpu_launches"])
PY
done
