#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout=200 -k "attention" -p no:cacheprovider > gpurun_out/attn_tests.log 2>&1; echo "attn tests exit $?"; tail -n 12 gpurun_out/attn_tests.log
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x --timeout=300 -p no:cacheprovider > gpurun_out/parity.log 2>&1; echo "parity exit $?"; tail -n 5 gpurun_out/parity.log
timeout 600 python bench.py --model phi3-mini --seq 4096 --batch 4 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_phi3_4096.json 2> gpurun_out/bench_phi3.err; echo "phi3 exit $?"; python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_phi3_4096.json") if l.startswith("{")][-1])
print("phi3 T=4096 B=4", round(d["value"],3), round(d["ms_per_step"],1), d["clocks"]["sm_mhz"], round(d["roofline"]["gemm_share_of_step"],3))
PY
timeout 600 python bench.py --model phi3-mini --seq 2048 --batch 8 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_phi3_2048.json 2> gpurun_out/bench_phi3.err; python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_phi3_2048.json") if l.startswith("{")][-1])
print("phi3 T=2048 B=8", round(d["value"],3), round(d["ms_per_step"],1), d["clocks"]["sm_mhz"], round(d["roofline"]["gemm_share_of_step"],3))
PY
