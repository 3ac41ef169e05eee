#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout=200 -k "attention" -p no:cacheprovider > gpurun_out/attn_tests.log 2>&1; echo "attn tests exit $?"; tail -n 6 gpurun_out/attn_tests.log
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x --timeout=300 -p no:cacheprovider > gpurun_out/parity.log 2>&1; echo "parity exit $?"; tail -n 6 gpurun_out/parity.log
timeout 600 python bench.py --model phi3-mini --seq 4096 --batch 4 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_phi3_4096.json 2> gpurun_out/bench_phi3.err; echo "phi3 exit $?"; tail -c 900 gpurun_out/bench_phi3_4096.json; tail -3 gpurun_out/bench_phi3.err
