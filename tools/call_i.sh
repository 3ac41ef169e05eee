#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout=120 -k "swiglu" -p no:cacheprovider > gpurun_out/swiglu_tests.log 2>&1; echo "swiglu tests exit $?"; tail -n 8 gpurun_out/swiglu_tests.log
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x --timeout=300 -p no:cacheprovider > gpurun_out/parity.log 2>&1; echo "parity exit $?"; tail -n 5 gpurun_out/parity.log
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ntp.json 2> gpurun_out/bench_ntp.err; echo "bench ntp exit $?"; tail -c 700 gpurun_out/bench_ntp.json
