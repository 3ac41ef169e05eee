#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x --timeout=400 -p no:cacheprovider > gpurun_out/parity.log 2>&1; echo "parity exit $?"; tail -n 25 gpurun_out/parity.log
timeout 600 python bench.py --steps 6 --warmup 3 --workload dsg --no-cpu-baseline > gpurun_out/bench_dsg.json 2> gpurun_out/bench_dsg.err; echo "bench dsg exit $?"; tail -c 600 gpurun_out/bench_dsg.json; tail -3 gpurun_out/bench_dsg.err
