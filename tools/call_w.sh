#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x --timeout=300 -p no:cacheprovider > gpurun_out/parity.log 2>&1; echo "parity exit $?"; tail -n 5 gpurun_out/parity.log
( time timeout 900 python bench.py --impl reference --steps 8 --warmup 3 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_ref.err ) 2>&1 | grep real; tail -c 700 gpurun_out/bench_reference_arm.json
