#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout=300 -p no:cacheprovider > gpurun_out/kernel_tests.log 2>&1; echo "kernel tests exit $?"; tail -n 5 gpurun_out/kernel_tests.log
timeout 300 python tools/kernel_bench.py attnprof 2>&1 | grep -v Warn > gpurun_out/attnprof.jsonl; echo "attnprof exit $?"; cat gpurun_out/attnprof.jsonl | tail -12
timeout 300 python tools/kernel_bench.py gemm > gpurun_out/gemm_bench.jsonl 2>&1; cat gpurun_out/gemm_bench.jsonl
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ntp.json 2> gpurun_out/bench_ntp.err; echo "bench ntp exit $?"; tail -c 1500 gpurun_out/bench_ntp.json
