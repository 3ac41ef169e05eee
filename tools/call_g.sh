#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout=120 -k "gemm" -p no:cacheprovider > gpurun_out/gemm_tests.log 2>&1; echo "gemm tests exit $?"; tail -n 15 gpurun_out/gemm_tests.log
timeout 300 python tools/kernel_bench.py gemm > gpurun_out/gemm_bench.jsonl 2>&1; cat gpurun_out/gemm_bench.jsonl
