#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout=200 -k "attention" -p no:cacheprovider > gpurun_out/attn_tests.log 2>&1; echo "attn tests exit $?"; tail -n 12 gpurun_out/attn_tests.log
timeout 300 python tools/kernel_bench.py attnprof > gpurun_out/attnprof.jsonl 2>&1; echo "attnprof exit $?"; cat gpurun_out/attnprof.jsonl | tail -20
