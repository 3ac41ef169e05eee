#!/bin/bash
# Attention kernels only: unit tests + per-kernel device times of every variant.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout=200 -k "attention" -p no:cacheprovider > gpurun_out/attn_tests.log 2>&1; echo "attn tests exit $?"; tail -n 12 gpurun_out/attn_tests.log
timeout 300 python tools/kernel_bench.py attnprof 2>&1 | grep -v -i warn > gpurun_out/attnprof.jsonl; echo "attnprof exit $?"; tail -30 gpurun_out/attnprof.jsonl
