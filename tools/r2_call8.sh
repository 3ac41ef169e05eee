#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -n 12
timeout 200 python tools/hbm_kernels_bench.py 2>&1 | grep -v -i warn | tee $O/r2c8_hbm_kernels.jsonl | head -6
