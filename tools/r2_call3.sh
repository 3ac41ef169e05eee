#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_variants_gpu.py tests/test_zero2_gpu.py -m gpu -x -q -p no:cacheprovider -k "attn or attention or zero2 or sinks or fused or adapter or rmsnorm" 2>&1 | tail -n 15
for v in 1 0 1 0; do
  echo "{\"VPB_ATTN_FWD_V1\": $v}"
  VPB_ATTN_FWD_V1=$v timeout 120 python tools/kernel_bench.py attn 2>&1 | grep -v -i warn | grep attention
done | tee $O/r2c3_attn_fwd_v3_ab.jsonl
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -n 3
