#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_variants_gpu.py -m gpu -x -q -p no:cacheprovider -k "attn or attention or persistent" 2>&1 | tail -n 6
for v in 1 0 1 0; do
  echo "{\"VPB_ATTN_BWD_DQ_R1\": $v}"
  VPB_ATTN_BWD_DQ_R1=$v timeout 120 python tools/kernel_bench.py attn 2>&1 | grep -v -i warn | grep attention | grep -v "1.27\|1.30\|1.29\|0.58\|hd\": 64"
done | tee $O/r2c23_attn_dq_persist_ab.jsonl
VPB_ATTN_BWD_DQ_R1=0 timeout 120 python tools/kernel_bench.py attnprof 2>&1 | grep -v -i warn | head -6
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -n 2
