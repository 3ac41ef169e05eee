#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_variants_gpu.py -m gpu -q -p no:cacheprovider -k "attn or attention" 2>&1 | tail -n 4
for v in 1 0 1 0; do
  echo "{\"VPB_ATTN_FWD_NS2\": $v}"
  VPB_ATTN_FWD_NS2=$v timeout 120 python tools/kernel_bench.py attn 2>&1 | grep -v -i warn | grep attention | grep -v "1.27\|1.30\|0.58"
done | tee $O/r2c21_attn_fwd_epilogue_wg_ab.jsonl
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -n 2
timeout 100 python tools/attn_overhead_probe.py 2>&1 | grep "fit\|variant\": 0" | tail -5
