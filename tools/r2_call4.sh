#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_variants_gpu.py tests/test_zero2_gpu.py -m gpu -q -p no:cacheprovider -k "attn or attention or zero2 or sinks or fused or adapter" 2>&1 | tail -n 12
for v in 1 0 1 0; do
  echo "{\"VPB_ATTN_FWD_NS2\": $v}"
  VPB_ATTN_FWD_NS2=$v timeout 120 python tools/kernel_bench.py attn 2>&1 | grep -v -i warn | grep attention | grep -v "1.27\|0.58"
done | tee $O/r2c4_attn_fwd_ns4_ab.jsonl
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -n 3
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:attn_fwd_tc_kernel' -s 3 -c 1 -o $O/r02_attn_fwd_ns4 python tools/kernel_bench.py attnprof > /dev/null 2>&1
ls -la $O/r02_attn_fwd_ns4.ncu-rep
