#!/bin/bash
# round 2, GPU call 1: the prepared-variant A/Bs of tools/experimental_check.sh, then an ncu --set full capture of the
# HBM-bound kernels at production row shapes (2-layer Llama-3-8B-width dsg step), then the whole GPU suite.
mkdir -p gpurun_out
bash tools/experimental_check.sh 2>&1 | tee gpurun_out/r2_call1_experimental.log
timeout 500 ncu --set full --clock-control none --import-source on \
  -k 'regex:norm_fwd|norm_bwd|ce_fwd_bwd|distill|gather_rows|swiglu_bwd|adamw|rope_inplace|colsum' \
  --profile-from-start off -c 90 -o gpurun_out/r02_hbm_kernels \
  python bench.py --layers 2 --workload dsg --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --profile > gpurun_out/r02_hbm_ncu.log 2>&1
ls -la gpurun_out/r02_hbm_kernels.ncu-rep
timeout 300 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -n 5
