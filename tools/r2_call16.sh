#!/bin/bash
# evidence pass: ncu launch list of the bench command (share of the step per kernel), ncu --set full of the HBM-bound
# kernels of the final tree, and of the dominant GEMM launches inside the step (roofline.traffic)
mkdir -p gpurun_out
O=gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 3200 --csv \
  --log-file $O/r02_launches_dsg.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --extras none --profile > $O/r02_launches_dsg.log 2>&1
ls -la $O/r02_launches_dsg.csv
timeout 400 ncu --set full --clock-control none --import-source off \
  -k 'regex:norm|swiglu_bwd|ce_fwd_bwd|distill|gather_rows|adamw|rope|colsum|sumsq' -c 36 \
  -o $O/r02_hbm2 python tools/hbm_kernels_bench.py --once > $O/r02_hbm2_ncu.log 2>&1
ncu -i $O/r02_hbm2.ncu-rep --page raw --csv > $O/r02_hbm2_raw.csv 2>/dev/null; rm -f $O/r02_hbm2.ncu-rep; ls -la $O/r02_hbm2_raw.csv
timeout 600 ncu --set full --clock-control none --import-source off --profile-from-start off -k 'regex:gemm_tcgen05_pair' -s 330 -c 8 \
  -o $O/r02_gemm_instep python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --extras none --profile > $O/r02_gemm_instep.log 2>&1
ncu -i $O/r02_gemm_instep.ncu-rep --page raw --csv > $O/r02_gemm_instep_raw.csv 2>/dev/null; rm -f $O/r02_gemm_instep.ncu-rep; ls -la $O/r02_gemm_instep_raw.csv
du -sh $O
