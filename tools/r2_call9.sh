#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_zero2_gpu.py -m gpu -x -q -p no:cacheprovider -k "rmsnorm or zero2 or sinks or fused or adapter" 2>&1 | tail -n 6
timeout 200 python tools/hbm_kernels_bench.py 2>&1 | grep -v -i warn | tee $O/r2c10_hbm_kernels.jsonl | head -7
for tag in "VPB_NORM_R1=1" "VPB_NORM_R1=0" "VPB_NORM_R1=1" "VPB_NORM_R1=0"; do
  env $tag timeout 300 python bench.py --workload ntp --extras none --no-cpu-baseline --steps 8 --warmup 3 2>/dev/null > $O/ab.json
  python -c "import json; d=json.load(open('gpurun_out/ab.json')); k=d['roofline']['kernels']; print('$tag', round(d['ms_per_step'],2), round(d['value'],3), d['clocks']['sm_mhz'], {n:(round(v['ms_per_step'],2), round(v['frac'],3)) for n,v in k.items()})" | tee -a $O/r2c10_ab.txt
done
