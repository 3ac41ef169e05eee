#!/bin/bash
# round 2, GPU call 2: whole GPU suite (incl. the full-size configs[0] parity test and the ZeRO-2 paths), the new
# default bench line (dsg + ntp extra + real cpu_baseline), the reference arm (CPU configs[0] + the reference's GPU
# build), A/Bs of the new norm kernels and the SwiGLU-backward fusion, HBM-kernel roofline numbers + ncu, and one
# source-level ncu capture of the three tcgen05 attention kernels.
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider -s > $O/r2c2_gpu_tests.log 2>&1; echo "gpu tests exit $?"
grep -E "^\[config0\]|\[parity\]|passed|failed|Error" $O/r2c2_gpu_tests.log | tail -40
timeout 120 python tools/hbm_kernels_bench.py 2>&1 | grep -v -i warn | tee $O/r2c2_hbm_kernels.jsonl
timeout 700 python bench.py --steps 8 --warmup 3 > $O/r2c2_bench_dsg.json 2> $O/r2c2_bench_dsg.err; echo "bench exit $?"; tail -c 1500 $O/r2c2_bench_dsg.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2c2_bench_dsg.json'))
    print('DSG', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'])
    print('NTP', d.get('ntp',{}).get('value'), d.get('ntp',{}).get('ms_per_step'))
    print('roof', json.dumps(d['roofline']['kernels']))
    print('cpu', json.dumps(d.get('cpu_baseline'))[:600])
except Exception as e: print('parse failed', e)
PY
timeout 900 python bench.py --impl reference --steps 8 --warmup 3 > $O/r2c2_bench_ref.json 2> $O/r2c2_bench_ref.err; echo "ref exit $?"; tail -c 1200 $O/r2c2_bench_ref.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2c2_bench_ref.json'))
    print('REF cpu', d['value'], d['ms_per_step'], d['steps'], d['cpu_baseline']['cores'])
    print('REF gpu', json.dumps(d.get('reference_gpu'))[:900])
except Exception as e: print('parse failed', e)
PY
for tag in "VPB_NORM_LEGACY=1" "VPB_NORM_LEGACY=0" "VPB_FUSE_SWIGLU_BWD=1" "VPB_NORM_LEGACY=1" "VPB_NORM_LEGACY=0" "VPB_FUSE_SWIGLU_BWD=1"; do
  env $tag timeout 300 python bench.py --workload ntp --extras none --no-cpu-baseline --steps 8 --warmup 3 2>/dev/null > $O/ab.json
  python -c "import json; d=json.load(open('gpurun_out/ab.json')); k=d['roofline']['kernels']; print('$tag', round(d['ms_per_step'],2), round(d['value'],3), d['clocks']['sm_mhz'], {n:(round(v['ms_per_step'],2), round(v['frac'],3)) for n,v in k.items()})" | tee -a $O/r2c2_ab.txt
done
timeout 400 ncu --set full --clock-control none --import-source off \
  -k 'regex:norm|swiglu_bwd|ce_fwd_bwd|distill|gather_rows|adamw|rope_inplace|colsum|sumsq' -c 40 \
  -o $O/r02_hbm python tools/hbm_kernels_bench.py --once > $O/r02_hbm_ncu.log 2>&1
ncu -i $O/r02_hbm.ncu-rep --page raw --csv > $O/r02_hbm_raw.csv 2>/dev/null; ls -la $O/r02_hbm.ncu-rep $O/r02_hbm_raw.csv
rm -f $O/r02_hbm.ncu-rep
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:attn_fwd_tc_kernel|attn_bwd_dkdv_tc2|attn_bwd_dq_tc2' -s 3 -c 3 \
  -o $O/r02_attn python tools/kernel_bench.py attnprof > $O/r02_attn_ncu.log 2>&1
ls -la $O/
du -sh $O
