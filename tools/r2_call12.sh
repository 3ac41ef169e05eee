#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 8 --warmup 3 2> $O/r2c12_bench_n2.err | grep '^{' > $O/r2c12_bench_n2.json; echo "n2 exit $?"; tail -c 1500 $O/r2c12_bench_n2.err | grep -v "^\*\|OMP"
timeout 400 python bench.py --workload ntp --train full --batch 4 --extras none --no-cpu-baseline --steps 8 --warmup 3 2> $O/r2c12_ift_n1.err | grep '^{' > $O/r2c12_ift_n1.json; echo "ift n1 exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2c12_bench_n2.json"))
print("DSG", round(d["value"],2), round(d["ms_per_step"],1), round(d["e2e"]["value"],2), d["comm"], d["clocks"]["sm_mhz"])
for k in ("ntp","ift"):
    e=d.get(k,{}); print(k.upper(), e.get("value"), e.get("ms_per_step"), e.get("comm"), e.get("peak_mem_gb"), e.get("error"))
print("DPCHECK", json.dumps(d.get("dp_check")))
e=json.load(open("gpurun_out/r2c12_ift_n1.json")); print("IFT N=1", e["value"], e["ms_per_step"], e["peak_mem_gb"], e["clocks"]["sm_mhz"])
PY
