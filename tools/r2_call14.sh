#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2> $O/r2c14_ref.err | grep '^{' > $O/r2c14_ref.json; echo "ref exit $?"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 2> $O/r2c14_bench.err | grep '^{' > $O/r2c14_bench.json; echo "bench exit $?"; tail -c 600 $O/r2c14_bench.err
python - <<'PY'
import json
r=json.load(open("gpurun_out/r2c14_ref.json")); d=json.load(open("gpurun_out/r2c14_bench.json"))
print("REF cpu", r["value"], r["steps"], r["ms_per_step"], "gpu", r["reference_gpu"]["value"], r["reference_gpu"]["ms_per_step"])
print("DSG", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["clocks"], "ratio vs ref gpu", d["e2e"]["value"]/r["reference_gpu"]["value"])
print("roof", d["roofline"]["frac"], d["roofline"]["step_model_flops_frac"], json.dumps(d["roofline"]["kernels"]))
for k in ("ntp","ift"): print(k, d[k].get("value"), d[k].get("ms_per_step"), d[k].get("e2e",{}).get("value"), d[k].get("error"))
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["step_s"])
PY
