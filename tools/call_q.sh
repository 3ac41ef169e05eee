#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/hf_library_baseline.py 2>&1 | grep -v -i warn | tail -2 | tee gpurun_out/hf_baseline_ckpt.json
timeout 600 python tools/hf_library_baseline.py --no-ckpt 2>&1 | grep -v -i warn | tail -2 | tee gpurun_out/hf_baseline_nockpt.json
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ntp_samebox.json 2>/dev/null; python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_ntp_samebox.json") if l.startswith("{")][-1])
print("ours same box", round(d["value"],3), round(d["ms_per_step"],1), d["clocks"]["sm_mhz"])
PY
