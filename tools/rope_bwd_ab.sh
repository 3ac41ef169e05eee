#!/bin/bash
# full re-validation + same-box A/B of the fused inverse-RoPE attention-backward epilogue
bash tools/full_check.sh
for f in 0 1; do
  VPB_FUSE_ROPE_BWD=$f timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_rope_bwd_$f.json
  python - <<PY
import json; d=json.load(open("gpurun_out/bench_rope_bwd_$f.json")); print("FUSE_ROPE_BWD=$f", d["ms_per_step"], d["value"], d["gpu_launches"], d["clocks"])
PY
done
