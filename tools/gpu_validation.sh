#!/bin/bash
# One GPU box, everything the driver runs at round end plus the kernel-level evidence, in this order:
#   gpurun --timeout 2400 -- 'bash tools/gpu_validation.sh'
# (results land in gpurun_out/; copy what should be judged into profiles/)
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -n 4
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout 200 python tools/hbm_kernels_bench.py 2>&1 | grep -v -i warn > $O/hbm_kernels.jsonl; head -3 $O/hbm_kernels.jsonl
timeout 200 python tools/kernel_bench.py attn 2>&1 | grep -v -i warn | grep attention > $O/attn_kernels.jsonl; head -2 $O/attn_kernels.jsonl
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2> $O/bench_ref.err | grep '^{' > $O/bench_ref.json; echo "reference arm exit $?"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 2> $O/bench.err | grep '^{' > $O/bench.json; echo "bench exit $?"
python - <<'PY'
import json
r = json.load(open("gpurun_out/bench_ref.json")); d = json.load(open("gpurun_out/bench.json"))
print("reference: cpu %.4f samples/s (%d steps), gpu %.3f samples/s" % (r["value"], r["steps"], r["reference_gpu"]["value"]))
print("this repo: dsg %.3f (e2e %.3f, %.1f ms/step, %s MHz), ntp %.3f, ift %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"],
      d["clocks"]["sm_mhz"], d["ntp"]["value"], d["ift"]["value"]))
print("e2e / reference_gpu = %.3f" % (d["e2e"]["value"] / r["reference_gpu"]["value"]))
print("rooflines:", d["roofline"]["frac"], {k: round(v["frac"], 3) for k, v in d["roofline"]["kernels"].items()})
PY
