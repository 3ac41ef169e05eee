#!/bin/bash
mkdir -p gpurun_out
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_ntp_n$N.log 2>&1; echo "ntp n$N exit $?"; grep '^{' gpurun_out/bench_ntp_n$N.log | tail -1 > gpurun_out/bench_ntp_n$N.json; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_ntp_n$N.json").read())
print("ntp n$N", round(d["value"],3), round(d["ms_per_step"],1), round(d["e2e"]["value"],3), d["clocks"])
PY
tail -3 gpurun_out/bench_ntp_n$N.log | cut -c1-300
