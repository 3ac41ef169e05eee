#!/bin/bash
# 2-GPU data-parallel check: NTP and dsg benches under torchrun (NCCL), NCCL_DEBUG for NVLS evidence
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/bench_ntp_n2.log 2>&1; echo "ntp n2 exit $?"; grep '^{' gpurun_out/bench_ntp_n2.log | tail -1 > gpurun_out/bench_ntp_n2.json; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_ntp_n2.json").read())
print("ntp n2", round(d["value"],3), round(d["ms_per_step"],1), d["e2e"]["value"], d["clocks"])
PY
grep -i "NVLS\|nvls" gpurun_out/bench_ntp_n2.log | head -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 6 --warmup 3 --workload dsg > gpurun_out/bench_dsg_n2.log 2>&1; echo "dsg n2 exit $?"; grep '^{' gpurun_out/bench_dsg_n2.log | tail -1 > gpurun_out/bench_dsg_n2.json; tail -c 400 gpurun_out/bench_dsg_n2.json
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ntp_n1_samebox.json 2>/dev/null; python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_ntp_n1_samebox.json") if l.startswith("{")][-1])
print("ntp n1 same box", round(d["value"],3), round(d["ms_per_step"],1))
PY
