#!/bin/bash
# GPU call: kernel + parity tests, NTP bench, in-step kernel profile, dsg bench, fused-vs-unfused A/B
mkdir -p gpurun_out
bash tools/gpu_check.sh tests/test_kernels_gpu.py tests/test_parity_gpu.py
timeout 600 python bench.py --steps 6 --warmup 3 > gpurun_out/bench_ntp.json 2> gpurun_out/bench_ntp.err; echo "bench ntp exit $?"; tail -c 600 gpurun_out/bench_ntp.json
timeout 300 python bench.py --steps 4 --warmup 3 --torch-profile --no-cpu-baseline > gpurun_out/bench_ntp_prof.jsonl 2>&1; echo "prof exit $?"
VPB_FUSE_SWIGLU=0 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_ntp_unfused.json 2>&1; echo "unfused exit $?"; tail -c 300 gpurun_out/bench_ntp_unfused.json
timeout 400 python bench.py --steps 6 --warmup 3 --workload dsg --no-cpu-baseline > gpurun_out/bench_dsg.json 2> gpurun_out/bench_dsg.err; echo "bench dsg exit $?"; tail -c 400 gpurun_out/bench_dsg.json
