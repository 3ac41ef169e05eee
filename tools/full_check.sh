#!/bin/bash
# One GPU call that re-validates the whole repo: kernel + parity tests, smoke, the default bench
# line (with the CPU baseline) and the in-step kernel-time profile.  Results land in gpurun_out/.
#   gpurun --timeout 1800 -- 'bash tools/full_check.sh'
mkdir -p gpurun_out
bash tools/gpu_check.sh tests/test_kernels_gpu.py tests/test_parity_gpu.py
python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_ntp.json 2> gpurun_out/bench_ntp.err; echo "bench exit $?"; tail -c 900 gpurun_out/bench_ntp.json
timeout 300 python bench.py --steps 4 --warmup 3 --torch-profile --no-cpu-baseline 2>&1 | grep -v -i warn > gpurun_out/bench_ntp_prof.jsonl; echo "profile exit $?"
timeout 200 python bench.py --workload dsg --teachers --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dsg_teachers.json 2> gpurun_out/bench_dsg_teachers.err; echo "dsg+teachers exit $?"; cut -c1-200 gpurun_out/bench_dsg_teachers.json
