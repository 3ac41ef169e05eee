#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_check.sh tests/test_kernels_gpu.py tests/test_parity_gpu.py
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ntp.json 2> gpurun_out/bench_ntp.err; echo "bench ntp exit $?"; tail -c 1200 gpurun_out/bench_ntp.json
VPB_FUSE_SWIGLU=0 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_ntp_unfused.json 2> gpurun_out/bench_ntp.err; echo "bench unfused exit $?"; tail -c 600 gpurun_out/bench_ntp_unfused.json
VPB_FUSE_SWIGLU_BWD=1 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_ntp_fusedbwd.json 2> gpurun_out/bench_ntp.err; echo "bench fusedbwd exit $?"; tail -c 600 gpurun_out/bench_ntp_fusedbwd.json
timeout 300 python bench.py --steps 4 --warmup 3 --torch-profile --no-cpu-baseline 2>&1 | grep -v -i warn > gpurun_out/bench_ntp_prof.jsonl; echo "prof exit $?"
