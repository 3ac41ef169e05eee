#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -n 6
timeout 200 python tools/teacher_profile.py seg depth gen 2>&1 | grep -v -i warn | grep "gpu_busy" | tee $O/r2c13_teachers.jsonl
timeout 200 python tools/teacher_profile.py seg 2>&1 | grep -v -i warn | head -8 | tee -a $O/r2c13_teachers.jsonl
