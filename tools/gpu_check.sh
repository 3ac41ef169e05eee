#!/bin/bash
# Runs on the GPU box (via gpurun): per-file pytest with timeouts + kernel micro-benchmarks.
# Everything is logged under gpurun_out/ so a cut-off call can still be read.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for f in "$@"; do
  name=$(basename "$f" .py)
  timeout 900 python -m pytest "$f" -m gpu -q -x --timeout=300 -p no:cacheprovider > "gpurun_out/${name}.log" 2>&1
  echo "== $f exit $?" | tee -a gpurun_out/summary.txt
  tail -n 15 "gpurun_out/${name}.log"
done
