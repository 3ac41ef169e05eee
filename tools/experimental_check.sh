#!/bin/bash
# First GPU call of the next round: validate the kernel variants that were written after round 1's GPU
# budget ran out (all OFF by default) and A/B them on one box.  Results land in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/experimental_check.sh'   (≈ 18 GPU-minutes; every block is independent — split it if the budget is tight)
mkdir -p gpurun_out
# fastest first: Python-free C-ABI A/B of the variants (bit identity / max diff + alternating timings), ~30 s of GPU.
# Can be run on its own:  gpurun --timeout 120 -- 'tools/_bin/variants_check | tee gpurun_out/variants_check.jsonl'
timeout 180 tools/_bin/variants_check | tee gpurun_out/variants_check.jsonl
VPB_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_experimental_gpu.py tests/test_ex2_poly.py -m gpu -q \
  -p no:cacheprovider > gpurun_out/experimental_tests.log 2>&1; echo "experimental tests exit $?"; tail -n 8 gpurun_out/experimental_tests.log
# polynomial exp2 in the tcgen05 attention kernels: isolated forward / backward times, off vs on (alternating)
for r in 1 2; do for v in 0 1; do
  echo "{\"VPB_ATTN_POLY_EXP2\": $v}" >> gpurun_out/poly_exp2_ab.jsonl
  VPB_ATTN_POLY_EXP2=$v timeout 120 python tools/kernel_bench.py attnprof 2>&1 | grep -v -i warn | grep "attn_" >> gpurun_out/poly_exp2_ab.jsonl
done; done; tail -n 12 gpurun_out/poly_exp2_ab.jsonl
# one-pass window attention in the seg teacher: per-kernel split, default vs the two occupancy variants
for v in 0 1 2; do
  echo "{\"VPB_WIN_ATTN_V2\": $v}" >> gpurun_out/win_attn_v2_ab.jsonl
  VPB_WIN_ATTN_V2=$v timeout 120 python tools/teacher_profile.py seg 2>&1 | grep -v -i warn | head -4 >> gpurun_out/win_attn_v2_ab.jsonl
done; cat gpurun_out/win_attn_v2_ab.jsonl
# whole step, off vs on
for v in 0 1; do
  VPB_ATTN_POLY_EXP2=$v timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_poly_$v.json
  python -c "import json; d=json.load(open('gpurun_out/bench_poly_$v.json')); print('POLY=$v', d['ms_per_step'], d['value'], d['clocks'])"
done
# depthwise 7x7 of the ConvNeXt tower: scalar vs packed-FMA (FFMA2) variant, parity + alternating timings (no Python)
timeout 60 tools/_bin/dwconv_check | tee gpurun_out/dwconv_ffma2_ab.jsonl | tail -n 20
# ConvNeXt-XXL tower in the step (BASELINE configs[3] tower): first bench lines
for v in 0 1; do
  VPB_DWCONV_FFMA2=$v timeout 400 python bench.py --tower convnext-xxl --workload dsg --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_convnext_dsg_ffma2_$v.json
  python -c "import json; d=json.load(open('gpurun_out/bench_convnext_dsg_ffma2_$v.json')); print('convnext dsg FFMA2=$v', d['ms_per_step'], d['value'], d['clocks'])"
done
# per-kernel split of the ConvNeXt-XXL tower (B = 8 @768) + one ncu capture of the depthwise kernel
timeout 200 python tools/teacher_profile.py convnext 2>&1 | grep -v -i warn | head -14 | tee gpurun_out/convnext_tower_kernel_times.jsonl
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dwconv7x7 -c 2 -o gpurun_out/dwconv7x7 tools/_bin/dwconv_check > /dev/null 2>&1; ls -la gpurun_out/dwconv7x7.ncu-rep
timeout 400 python -m pytest tests/test_tower_convnext_gpu.py -m gpu -q -s -p no:cacheprovider 2>&1 | tail -n 15 | tee gpurun_out/convnext_gpu_tests.log
# ViT attention (head_dim 64, non-causal) on the tcgen05 forward kernel: depth teacher (DINOv2-L) per-kernel split and
# the whole step (CLIP ViT-L tower), off vs on
for v in 0 1; do
  echo "{\"VPB_ATTN_FWD_TC64\": $v}" >> gpurun_out/attn_tc64_ab.jsonl
  VPB_ATTN_FWD_TC64=$v timeout 120 python tools/teacher_profile.py depth 2>&1 | grep -v -i warn | head -5 >> gpurun_out/attn_tc64_ab.jsonl
  VPB_ATTN_FWD_TC64=$v timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_tc64_$v.json
  python -c "import json; d=json.load(open('gpurun_out/bench_tc64_$v.json')); print('TC64=$v', d['ms_per_step'], d['value'], d['clocks'])"
done; cat gpurun_out/attn_tc64_ab.jsonl
# CTA-pair GEMM with eight epilogue warps for K <= 1024: seg teacher (Swin-L, K = 192..768 GEMMs) and the ConvNeXt tower, off vs on
for v in 0 1; do
  echo "{\"VPB_GEMM_EPI8\": $v}" >> gpurun_out/gemm_epi8_ab.jsonl
  VPB_GEMM_EPI8=$v timeout 200 python tools/teacher_profile.py seg convnext 2>&1 | grep -v -i warn | grep -i "gpu_busy\|gemm" >> gpurun_out/gemm_epi8_ab.jsonl
done; cat gpurun_out/gemm_epi8_ab.jsonl
# flat grid-stride gather (window partition / reverse / merges of the seg teacher and the ConvNeXt tower), off vs on
for v in 0 1; do
  echo "{\"VPB_GATHER_FLAT\": $v}" >> gpurun_out/gather_flat_ab.jsonl
  VPB_GATHER_FLAT=$v timeout 200 python tools/teacher_profile.py seg 2>&1 | grep -v -i warn | grep -i "gpu_busy\|gather" >> gpurun_out/gather_flat_ab.jsonl
done; cat gpurun_out/gather_flat_ab.jsonl
# Q-in-TMEM attention forward in the whole step (kernel-level numbers come from variants_check above), off vs on
for v in 0 1; do
  VPB_ATTN_FWD_QTM=$v timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_qtm_$v.json
  python -c "import json; d=json.load(open('gpurun_out/bench_qtm_$v.json')); print('QTM=$v', d['ms_per_step'], d['value'], d['clocks'])"
done
