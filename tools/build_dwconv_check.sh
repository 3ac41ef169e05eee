#!/bin/bash
# Builds the stand-alone C-ABI checks tools/_bin/{dwconv_check,variants_check} (git-ignored, they travel with gpurun)
# against the in-tree library.
set -e
NVCC=${NVCC:-$(command -v nvcc || echo /usr/local/cuda/bin/nvcc)}
cd "$(dirname "$0")/.."
${PYTHON:-python} -m visper_lm_b200.build > /dev/null
mkdir -p tools/_bin
gcc -O2 -c oracle/c/dwconv_ref.c -o tools/_bin/dwconv_ref.o
$NVCC -gencode arch=compute_100a,code=sm_100a -O2 -Iinclude tools/dwconv_check.cu tools/_bin/dwconv_ref.o \
  -o tools/_bin/dwconv_check -Lvisper_lm_b200 -l:libvisper_b200.so -Xlinker -rpath -Xlinker '$ORIGIN/../../visper_lm_b200'
$NVCC -gencode arch=compute_100a,code=sm_100a -O2 -Iinclude tools/variants_check.cu \
  -o tools/_bin/variants_check -Lvisper_lm_b200 -l:libvisper_b200.so -Xlinker -rpath -Xlinker '$ORIGIN/../../visper_lm_b200'
echo tools/_bin/dwconv_check tools/_bin/variants_check
