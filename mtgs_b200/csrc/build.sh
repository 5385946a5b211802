#!/bin/bash
# Builds mtgs_b200/libb200splat.so for sm_100a (in-tree; the .so travels to the GPU box with the snapshot).
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libb200splat.so
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O2 --shared -ccbin /usr/bin/g++"
$NVCC $FLAGS ${B2S_PTXAS_V:+-Xptxas -v} -o $OUT api.cu project.cu depthsort.cu tilelists.cu blend.cu sh.cu exchange.cu ssim.cu losses.cu optim.cu arena.cu
echo "built $(readlink -f $OUT)"
