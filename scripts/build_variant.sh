#!/bin/bash
# Diagnostic build of the library: scripts/build_variant.sh <suffix> [-Dflags...]; load the result with NA_LIB_PATH.
#   trace: -DNA_TM_TRACE (per-tile clock trace of the TMEM kernel, read with scripts/trace_show.py)
set -e
SUF=$1; shift
cd "$(dirname "$0")/../nerf-art_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --compiler-options -fPIC -shared "$@" \
  -o ../libnerfart_b200_$SUF.so api.cu mlp_simt.cu volsdf_render.cu neus_render.cu surface_render.cu mlp_tc.cu mlp_tmem.cu train.cu clip_vit.cu tgemm.cu wgrad_f16.cu
echo built libnerfart_b200_$SUF.so
