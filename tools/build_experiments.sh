#!/bin/bash
# Build the default-off experiment variants next to the product library (CPU, nvcc only).  The .so files are git-ignored
# and travel to the GPU box with the snapshot; select one with GB200_LIB=gypsum_b200/exp_<name>.so.
set -e
cd "$(dirname "$0")/../gypsum_b200/csrc"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -shared -cudart static"
SRC="kernels.cu tracker.cu bits.cu fused.cu engine.cu"
nvcc $FLAGS -DGB_W2048_LAYOUT_B=1 -o ../exp_layout_b.so $SRC
nvcc $FLAGS -DGB_SPEC_ALIAS=1 -o ../exp_spec_alias.so $SRC
nvcc $FLAGS -DGB_W2048_LAYOUT_B=1 -DGB_SPEC_ALIAS=1 -o ../exp_both.so $SRC
nvcc $FLAGS -DGB_TRACK_FAST_ANGLE=1 -o ../exp_fast_angle.so $SRC
ls -la ../exp_*.so
