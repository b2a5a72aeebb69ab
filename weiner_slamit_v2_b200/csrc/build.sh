#!/bin/sh
# Builds liborb_b200.so (the C-ABI library, include/orb_b200.h) for sm_100a, in-tree.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
SRCS="orb_extractor.cu"
[ -f orb_matcher.cu ] && SRCS="$SRCS orb_matcher.cu orb_matcher_proj.cu orb_matcher_bow.cu"
[ -f orb_frame.cu ] && SRCS="$SRCS orb_frame.cu"
[ -f orb_stereo.cu ] && SRCS="$SRCS orb_stereo.cu"
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false \
      -Xcompiler -fPIC -shared -o ../liborb_b200.so $SRCS "$@"
