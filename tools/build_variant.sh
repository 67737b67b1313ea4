#!/bin/sh
# build an experimental variant of the library into ab/ (git-ignored): tools/build_variant.sh NAME -DFLAG ...
set -e
NAME=$1; shift
cd "$(dirname "$0")/../ei-keyword-spotting_b200/csrc"
mkdir -p ../../ab
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -fmad=false \
  -Xcompiler -fPIC,-ffp-contract=off,-Wno-unused-function -I../../include -I. -shared -o ../../ab/libeikws_$NAME.so \
  api.cpp plan.cpp model_graph.cpp tflm_capture.cpp kernels.cu "$@" -Xptxas -v 2> ../../ab/$NAME.log
grep -A2 "kernelIsLb1ELi2" ../../ab/$NAME.log | grep -E "Used|spill" | head -3
