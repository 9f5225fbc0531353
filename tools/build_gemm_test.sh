#!/bin/bash
# builds tools/_bin/gemm_test (standalone check of csrc/gemm_tc.cu; ships to the GPU box with the snapshot)
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/_bin
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -I include \
  -DSNB_API='extern "C"' tools/gemm_test.cu sparenet_b200/csrc/gemm_tc.cu -o tools/_bin/gemm_test -lcublas
