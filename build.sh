#!/bin/bash
# Builds dentist_b200/libdentist_b200.so for sm_100a (in-tree, travels with gpurun snapshots).
set -e
cd "$(dirname "$0")"
OUT=dentist_b200/libdentist_b200.so
SRC="dentist_b200/csrc/scan.cu dentist_b200/csrc/radix.cu dentist_b200/csrc/seed.cu dentist_b200/csrc/segsort.cu dentist_b200/csrc/extend.cu dentist_b200/csrc/engine.cu dentist_b200/csrc/api.cu dentist_b200/csrc/api_pile.cu dentist_b200/csrc/pile.cu dentist_b200/csrc/dust.cu dentist_b200/csrc/chain.cu dentist_b200/csrc/collect.cu dentist_b200/csrc/dazzdb.cpp"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -rdc=false \
  -Xcompiler -fPIC,-Wall,-Wno-unused-function -shared -cudart static ${DN_NVCC_EXTRA} -o $OUT $SRC
echo built $OUT
