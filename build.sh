#!/bin/bash
# Builds dentist_b200/libdentist_b200.so for sm_100a (in-tree, travels with gpurun snapshots).
set -e
cd "$(dirname "$0")"
OUT=dentist_b200/libdentist_b200.so
SRC="dentist_b200/csrc/scan.cu dentist_b200/csrc/radix.cu dentist_b200/csrc/seed.cu dentist_b200/csrc/segsort.cu dentist_b200/csrc/extend.cu dentist_b200/csrc/engine.cu dentist_b200/csrc/api.cu dentist_b200/csrc/api_pile.cu dentist_b200/csrc/pile.cu dentist_b200/csrc/dust.cu dentist_b200/csrc/chain.cu dentist_b200/csrc/collect.cu dentist_b200/csrc/maskcov.cu dentist_b200/csrc/dazzdb.cpp"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -rdc=false \
  -Xcompiler -fPIC,-Wall,-Wno-unused-function -shared -cudart static ${DN_NVCC_EXTRA} -o $OUT $SRC
echo built $OUT
# command-line stand-ins for the tools the workflow calls directly (tools/dn_cli.cpp)
mkdir -p bin
i=0
for t in dn-damapper dn-daligner dn-dbdust; do
  g++ -O2 -std=c++17 -Iinclude -DDN_TOOL=$i tools/dn_cli.cpp -o bin/$t -Ldentist_b200 -ldentist_b200 -Wl,-rpath,'$ORIGIN/../dentist_b200' -ldl -lpthread -lrt
  i=$((i+1))
done
echo built bin/dn-damapper bin/dn-daligner bin/dn-dbdust
