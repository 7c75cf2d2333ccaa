#!/bin/bash
# Builds dentist_b200/libdentist_b200.so for sm_100a (in-tree, travels with gpurun snapshots).
# One object per source, compiled in parallel and only when the source (or a header) is newer.
set -e
cd "$(dirname "$0")"
OUT=dentist_b200/libdentist_b200.so
CS=dentist_b200/csrc
OBJ=build/obj
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -rdc=false -Xcompiler -fPIC,-Wall,-Wno-unused-function ${DN_NVCC_EXTRA}"
mkdir -p $OBJ bin
newest_hdr=$(ls -t $CS/*.cuh $CS/*.hpp include/*.h 2>/dev/null | head -1)
pids=()
objs=()
for src in $CS/*.cu $CS/*.cpp; do
  o=$OBJ/$(basename $src).o
  objs+=($o)
  if [ ! -f $o ] || [ $src -nt $o ] || [ $newest_hdr -nt $o ] || [ build.sh -nt $o ]; then
    $NVCC $FLAGS -c $src -o $o &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o $OUT "${objs[@]}" -ldl -lpthread
echo built $OUT
# command-line stand-ins for the tools the workflow calls directly (tools/dn_cli.cpp)
i=0
for t in dn-damapper dn-daligner dn-dbdust; do
  g++ -O2 -std=c++17 -Iinclude -DDN_TOOL=$i tools/dn_cli.cpp -o bin/$t -Ldentist_b200 -ldentist_b200 -Wl,-rpath,'$ORIGIN/../dentist_b200' -ldl -lpthread -lrt &
  i=$((i+1))
done
wait
echo built bin/dn-damapper bin/dn-daligner bin/dn-dbdust
