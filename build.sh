#!/bin/bash
# Builds libmcbb_b200.so (sm_100a only) in-tree; the .so travels to the GPU box with the snapshot.
set -e
cd "$(dirname "$0")"
SRC=mcbb.jl_b200/csrc
OUT=mcbb.jl_b200/libmcbb_b200.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
CCBIN=/usr/bin/g++
[ -x $CCBIN ] || CCBIN=g++
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -ccbin $CCBIN -Xcompiler -fPIC,-O2 --expt-relaxed-constexpr -Xptxas -v"
mkdir -p build
pids=()
for f in api solve cluster batched batched_imma batched_umma batched_umma_sl multi dist_mat hist group peaks; do
  ( $NVCC $FLAGS -c $SRC/$f.cu -o build/$f.o > build/$f.log 2>&1 || { cat build/$f.log; exit 1; } ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -ccbin $CCBIN -o $OUT build/api.o build/solve.o build/cluster.o build/batched.o build/batched_imma.o build/batched_umma.o build/batched_umma_sl.o build/multi.o build/dist_mat.o build/hist.o build/group.o build/peaks.o -lcudart_static -lpthread -ldl -lrt
echo "built $OUT"
