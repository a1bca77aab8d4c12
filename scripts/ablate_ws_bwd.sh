#!/bin/bash
# Developer aid: builds of the library whose pipelined backward skips one piece of work (WRONG results, timing only):
# py4cast_b200/lib/libpy4cast_b200_abl<N>.so with N = 1 (no dgamma / dbeta column sums), 2 (no dh1 store),
# 4 (no dX epilogue); time them with P4C_LIB_AB=libpy4cast_b200_abl<N>.so python scripts/bwd_micro.py ...
set -e
cd "$(dirname "$0")/../py4cast_b200/csrc"
make -s >/dev/null
for n in "$@"; do
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr \
    -DP4C_WB_ABLATE=$n -c mlp2_ws_bwd.cu -o /tmp/mlp2_ws_bwd_abl$n.o
  objs=$(ls ../../build/csrc/*.o | grep -v mlp2_ws_bwd.o)
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/libpy4cast_b200_abl$n.so $objs /tmp/mlp2_ws_bwd_abl$n.o
done
ls ../lib
