#!/bin/bash
# Dev tool: answers and kernel times of two builds of the library on the same batches (A = build/libcbfnmpc_<tagA>.so, B = the in-tree one).
# usage: bash scripts/dev/ab_compare.sh <tagA> <tagB>       results: gpurun_out/status_<tag>_<B>.npz, compared by scripts/dev/ab_diff.py
A=${1:-v15}; B=${2:-new}
mkdir -p gpurun_out
for n in 4096 32768; do
  CBFNMPC_LIB=build/libcbfnmpc_$A.so timeout 300 python scripts/dev/dump_status.py $n $A
  timeout 300 python scripts/dev/dump_status.py $n $B
done
CBFNMPC_LIB=build/libcbfnmpc_$A.so timeout 300 python scripts/dev/bench_quick.py 4096 1 | tail -3
timeout 300 python scripts/dev/bench_quick.py 4096 1 | tail -3
timeout 300 python scripts/dev/bench_quick.py 32768 1 | tail -2
timeout 300 python scripts/dev/bench_quick.py 4096 0 | tail -2
timeout 300 python scripts/dev/single.py 2>&1 | tail -3
