#!/bin/bash
# Dev tool: kernel times of two builds of the library (A = build/libcbfnmpc_<tagA>.so, B = the in-tree one), no dumps.
A=${1:-v19}
for n in 4096 32768; do
  CBFNMPC_LIB=build/libcbfnmpc_$A.so timeout 300 python scripts/dev/bench_quick.py $n 1 | tail -3 | head -2
  timeout 300 python scripts/dev/bench_quick.py $n 1 | tail -3 | head -2
done
timeout 300 python scripts/dev/bench_quick.py 4096 0 | tail -2 | head -1
timeout 300 python scripts/dev/single.py 2>&1 | tail -1
