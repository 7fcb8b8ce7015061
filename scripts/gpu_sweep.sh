set -x; mkdir -p gpurun_out
for w in 28 14 7 4; do echo "== warps/CTA $w"; CBFNMPC_LIB=$PWD/srb-cbf-nmpc_b200/lib/exp_w$w.so python scripts/dev/bench_quick.py 4096 2>&1 | grep "mode 1" | tail -1; done
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err; tail -c 2500 gpurun_out/bench_v2.json; tail -5 gpurun_out/bench_v2.err
python scripts/dev/bench_quick.py 4096 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:cbfnmpc_solve -s 5 -c 1 -o gpurun_out/prof_v2 python scripts/dev/bench_quick.py 4096 > gpurun_out/ncu_v2.log 2>&1; tail -3 gpurun_out/ncu_v2.log
