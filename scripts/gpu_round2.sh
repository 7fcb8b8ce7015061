#!/bin/bash
# One GPU session (round 2): bench line, ncu launch list of the same command, one full capture of the faithful solve kernel and one
# of the extended-mode kernel (config-4 teams).  usage: bash scripts/gpu_round2.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
set -x
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 600 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
tail -2 gpurun_out/ncu_launch_$TAG.log
python scripts/dev/bench_quick.py 4096 1 > gpurun_out/plainq_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cbfnmpc_lane -s 5 -c 1 -f -o gpurun_out/prof_$TAG \
    python scripts/dev/bench_quick.py 4096 1 > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
python scripts/dev/ext_quick.py cfg4 > gpurun_out/plainx_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cbfnmpc_ext_kernel -s 2 -c 1 -f -o gpurun_out/prof_ext_$TAG \
    python scripts/dev/ext_quick.py cfg4 > gpurun_out/ncu_fullx_$TAG.log 2>&1
tail -3 gpurun_out/ncu_fullx_$TAG.log
cp srb-cbf-nmpc_b200/lib/libcbfnmpc.so gpurun_out/libcbfnmpc_$TAG.so
