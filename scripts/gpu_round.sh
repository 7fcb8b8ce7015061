#!/bin/bash
# (the full capture is launch #5 of `bench_quick.py 4096 1`: the second device-resident QP->NLP launch over 4096 instances)
# One GPU session: parity tests, bench line, ncu launch list and one full capture of the solve kernel.
# usage: bash scripts/gpu_round.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 4000 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
tail -2 gpurun_out/ncu_launch_$TAG.log
python scripts/dev/bench_quick.py 4096 1 > gpurun_out/plainq_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cbfnmpc_lane -s 5 -c 1 -f -o gpurun_out/prof_$TAG \
    python scripts/dev/bench_quick.py 4096 1 > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
