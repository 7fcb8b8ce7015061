timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --no-extended > gpurun_out/bench_v22.json 2> gpurun_out/bench_v22.err; tail -c 1500 gpurun_out/bench_v22.json; tail -3 gpurun_out/bench_v22.err
