timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5
timeout 600 python bench.py --no-extended > gpurun_out/bench_v22b.json 2> gpurun_out/bench_v22b.err; tail -3 gpurun_out/bench_v22b.err
timeout 100 python scripts/dev/step_overhead.py 2>&1 | tail -8
