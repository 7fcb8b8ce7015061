set -x; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/dev/bench_quick.py 4096 2>&1 | tail -6
python scripts/dev/bench_quick.py 65536 2>&1 | grep "mode 1" | tail -2
