timeout 200 python scripts/dev/ext_quick.py cfg4 cfg5 2>&1 | tail -2
CBFNMPC_LIB=$PWD/build/libcbfnmpc_extn.so timeout 200 python scripts/dev/ext_quick.py cfg4 cfg5 2>&1 | tail -2
CBFNMPC_LIB=$PWD/build/libcbfnmpc_extn.so timeout 600 python -m pytest tests/test_gpu_ext.py -x -q 2>&1 | tail -3
