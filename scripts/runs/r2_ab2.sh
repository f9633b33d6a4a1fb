#!/bin/bash
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
echo "== prev"; SQG_LIB=$PWD/build/libsqg_prev.so python scripts/perf_workloads.py 16384 2>&1 | grep -v "^$" | cut -c1-19,60-200
echo "== new, no L2 window"; SQG_L2_PERSIST=0 python scripts/perf_workloads.py 16384 2>&1 | grep -v "^$" | cut -c1-19,60-200
echo "== new, L2 window"; python scripts/perf_workloads.py 16384 2>&1 | grep -v "^$" | cut -c1-19,60-200
