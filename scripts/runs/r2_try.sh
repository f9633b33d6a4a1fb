#!/bin/bash
# parity of the kernel under test, then kernel times: previous build vs this one
python -m pytest tests/test_gpu_parity.py tests/test_law.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -3
bash scripts/runs/r2_ab.sh build/libsqg_prev.so squigulator_b200/libsqg.so "$@"
