#!/bin/bash
for v in "" _ko_A; do
  echo "== variant $v"
  SQG_LIB=$PWD/squigulator_b200/libsqg$v.so python scripts/perf_matrix.py 16384 2>&1 | grep noisy
done
