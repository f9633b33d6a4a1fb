#!/bin/bash
# same-box A/B of signal-kernel builds: scripts/runs/r3_ab.sh lib1 lib2 ...   (kernel time and roofline fraction per workload), then parity of the last one
for rep in 1 2; do
for lib in "$@"; do
  echo "== $lib"
  SQG_LIB=$PWD/$lib python scripts/perf_workloads.py 16384 dna-r10-prom dna-r9-prom rna004-prom 2>&1 | grep -v "^$" | cut -c1-19,60-200
done
done
last="${@: -1}"
SQG_LIB=$PWD/$last timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
