#!/bin/bash
# kernel times of several builds of the library on all workloads, same box
for lib in "$@"; do
  echo "== $lib"
  SQG_LIB=$PWD/$lib python scripts/perf_workloads.py 16384 2>&1 | grep -v "^$" | cut -c1-19,60-200
done
