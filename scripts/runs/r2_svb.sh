#!/bin/bash
python -m pytest tests/test_gpu_parity.py tests/test_records.py tests/test_gpu_full_size.py -m gpu -x -q -k "svb or records or full" 2>&1 | tail -1
for lib in build/libsqg_prev.so squigulator_b200/libsqg.so "$@"; do
  echo "== $lib"
  SQG_LIB=$PWD/$lib ncu --metrics gpu__time_duration.sum --clock-control none -k regex:svb_encode --csv python scripts/perf_svb.py 2>/dev/null | grep -a "svb_" | awk -F'","' '{print $5, $NF}' | sed 's/"//g' | sort | uniq -c | head -4
done
