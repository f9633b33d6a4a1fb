#!/bin/bash
SQG_LIB=$PWD/squigulator_b200/libsqg_svb128.so timeout 120 python -m pytest tests/test_gpu_parity.py tests/test_records.py -m gpu -x -q -k "svb or records" 2>&1 | tail -1
for c in 8 12 16; do
  echo "== 128 threads, launched ctas/SM $c"
  SQG_SVB_CTAS=$c SQG_LIB=$PWD/squigulator_b200/libsqg_svb128.so timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:svb_encode --csv python scripts/perf_svb.py 2>/dev/null | grep -a "svb_" | awk -F'","' '{print $5, $NF}' | sed 's/"//g' | sort | uniq -c | head -3
done
