#!/bin/bash
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1
for lib in build/libsqg_prev.so squigulator_b200/libsqg.so; do
  echo "== $lib"
  SQG_LIB=$PWD/$lib ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"read_offsets|dwell_kernel|tile_desc|read_plan" -c 12 --csv python bench.py --steps 2 --warmup 1 --no-extra 2>/dev/null | grep -a "offsets\|dwell\|tile_desc\|read_plan" | awk -F'","' '{print $5, $(NF-2), $NF}' | sed 's/"//g' | sort | uniq -c | sort -k2 | head -40
done
