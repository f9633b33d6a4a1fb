#!/bin/bash
# same-box A/B of the device-resident step: previous library vs 16 warps (smaller buffers) vs 18 warps, alternating
for lib in squigulator_b200/libsqg.so build/libsqg_w18.so; do
SQG_LIB=$PWD/$lib python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "equals_oracle or random_profiles or dwell_extremes or edge or sample_range" 2>&1 | tail -1
done
for i in 1 2; do
  for lib in build/libsqg_prev.so squigulator_b200/libsqg.so build/libsqg_w18.so; do
    SQG_LIB=$PWD/$lib python bench.py --steps 30 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib', round(d['value']/1e9,1), round(d['roofline']['kernel_ms'],4), {k:round(v['kernel_ms'],4) for k,v in d['other_workloads'].items()})"
  done
done
