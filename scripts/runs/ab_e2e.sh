#!/bin/bash
# same-box A/B of the end-to-end legs: previous library vs current, alternating
for i in 1 2; do
  for lib in build/libsqg_prev.so squigulator_b200/libsqg.so; do
    SQG_LIB=$PWD/$lib python bench.py --no-extra --steps 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib', round(d['e2e']['value']/1e9,2), round(d['e2e_svb']['value']/1e9,2), round(d['e2e_coords_svb']['value']/1e9,2), round(d['value']/1e9,1))"
  done
done
