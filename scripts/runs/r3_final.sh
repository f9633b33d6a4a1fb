#!/bin/bash
# r03 (second session of round 2): full GPU test suite, smoke, bench line, launch list
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r3_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 > gpurun_out/r3_smoke.log
python bench.py > gpurun_out/r3_bench_n1.json 2> gpurun_out/r3_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 220 --csv --log-file gpurun_out/r3_launches.csv python bench.py --steps 4 --warmup 3 --no-extra > gpurun_out/r3_launch.log 2>&1
cat gpurun_out/r3_tests.log gpurun_out/r3_smoke.log; cut -c1-200 gpurun_out/r3_bench_n1.json
