#!/bin/bash
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r2l_bench_n1.json 2> gpurun_out/r2l_bench_n1.err; cut -c1-160 gpurun_out/r2l_bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 220 --csv --log-file gpurun_out/r2l_launches.csv python bench.py --steps 4 --warmup 3 --no-extra > gpurun_out/r2l_launch.log 2>&1
