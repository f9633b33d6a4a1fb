#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 3 > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; cut -c1-200 gpurun_out/r2b_bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 4 --warmup 3 --no-extra > gpurun_out/r2b_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:signal_kernel -s 3 -c 1 -f -o gpurun_out/r2b_prof python bench.py --steps 2 --warmup 3 --no-extra --reads-per-step 16384 > gpurun_out/r2b_ncu.log 2>&1
ls -la gpurun_out/
