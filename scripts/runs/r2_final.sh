#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; cut -c1-200 gpurun_out/r2f_bench_n1.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; cut -c1-200 gpurun_out/r2f_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 4 --warmup 3 --no-extra > gpurun_out/r2f_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:signal_kernel -s 3 -c 1 -f -o gpurun_out/r2f_prof python bench.py --steps 2 --warmup 3 --no-extra --reads-per-step 16384 > gpurun_out/r2f_ncu.log 2>&1
for v in "" _ko_PHASEA _ko_EMIT _ko_STORE _ko_MAP _ko_GATHER; do
  echo "== variant $v"
  SQG_LIB=$PWD/squigulator_b200/libsqg$v.so python scripts/perf_workloads.py 16384 dna-r10-prom dna-r9-prom 2>&1 | grep -v "^$"
done > gpurun_out/r2f_ko.log 2>&1
cat gpurun_out/r2f_ko.log
