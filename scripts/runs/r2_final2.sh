#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err; cut -c1-200 gpurun_out/r2h_bench_n1.json; tail -2 gpurun_out/r2h_bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2h_bench_ref.json 2> gpurun_out/r2h_bench_ref.err; cut -c1-200 gpurun_out/r2h_bench_ref.json
