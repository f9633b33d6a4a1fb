#!/bin/bash
# (a) full ncu capture of the signal kernel at the bench size (traffic per launch for the bench line)
ncu --set full --clock-control none --import-source on -k regex:signal_kernel -s 3 -c 1 -f -o gpurun_out/r2g_prof python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/r2g_ncu.log 2>&1
# (b) full capture of the svb encoder
ncu --set full --clock-control none --import-source on -k regex:svb_encode -c 1 -f -o gpurun_out/r2g_prof_svb python scripts/perf_svb.py > gpurun_out/r2g_ncu_svb.log 2>&1
tail -2 gpurun_out/r2g_ncu_svb.log
# (c) sanitizer
S=gpurun_out/r2g_san
timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > ${S}_mem_smoke.log 2>&1; grep -a "ERROR SUMMARY\|smoke ok" ${S}_mem_smoke.log | tail -2
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_records.py -m gpu -x -q -k "random_profiles or equals_oracle or golden or records or svb or ss_text or edge or sample_range or dwell_extremes or window" > ${S}_mem_tests.log 2>&1; grep -a "ERROR SUMMARY\|passed\|failed" ${S}_mem_tests.log | tail -3
timeout 600 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > ${S}_race_smoke.log 2>&1; grep -a "RACECHECK SUMMARY\|smoke ok" ${S}_race_smoke.log | tail -2
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py tests/test_records.py -m gpu -x -q -k "random_profiles or svb or ss_text or records or window" > ${S}_race_tests.log 2>&1; grep -a "RACECHECK SUMMARY\|passed\|failed" ${S}_race_tests.log | tail -3
