#!/bin/bash
# sanitizer over what changed in the last session: the read-extraction kernels (vector path, scan) and the signal kernel's emit loop
S=gpurun_out/r3_san
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_read_extraction.py -m gpu -x -q > ${S}_mem_extract.log 2>&1; grep -a "ERROR SUMMARY\|passed\|failed" ${S}_mem_extract.log | tail -3
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_read_extraction.py -m gpu -x -q > ${S}_race_extract.log 2>&1; grep -a "RACECHECK SUMMARY\|passed\|failed" ${S}_race_extract.log | tail -3
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "random_profiles or edge or window" > ${S}_mem_k4.log 2>&1; grep -a "ERROR SUMMARY\|passed\|failed" ${S}_mem_k4.log | tail -3
timeout 900 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > ${S}_race_smoke.log 2>&1; grep -a "RACECHECK SUMMARY\|smoke ok" ${S}_race_smoke.log | tail -2
