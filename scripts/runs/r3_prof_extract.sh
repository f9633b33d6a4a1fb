#!/bin/bash
ncu --set full --clock-control none --import-source on -k regex:extract_ -c 2 -f -o gpurun_out/r3_prof_extract python scripts/perf_extract.py > gpurun_out/r3_ncu_extract.log 2>&1
tail -2 gpurun_out/r3_ncu_extract.log
