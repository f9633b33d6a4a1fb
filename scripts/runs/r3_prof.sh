#!/bin/bash
ncu --set full --clock-control none --import-source on -k regex:signal_kernel -s 3 -c 1 -f -o gpurun_out/r3_prof python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/r3_ncu.log 2>&1
tail -2 gpurun_out/r3_ncu.log; ls -la gpurun_out/r3_prof.ncu-rep
