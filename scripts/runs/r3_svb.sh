#!/bin/bash
# svb-zd encoder: parity tests, then its kernel time (ncu launch list of scripts/perf_svb.py: 4096 reads, 0.53 G samples), pipelined kernel and the old one
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_records.py tests/test_svb_zd.py tests/test_gpu_full_size.py -m gpu -x -q -k "svb or records or full" 2>&1 | tail -2
for v in "" 1; do
echo "== SQG_SVB_V1=$v"
[ -n "$v" ] && export SQG_SVB_V1=1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:svb_encode --csv python scripts/perf_svb.py 2>/dev/null | grep -a "svb_" | awk -F'","' '{print $5, $NF}' | sed 's/"//g' | sort | uniq -c | head -8
done
