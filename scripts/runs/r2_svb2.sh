#!/bin/bash
for c in 2 3 4 5; do
  echo "== ctas/SM $c"
  SQG_SVB_CTAS=$c ncu --metrics gpu__time_duration.sum --clock-control none -k regex:svb_encode --csv python scripts/perf_svb.py 2>/dev/null | grep -a "svb_" | awk -F'","' '{print $5, $NF}' | sed 's/"//g' | sort | uniq -c | head -3
done
