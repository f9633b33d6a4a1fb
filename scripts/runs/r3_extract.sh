#!/bin/bash
timeout 300 python -m pytest tests/test_read_extraction.py -m gpu -x -q 2>&1 | tail -3
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:extract --csv python scripts/perf_extract.py 2>/dev/null | grep -a "extract_" | awk -F'","' '{print $5, $NF}' | sed 's/"//g' | sort | uniq -c | head -12
