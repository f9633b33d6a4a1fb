#!/usr/bin/env python3
"""Per-source-line view of an ncu capture: scripts/ncu_lines.py <rep> [min_pct]  (instructions executed and stall samples)"""
import csv, io, subprocess, sys
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur = hdr = None; agg = []
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if r and r[0] == "Function Name": continue
    if hdr and r and r[0].isdigit() and r[2] == "-":
        ci = hdr.index("Instructions Executed"); si = hdr.index("# Samples")
        st = {k: int(r[hdr.index(k)] or 0) for k in ("stall_long_sb", "stall_short_sb", "stall_wait", "stall_math", "stall_mio", "stall_lg", "stall_not_selected", "stall_selected", "stall_dispatch", "stall_no_inst", "stall_branch_resolving", "stall_barrier", "stall_membar")}
        agg.append((cur, int(r[0]), r[1].strip(), int(r[ci]), int(r[si]), st))
tot = sum(a[3] for a in agg); ts = sum(a[4] for a in agg)
print("warp instructions", tot, "samples", ts)
tst = {}
for a in agg:
    for k, v in a[5].items(): tst[k] = tst.get(k, 0) + v
print("stalls:", ", ".join(f"{k[6:]} {100*v/ts:.1f}%" for k, v in sorted(tst.items(), key=lambda kv: -kv[1])))
for a in agg:
    if a[3] > thr / 100 * tot or a[4] > thr / 100 * ts:
        top = sorted(a[5].items(), key=lambda kv: -kv[1])[:2]
        print(f"{a[0][:12]:12s} {a[1]:4d} {100*a[3]/tot:5.1f}% inst {100*a[4]/ts:5.1f}% time [{', '.join(f'{k[6:]} {v}' for k, v in top)}]  {a[2][:90]}")
