#!/usr/bin/env python3
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) of bench.py: scripts/launch_summary.py <csv> [step kernels...]
Per kernel: launches, mean time, the cluster of the largest launches (the bench-size batches), share of a step."""
import csv, sys, collections
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')) if len(r) > 14 and r[12] == "gpu__time_duration.sum"]
k = collections.defaultdict(list)
for r in rows:
    name = r[4].split("(")[0].replace("sqg::", "").replace("void ", "")
    k[name].append(float(r[14]) / 1e3)
step = ("signal_kernel", "dwell_kernel", "read_offsets_kernel", "tile_desc_kernel", "read_plan_kernel")
def cluster(ts):
    n = [x for x in ts if x >= 0.9 * max(ts)]   # the launches of the device-resident step (the largest batches)
    return (len(n), sum(n) / len(n))
cl = {n: cluster(ts) for n, ts in k.items()}
tot = sum(cl[n][1] for n in cl if n.startswith(step))
for n, ts in sorted(k.items(), key=lambda kv: -sum(kv[1])):
    c = cl[n]
    share = f"share of a step = {100 * c[1] / tot:5.1f}%" if n.startswith(step) else "(not part of a step)"
    print(f"{n[:52]:52s} launches={len(ts):3d} mean={sum(ts) / len(ts):9.1f} us  largest cluster: {c[0]:3d} x {c[1]:9.1f} us  {share}")
