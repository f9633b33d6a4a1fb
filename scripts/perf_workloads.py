#!/usr/bin/env python3
"""Kernel time of every bench workload with the reference's tables (profiling aid; run on the GPU box):
   scripts/perf_workloads.py [reads] [workload ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json
import numpy as np
import squigulator_b200 as sq
from squigulator_b200.api import PROFILES
from bench import WORKLOADS, synth_reads, load_model, methylate, measured_peak

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
names = sys.argv[2:] or sorted(WORKLOADS)
PEAK = measured_peak()[0] * 1e9
for name in names:
    w = WORKLOADS[name]
    bases, off = synth_reads(n_reads * (4 if w["rna"] else 1), 10000, w["rna"], seed=1)
    if w["meth"]:
        bases = methylate(bases, off, 0.7, seed=2)
    d, f = PROFILES[w["profile"]]
    g = sq.SignalGenerator(dict(d), load_model(name, w["k"], w["meth"])[0], w["k"], flags=f, seed=1, meth=w["meth"])
    b = g.dev_batch(bases, off)
    g.dev_batch_run(b, 3)
    t, tk = g.dev_batch_run(b, 10)
    info = g.dev_batch_info(b)
    alg = 2.0 * info["samples"] + info["kmers"]
    print(f"{name:18s} samples={info['samples']/1e9:.3f}G step={t/10:.3f} ms kernel={tk/10:.3f} ms -> {info['samples']/(t/10*1e-3)/1e9:.0f} Gsamples/s whole step, "
          f"signal kernel {alg/(tk/10*1e-3)/PEAK*100:.1f}% of the HBM roofline")
    g.dev_batch_destroy(b)
    g.close()
