#!/usr/bin/env python3
"""svb-zd / records kernels on one batch of 4096 reads (profiling aid; run under ncu --metrics gpu__time_duration.sum)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import squigulator_b200 as sq
from squigulator_b200.api import PROFILES, WANT_SVB
from bench import synth_reads, load_model
bases, off = synth_reads(4096, 10000, False, seed=1)
d, f = PROFILES["dna-r10-prom"]
g = sq.SignalGenerator(dict(d), load_model("dna-r10-prom", 9, False)[0], 9, flags=f, seed=1)
for _ in range(3):
    r = g.gen_batch_raw(bases, off, want=WANT_SVB)
print("samples", r.total_samples, "svb bytes", int(r.svb_off[r.n_reads]), "bytes/sample", int(r.svb_off[r.n_reads]) / r.total_samples)
g.close()
