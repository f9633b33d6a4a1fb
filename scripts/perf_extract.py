#!/usr/bin/env python3
"""read extraction kernels on one batch of 8192 reads by coordinates (profiling aid; run under ncu --metrics gpu__time_duration.sum)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import squigulator_b200 as sq
from squigulator_b200.api import PROFILES, WANT_SVB
from bench import synth_reads, load_model
bases, off, genome, coords = synth_reads(8192, 10000, False, seed=1, with_coords=True)
d, f = PROFILES["dna-r10-prom"]
g = sq.SignalGenerator(dict(d), load_model("dna-r10-prom", 9, False)[0], 9, flags=f, seed=1)
g.load_genome([genome.tobytes()])
for _ in range(3):
    t = g.submit_coords(np.ascontiguousarray(coords), first_read_index=0, want=WANT_SVB)
    res = g.wait(t)
    n = int(res.total_samples)
    g.release(t)
print("reads", len(coords), "bases", int(off[-1]), "samples", n)
g.close()
