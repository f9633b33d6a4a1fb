#!/usr/bin/env python3
"""Kernel-time matrix over the signal-kernel variants (profiling aid; run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import squigulator_b200 as sq
from squigulator_b200.api import PROFILES
from bench import synth_reads, synth_model

import json
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] * 1e9
except Exception:
    PEAK = 6650e9   # B200_PROFILING.md fallback
bases, off = synth_reads(n_reads, 10000, False, seed=1)
rows = []
for prof, k in (("dna-r10-prom", 9), ("dna-r9-prom", 6)):
    for name, flags in (("noisy", 0), ("ideal_time", sq.SQ_IDEAL_TIME), ("ideal_amp", sq.SQ_IDEAL_AMP), ("ideal", sq.SQ_IDEAL)):
        d, f = PROFILES[prof]
        g = sq.SignalGenerator(dict(d), synth_model(4 ** k), k, flags=f | flags, seed=1)
        b = g.dev_batch(bases, off)
        g.dev_batch_run(b, 3)
        t, tk = g.dev_batch_run(b, 10)
        info = g.dev_batch_info(b)
        print(f"{prof:14s} {name:11s} samples={info['samples']/1e9:.3f}G step={t/10:.3f} ms kernel={tk/10:.3f} ms "
              f"-> {info['samples']/(tk/10*1e-3)/1e9:.0f} Gsamples/s ({2.08*info['samples']/(tk/10*1e-3)/PEAK*100:.1f}% of HBM roofline)")
        g.dev_batch_destroy(b)
        g.close()
