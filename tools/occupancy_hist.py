#!/usr/bin/env python
"""Families per occupied node after a few steps of a bench configuration (GPU box). usage: occupancy_hist.py c2|c3 steps"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "dispersal-simulation_b200", "python")); sys.path.insert(0, REPO)
import numpy as np, ctypes as C
import dsim_b200 as D, bench
cfg = bench.CONFIGS[sys.argv[1]]; steps = int(sys.argv[2])
synth = D.load_synth(); lib = D.load()
g, res, mx, fam = bench.build_workload(cfg, 1, synth)
opt = D.Options(); lib.dsim_default_options(C.byref(opt))
sim = D.Sim(lib, g, seed=7, options=opt); sim.set_patches(res, mx); sim.set_families(fam)
sim.step(steps, allow_model_errors=True)
f = sim.get_families(history=False, adaptation=False)
cnt = np.bincount(f.location.astype(np.int64))
cnt = cnt[cnt > 0]
h = np.bincount(cnt)
print(sys.argv[1], "families", len(f.location), "occupied", len(cnt), "max", cnt.max())
cum = 0
for k in range(1, len(h)):
    if h[k]: print(f"  n={k}: nodes {h[k]} families {k*h[k]}")
