#!/usr/bin/env python
"""GPU box: the reach table of a bench configuration built on the device vs on the host — build times, table sizes,
and the states after a few steps from the same start (must be identical).  usage: reach_device_check.py [c2] [steps]"""
import ctypes as C, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "dispersal-simulation_b200", "python")); sys.path.insert(0, REPO)
import numpy as np
import dsim_b200 as D, bench
cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c2"]; steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
synth = D.load_synth(); lib = D.load()
g, res, mx, fam = bench.build_workload(cfg, 1, synth)
out = []
for flags in (D.OPT_REACH_ON_DEVICE, 0):
    opt = D.Options(); lib.dsim_default_options(C.byref(opt)); opt.flags = flags
    t0 = time.perf_counter()
    sim = D.Sim(lib, g, seed=7, options=opt)
    dt = time.perf_counter() - t0
    sim.set_patches(res, mx); sim.set_families(fam)
    st = sim.step(steps, allow_model_errors=True)
    f = sim.get_families(history=False, adaptation=False)
    r, _ = sim.get_patches(maximum=False)
    print(f"flags {flags}: handle + reach table {dt:.2f} s; reach counts {sim.reach_counts() if hasattr(sim, 'reach_counts') else ''}; "
          f"after {steps} steps: {st.live_families} families, errors {st.model_errors}", flush=True)
    out.append((f, r))
    sim.close()
same = all((getattr(out[0][0], n) == getattr(out[1][0], n)).all() for n, _ in D._FAM_FIELDS) and (out[0][1].view(np.uint64) == out[1][1].view(np.uint64)).all()
print("IDENTICAL" if same else "DIFFERENT")
sys.exit(0 if same else 1)
