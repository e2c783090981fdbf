import os, sys
sys.path.insert(0, "dispersal-simulation_b200/python"); sys.path.insert(0, ".")
import numpy as np, ctypes as C
import dsim_b200 as D, bench
cfg = bench.CONFIGS["c2"]
synth = D.load_synth(); lib = D.load()
g, res, mx, fam = bench.build_workload(cfg, 1, synth)
opt = D.Options(); lib.dsim_default_options(C.byref(opt))
sim = D.Sim(lib, g, seed=7, options=opt); sim.set_patches(res, mx); sim.set_families(fam)
done = 0
for steps in (1000, 2000, 5000, 10000):
    sim.step(steps - done, allow_model_errors=True); done = steps
    f = sim.get_families(history=False, adaptation=False)
    cnt = np.bincount(f.location.astype(np.int64)); cnt = cnt[cnt > 0]
    cs = np.sort(cnt)[::-1]
    ncult = []
    for node in np.argsort(np.bincount(f.location.astype(np.int64)))[::-1][:3]:
        m = f.location == node
        ncult.append((int(m.sum()), len(np.unique(f.culture[m]))))
    print(f"t={steps}: families {len(f.location)} occupied {len(cnt)} top sizes {cs[:8].tolist()} nodes>32 {(cnt>32).sum()} nodes>512 {(cnt>512).sum()} families on nodes>32 {cnt[cnt>32].sum()} (size, distinct cultures) of top3 {ncult}", flush=True)
