#!/usr/bin/env python
"""BASELINE.json configs[4]: the gavin2017processbased model batched over a parameter sweep on one B200.

    python tools/bench_gavin.py [--members 1000] [--rows 64] [--cols 64] [--cpu-members 2]

One JSON line: sweep members per second and cells visited by `grow` per second (device time and end to end through
`gavin_run` from host buffers), next to the CPU oracle (oracle/gavin.py: pure Python, like the reference) timed on
`--cpu-members` members of the same sweep.  Synthetic precipitation and language-cap tables stand in for the
reference's inputs that are not in its repository (oracle/gavin.py).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "dispersal-simulation_b200", "python"))
sys.path.insert(0, os.path.join(REPO, "oracle"))
import dsim_b200 as D  # noqa: E402
from dsim_b200 import gavin as GB  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--members", type=int, default=1000)
    ap.add_argument("--rows", type=int, default=64)
    ap.add_argument("--cols", type=int, default=64)
    ap.add_argument("--cpu-members", type=int, default=2)
    args = ap.parse_args()
    import gavin as G  # the oracle: inputs and CPU leg only
    S, R, C = args.members, args.rows, args.cols
    P = G.synthetic_precipitation(R, C)
    rng = np.random.default_rng(4)
    alpha = 10 ** rng.uniform(-8.3, -7.8, (S, 1))      # around the published 10^-8.07, model.py:269
    beta = rng.uniform(2.5, 2.75, (S, 1))              # around 2.64, model.py:270
    popcap = GB.population_capacity(P[None, :], alpha, beta)
    gf = rng.uniform(1.05, 1.3, S)                     # around growth_factor = 1.1, language.py:7
    seed = rng.integers(1, 2 ** 62, S).astype(np.uint64)
    caps = G.synthetic_lang_caps()
    lib = D.load()
    GB.run_sweep(lib, R, C, popcap[:8], gf[:8], seed[:8], caps)   # warm-up (context, module load)
    t0 = time.perf_counter()
    out = GB.run_sweep(lib, R, C, popcap, gf, seed, caps)
    wall = time.perf_counter() - t0
    cpu = None
    if args.cpu_members > 0:
        t0 = time.perf_counter()
        for k in range(args.cpu_members):
            m = G.Member(R, C, popcap[k], gf[k], int(seed[k]), caps)
            m.run()
            assert (m.language == out.language[k]).all() and (m.population == out.population[k]).all(), "GPU != oracle"
        t_cpu = time.perf_counter() - t0
        cpu = {"value": args.cpu_members / t_cpu, "unit": "members/s", "cores": 1, "kind": "port",
               "sample": f"{args.cpu_members} members of the same sweep, oracle/gavin.py (pure Python, as the reference), {t_cpu:.1f} s; "
                         "their results equal the GPU's"}
    line = {"metric": "sweep members/s (gavin2017processbased, one B200)", "value": S / (out.device_ms * 1e-3), "unit": "members/s",
            "cell_updates_per_s": out.cell_updates / (out.device_ms * 1e-3), "device_ms": out.device_ms,
            "e2e": {"value": S / wall, "unit": "members/s", "h2d_bytes": int(popcap.nbytes + gf.nbytes + seed.nbytes + caps.nbytes),
                    "d2h_bytes": int(out.language.nbytes + out.population.nbytes + out.n_languages.nbytes)},
            "config": {"workload": f"{S} members x 2x{R}x{C} cells, synthetic precipitation and language caps",
                       "languages_per_member": float(out.n_languages.mean()), "grow_calls": out.grow_calls,
                       "cell_updates": out.cell_updates},
            "cpu_baseline": cpu, "data": "synthetic", "dtype": "f64/int32"}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
