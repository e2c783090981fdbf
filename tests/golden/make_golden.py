#!/usr/bin/env python
"""Generate the golden trajectories of tests/golden/*.npz from the CPU oracle.

The reference ships no golden vectors and cannot be run here, so these fixtures pin the
ORACLE against regressions and give the GPU tests inputs and outputs that do not depend on rebuilding
the oracle on the GPU box.  Run from the repository root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers  # noqa: E402
from helpers import D  # noqa: E402

CASES = {
    # name: (W, H, families, n_occupied, hist_fill, steps, params edits)
    "default_params": (20, 24, 160, 0, 0, 25, {}),
    "mixed_hostility": (30, 36, 300, 0, 0, 30, {"cooperation_threshold": 30}),
    "peaceful_growth": (24, 30, 120, 0, 0, 40, {"cooperation_threshold": 65}),
    "dense_history": (16, 20, 600, 5, 30, 6, {"maximum_resources_one_adult_can_harvest": 0.25}),
    # histories longer than the memory loop reaches (625 entries at season 1/6): the head of the memory draws AND the
    # skip-sampled tail (canonical stream v2)
    "long_memory": (20, 24, 60, 0, 700, 12, {}),
}


HIST_KEEP = 628   # the default history ring of the device path at season 1/6: every entry the memory loop can sample (625), rounded up to 4


def trim_histories(f, keep):
    """The newest `keep` entries of every history: the oracle keeps the reference's unbounded vector (lib.rs:218), the
    device path a ring of the entries the memory loop of lib.rs:854-866 can still reach."""
    lens = np.diff(f.hist_off.astype(np.int64))
    new_lens = np.minimum(lens, keep)
    off = np.zeros(len(lens) + 1, np.uint64)
    off[1:] = np.cumsum(new_lens)
    idx = np.concatenate([np.arange(int(f.hist_off[i + 1]) - int(new_lens[i]), int(f.hist_off[i + 1])) for i in range(len(lens))] + [np.zeros(0, np.int64)]).astype(np.int64)
    f.hist_patch, f.hist_harvest, f.hist_off = f.hist_patch[idx], f.hist_harvest[idx], off


def state_arrays(sim):
    f = sim.get_families()
    trim_histories(f, HIST_KEEP)
    res, _ = sim.get_patches()
    node, cult, size = sim.get_storage()
    out = {name: getattr(f, name) for name, _ in D._FAM_FIELDS}
    out.update(hist_off=f.hist_off, hist_patch=f.hist_patch, hist_harvest=f.hist_harvest, adapt_off=f.adapt_off,
               adapt_eco=f.adapt_eco, adapt_val=f.adapt_val, res=res, st_node=node, st_culture=cult, st_size=size,
               t=np.array([sim.get_time()], dtype=np.uint64))
    return out


def run_case(lib, synth, name, sim_factory=None):
    W, H, n, n_occ, hist, steps, edits = CASES[name]
    g, res, mx = D.synth_grid(W, H, 1, lib=synth)
    fam = D.synth_families(W, H, n, 1, n_occupied=n_occ, hist_fill=hist, lib=synth)
    p = helpers.default_params(lib)
    for k, v in edits.items():
        setattr(p, k, v)
    sim = sim_factory(g, p) if sim_factory else helpers.oracle_sim(lib, g, params=p, seed=11, mode=0)
    sim.set_patches(res, mx)
    sim.set_families(fam)
    trace = []
    for _ in range(steps):
        st = sim.step(1, allow_model_errors=True)
        trace.append((st.live_families, st.births, st.deaths, st.occupied_patches))
    out = state_arrays(sim)
    out["trace"] = np.array(trace, dtype=np.uint64)
    return out


if __name__ == "__main__":
    oracle = helpers.load_oracle()
    synth = helpers.load_synth()
    for name in CASES:
        out = run_case(oracle, synth, name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "families", len(out["id"]), "t", int(out["t"][0]), "bytes", os.path.getsize(os.path.join(HERE, name + ".npz")))
