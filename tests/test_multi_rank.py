"""The band-partitioned multi-GPU step on "virtual ranks": several handles in one process, messages moved
with plain copies.  The partitioned run must be bit-identical to the single-rank run (and hence to the
oracle) for every number of ranks — every random draw is keyed by entity id, never by rank or position.
CPU: kernels under the test-only emulator.  GPU: the same with the real library on one device."""
import ctypes as C

import numpy as np
import pytest

import helpers
import parity_cases as pc
from helpers import D


def gather(sims):
    """Families of all ranks sorted by id, own patches merged, storage concatenated."""
    fams = [s.get_families() for s in sims]
    ids = np.concatenate([f.id for f in fams])
    order = np.argsort(ids, kind="stable")
    out = D.Families(len(ids))
    for name, _ in D._FAM_FIELDS:
        setattr(out, name, np.concatenate([getattr(f, name) for f in fams])[order])
    hist = [f.history_of(i) for f in fams for i in range(f.n)]
    adap = [f.adaptation_of(i) for f in fams for i in range(f.n)]
    hist = [hist[k] for k in order]
    adap = [adap[k] for k in order]
    out.hist_off = np.zeros(len(ids) + 1, np.uint64)
    out.hist_off[1:] = np.cumsum([len(h[0]) for h in hist])
    out.hist_patch = np.concatenate([h[0] for h in hist] + [np.zeros(0, np.uint32)])
    out.hist_harvest = np.concatenate([h[1] for h in hist] + [np.zeros(0)])
    out.adapt_off = np.zeros(len(ids) + 1, np.uint64)
    out.adapt_off[1:] = np.cumsum([len(a[0]) for a in adap])
    out.adapt_eco = np.concatenate([a[0] for a in adap] + [np.zeros(0, np.uint32)])
    out.adapt_val = np.concatenate([a[1] for a in adap] + [np.zeros(0)])
    res = None
    for s in sims:
        r, _ = s.get_patches()
        if res is None:
            res = np.zeros_like(r)
        lo, hi = int(s.graph.eco_off[s.mg_bounds[s.mg_rank]]), int(s.graph.eco_off[s.mg_bounds[s.mg_rank + 1]])
        res[lo:hi] = r[lo:hi]
    st = [s.get_storage() for s in sims]
    node = np.concatenate([x[0] for x in st]); cult = np.concatenate([x[1] for x in st]); size = np.concatenate([x[2] for x in st])
    o = np.lexsort((cult, node))
    return out, res, (node[o], cult[o], size[o])


def run_partitioned(lib, oracle, synth, world, W, H, n_fam, steps, hist_fill=0, params_edit=None, check_every=1, hist_cap=96):
    g, res, mx = D.synth_grid(W, H, 1, lib=synth)
    fam = D.synth_families(W, H, n_fam, 1, hist_fill=hist_fill, lib=synth)
    p = helpers.default_params(oracle)
    if params_edit:
        params_edit(p)
    ref = helpers.oracle_sim(oracle, g, params=p, seed=7)
    ref.set_patches(res, mx)
    ref.set_families(fam)
    sims = []
    for r in range(world):
        opt = D.Options()
        lib.dsim_default_options(C.byref(opt))
        opt.history_capacity = hist_cap
        opt.flags |= D.OPT_SHORT_HISTORY
        opt.family_capacity = 4 * n_fam + 1024
        s = D.Sim(lib, g, params=p, seed=7, options=opt)
        if r == 0:
            bounds = D.mg_band_bounds(fam.location, g.n_nodes, world, s.mg_halo_width())
        s.mg_configure(r, world, bounds)
        s.set_patches(res, mx)
        s.set_families(fam)
        sims.append(s)
    D.mg_halo_sync_in_process(sims)
    moved = 0
    for k in range(steps):
        before = [set(s.get_families(history=False, adaptation=False).id.tolist()) for s in sims]
        ref.step(1, allow_model_errors=True)
        D.mg_step_in_process(sims)
        after = [set(s.get_families(history=False, adaptation=False).id.tolist()) for s in sims]
        moved += sum(len(after[r] & before[q]) for r in range(world) for q in range(world) if q != r)
        if k % check_every == 0 or k == steps - 1:
            f, r_, st = gather(sims)
            msgs = helpers.compare_families(ref.get_families(), f, hist_cap=hist_cap, what=f"world {world} step {k}")
            ra, _ = ref.get_patches()
            if (ra.view(np.uint64) != r_.view(np.uint64)).any():
                msgs.append(f"world {world} step {k}: patch resources differ")
            na, ca, sa = ref.get_storage()
            if len(na) != len(st[0]) or (na != st[0]).any() or (ca != st[1]).any() or (sa != st[2]).any():
                msgs.append(f"world {world} step {k}: storage differs")
            assert not msgs, "\n".join(msgs)
            assert all(s.get_time() == ref.get_time() for s in sims)
    return moved, sims


@pytest.mark.parametrize("world", [1, 2, 3])
def test_virtual_ranks_match_single_rank(oracle, emu, synth, world):
    moved, _ = run_partitioned(emu, oracle, synth, world, 16, 60, 260, 10)
    if world > 1:
        assert moved > 0, "scenario must exercise migration between bands"


def test_virtual_ranks_growth_and_children_ids(oracle, emu, synth):
    def peaceful(p):
        p.cooperation_threshold = 65
    moved, sims = run_partitioned(emu, oracle, synth, 2, 16, 60, 150, 45, params_edit=peaceful, check_every=5)
    kids = sum(int((s.get_families(history=False, adaptation=False).parent_id != D.NO_PARENT).sum()) for s in sims)
    assert kids > 5 and moved > 0


def test_virtual_ranks_with_history(oracle, emu, synth):
    run_partitioned(emu, oracle, synth, 2, 16, 60, 120, 5, hist_fill=150, hist_cap=160)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_virtual_ranks_on_gpu(oracle, gpu_lib, synth, world):
    moved, sims = run_partitioned(gpu_lib, oracle, synth, world, 24, 120, 900, 12, hist_fill=40)
    assert moved > 0
    # the reach table of a rank holds the rows of its band and halo only (device builder: counting pass for all nodes,
    # fill pass for [own_lo - halo, own_hi + halo)): bit-identical results above, a fraction of the entries here
    full = D.Sim(gpu_lib, sims[0].graph, seed=7)
    n_full = full.reach_counts()[1]
    hw = sims[0].mg_halo_width()
    for s in sims:
        lo, hi = int(s.mg_bounds[s.mg_rank]), int(s.mg_bounds[s.mg_rank + 1])
        n_band = s.reach_counts()[1]
        assert n_band < n_full, "a band's table must be smaller than the full one"
        assert n_band <= n_full * (hi - lo + 2 * hw) / s.graph.n_nodes * 1.25 + 1
    full.close()
