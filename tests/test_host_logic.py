"""Host side of the C ABI (state conversion, growth, error behaviour), exercised through the CPU
debugging build of the CUDA sources (tests/emu) so that it runs without a GPU."""
import ctypes as C

import numpy as np
import pytest

import helpers
import parity_cases as pc
from helpers import D


def _sim(emu, synth, W=12, H=14, n=60, hist_cap=64, cap=0, **kw):
    g, res, mx = D.synth_grid(W, H, 1, lib=synth)
    opt = D.Options()
    emu.dsim_default_options(C.byref(opt))
    opt.history_capacity = hist_cap
    opt.flags |= D.OPT_SHORT_HISTORY
    opt.family_capacity = cap
    s = D.Sim(emu, g, seed=5, options=opt)
    s.set_patches(res, mx)
    return s, g, res, mx


def test_family_round_trip(emu, synth):
    s, g, res, mx = _sim(emu, synth)
    n = 37
    rng = np.random.default_rng(3)
    f = D.synth_families(12, 14, n, 2, hist_fill=5, lib=synth)
    f.id = (np.arange(n, dtype=np.uint64) * 3 + 10)
    f.parent_id[::2] = 7
    f.birth_index[:] = rng.integers(0, 9, n)
    f.has_mutation_clock[:] = rng.integers(0, 2, n)
    f.till_mutation[:] = rng.integers(0, 500, n) * f.has_mutation_clock
    f.n_offspring[:] = rng.integers(0, 5, n)
    f.last_harvest[:] = rng.random(n)
    # ragged histories and sparse adaptation
    lens = rng.integers(0, 9, n)
    f.hist_off = np.zeros(n + 1, np.uint64); f.hist_off[1:] = np.cumsum(lens)
    f.hist_patch = rng.integers(0, g.n_nodes, int(lens.sum())).astype(np.uint32)
    f.hist_harvest = rng.random(int(lens.sum()))
    alens = rng.integers(0, 4, n)
    f.adapt_off = np.zeros(n + 1, np.uint64); f.adapt_off[1:] = np.cumsum(alens)
    f.adapt_eco = np.concatenate([np.sort(rng.choice(304, k, replace=False)) for k in alens] + [np.zeros(0, int)]).astype(np.uint32)
    f.adapt_val = 0.5 + 0.5 * rng.random(int(alens.sum()))
    s.set_families(f)
    out = s.get_families()
    assert not helpers.compare_families(f, out)
    assert s.family_counts() == (n, int(lens.sum()), int(alens.sum()))
    # counts and columns that are not asked for are not computed; the partial snapshots agree with the full one
    s.set_families(f)
    assert s.family_counts(history=False, adaptation=False) == (n, 0, 0)
    core = s.get_families(history=False, adaptation=False)
    assert core.hist_off is None and core.adapt_off is None
    for name, _ in D._FAM_FIELDS:
        assert (getattr(core, name) == getattr(f, name)).all(), name
    assert s.family_counts(adaptation=False) == (n, int(lens.sum()), 0)
    assert not helpers.compare_families(f, s.get_families())  # adaptation fetched after the record columns


def test_history_longer_than_ring_keeps_newest(emu, synth):
    s, g, *_ = _sim(emu, synth, hist_cap=8)
    f = D.synth_families(12, 14, 5, 2, hist_fill=20, lib=synth)
    s.set_families(f)
    out = s.get_families()
    for i in range(5):
        pa, ha = f.history_of(i)
        pb, hb = out.history_of(i)
        assert (pa[-8:] == pb).all() and (ha[-8:] == hb).all()


def test_patches_time_storage_round_trip(emu, synth):
    s, g, res, mx = _sim(emu, synth)
    r2, m2 = s.get_patches()
    assert (r2 == res).all() and (m2 == mx).all()
    s.set_time(41)
    assert s.get_time() == 41
    f = D.synth_families(12, 14, 10, 2, lib=synth)
    s.set_families(f)
    assert s.get_time() == 41
    node = np.array([5, 5, 9, 5], np.uint32)
    cult = np.array([3, 1, 7, 3], np.uint64)
    size = np.array([4, 2, 6, 9], np.uint64)     # the second (5, 3) overwrites the first: HashMap::insert
    s.set_storage(node, cult, size)
    n, c, z = s.get_storage()
    assert n.tolist() == [5, 5, 9] and c.tolist() == [1, 3, 7] and z.tolist() == [2, 9, 6]


def test_capacity_grows_between_steps(oracle, emu, synth):
    def peaceful(p):
        p.cooperation_threshold = 65
    g, res, mx = D.synth_grid(24, 30, 1, lib=synth)
    fam = D.synth_families(24, 30, 120, 1, lib=synth)
    p = helpers.default_params(oracle)
    peaceful(p)
    opt = D.Options()
    emu.dsim_default_options(C.byref(opt))
    opt.history_capacity = 96
    opt.flags |= D.OPT_SHORT_HISTORY
    opt.family_capacity = 128            # will have to grow: the population passes 180
    d = D.Sim(emu, g, params=p, seed=7, options=opt)
    o = helpers.oracle_sim(oracle, g, params=p, seed=7)
    for s in (o, d):
        s.set_patches(res, mx)
        s.set_families(fam)
    for _ in range(12):
        so, sd = o.step(5), d.step(5)
        assert so.as_dict() == sd.as_dict()
    assert sd.live_families > 128
    assert not helpers.compare_sims(o, d, hist_cap=96)


def test_errors_are_status_codes(emu, synth):
    s, g, *_ = _sim(emu, synth)
    f = D.synth_families(12, 14, 4, 2, lib=synth)
    f.location[2] = g.n_nodes + 3
    with pytest.raises(D.DsimError) as e:
        s.set_families(f)
    assert e.value.code == -1
    f = D.synth_families(12, 14, 4, 2, lib=synth)
    f.id[:] = [4, 3, 2, 1]
    with pytest.raises(D.DsimError):
        s.set_families(f)
    f = D.synth_families(12, 14, 4, 2, lib=synth)
    s.set_families(f)
    with pytest.raises(D.DsimError):
        s.step_part(2)                    # part 2 without part 1
    s.step_part(1)
    with pytest.raises(D.DsimError):
        s.step(1)                         # full step between the parts
    s.step_part(2)
    bad = D.synth_grid(4, 4, 1, lib=synth)[0]
    bad.eco_id[3] = 400                   # ecoregion id out of range (src/ecology.rs:137)
    with pytest.raises(D.DsimError):
        D.Sim(emu, bad, seed=1)


def test_empty_population(emu, synth):
    s, g, *_ = _sim(emu, synth)
    s.set_families(D.Families.empty(0))
    st = s.step(3)
    assert st.live_families == 0 and st.t == 3 and st.family_updates == 0
    assert s.get_families().n == 0


def test_stats_accumulate_per_call(oracle, emu, synth):
    o, d = pc.build_pair(oracle, emu, synth, 20, 24, 200, hist_cap=64)
    for k in (1, 3, 7):
        so, sd = o.step(k), d.step(k)
        assert so.as_dict() == sd.as_dict()


def test_step_timed_counts_only_its_own_steps(oracle, emu, synth):
    """dsim_step_timed reports the births, deaths and family-updates of exactly the steps it ran, whatever ran before
    it (bench.py warms up with dsim_step and then times: the warm-up's updates must not be in `value`)."""
    o, d = pc.build_pair(oracle, emu, synth, 20, 24, 200, hist_cap=64)
    o.step(5)
    d.step(5, want_stats=False)                       # warm-up without a read-back, as bench.py enqueues it
    so = o.step(20)
    sd, ms = d.step_timed(20)
    assert ms >= 0.0
    assert (sd.family_updates, sd.births, sd.deaths) == (so.family_updates, so.births, so.deaths)
    assert sd.live_families == so.live_families and sd.t == so.t == 25
    so2 = o.step(3)                                   # and again, after a stats-carrying call
    sd2, _ = d.step_timed(3)
    assert (sd2.family_updates, sd2.births, sd2.deaths) == (so2.family_updates, so2.births, so2.deaths)


def test_get_families_into_caller_buffers(emu, synth):
    """get_families(out=): the core columns land in caller-owned arrays (a checkpoint writer's reusable buffers) and equal
    a fresh download; too small a buffer and a request for histories are refused."""
    s, g, *_ = _sim(emu, synth)
    s.set_families(D.synth_families(12, 14, 50, 2, hist_fill=3, lib=synth))
    s.step(3, allow_model_errors=True)
    a = s.get_families(history=False, adaptation=False)
    buf = D.Families.empty(200)
    b = s.get_families(history=False, adaptation=False, out=buf)
    assert b.n == a.n and len(b.id) == a.n
    for name, _ in D._FAM_FIELDS:
        assert (getattr(a, name) == getattr(b, name)).all(), name
        assert np.shares_memory(getattr(b, name), getattr(buf, name))
    with pytest.raises(ValueError):
        s.get_families(history=False, adaptation=False, out=D.Families.empty(max(1, a.n - 1)))
    with pytest.raises(ValueError):
        s.get_families(history=True, adaptation=False, out=buf)


def test_short_history_ring_needs_the_flag(emu, synth):
    """A history ring shorter than what the memory loop of lib.rs:854-866 can sample changes decisions silently once a
    history outgrows it: dsim_create refuses it unless DSIM_OPT_SHORT_HISTORY says the caller knows."""
    g, _, _ = D.synth_grid(8, 8, 1, lib=synth)
    opt = D.Options()
    emu.dsim_default_options(C.byref(opt))
    opt.history_capacity = 96
    with pytest.raises(D.DsimError) as e:
        D.Sim(emu, g, seed=1, options=opt)
    assert e.value.code == -1
    opt.flags |= D.OPT_SHORT_HISTORY
    D.Sim(emu, g, seed=1, options=opt).close()
    opt.flags = 0
    opt.history_capacity = 0              # exact
    D.Sim(emu, g, seed=1, options=opt).close()
