"""Parity of the sm_100a kernels with the oracle, through the C ABI of include/dsim.h, on a real GPU."""
import numpy as np
import pytest

import helpers
import parity_cases as pc

pytestmark = pytest.mark.gpu


def test_reach_table(oracle, gpu_lib, synth):
    o, d = pc.build_pair(oracle, gpu_lib, synth, 16, 20, 10)
    pc.check_reach(o, d)
    o, d = pc.build_pair(oracle, gpu_lib, synth, 16, 20, 10, graph_edit=pc.sparse_edges_graph)
    pc.check_reach(o, d)


def test_reach_table_built_on_device(oracle, gpu_lib, synth, capfd, monkeypatch):
    """DSIM_OPT_REACH_ON_DEVICE: the label-correcting device build gives the oracle's Dijkstra table bit for bit (regular,
    ragged and long-range graphs), and its rows and bucket indices drive the step like the host-built ones."""
    monkeypatch.setenv("DSIM_DEBUG", "1")
    for edit in (None, pc.sparse_edges_graph, pc.long_range_graph(16)):
        capfd.readouterr()
        o, d = pc.build_pair(oracle, gpu_lib, synth, 16, 20, 10, hist_cap=96, graph_edit=edit, flags=pc.D.OPT_REACH_ON_DEVICE)
        assert "reach table filled on the device" in capfd.readouterr().err, "the host builder must not have been used"
        pc.check_reach(o, d)
    o, d = pc.build_pair(oracle, gpu_lib, synth, 20, 30, 150, hist_cap=96, hist_fill=40, flags=pc.D.OPT_REACH_ON_DEVICE)
    pc.run_lockstep(o, d, 6, hist_cap=96)


def test_small_grid_by_parts(oracle, gpu_lib, synth):
    o, d = pc.build_pair(oracle, gpu_lib, synth, 20, 30, 150, hist_cap=96)
    pc.run_lockstep(o, d, 12, hist_cap=96)


@pytest.mark.parametrize("use_graphs", [0, 1])
def test_growth_and_splits(oracle, gpu_lib, synth, use_graphs):
    def peaceful(p):
        p.cooperation_threshold = 65
    o, d = pc.build_pair(oracle, gpu_lib, synth, 24, 30, 120, params_edit=peaceful, hist_cap=96, use_graphs=use_graphs)
    trace = pc.run_lockstep(o, d, 45, by_parts=False, hist_cap=96, check_every=3)
    f = o.get_families(history=False, adaptation=False)
    assert (f.parent_id != pc.D.NO_PARENT).sum() > 5
    assert trace[-1] > 100


def test_dense_coresidence_spill_path(oracle, gpu_lib, synth):
    """Nodes with more families than k_patch holds in shared memory (512): the HBM scratch path."""
    o, d = pc.build_pair(oracle, gpu_lib, synth, 30, 40, 3600, n_occupied=5, hist_cap=96)
    f = o.get_families(history=False, adaptation=False)
    assert np.unique(f.location, return_counts=True)[1].max() > 512, "scenario must exceed the shared-memory path"
    pc.run_lockstep(o, d, 2, hist_cap=96)


def test_crowded_nodes_in_shared_memory(oracle, gpu_lib, synth):
    """Nodes with 33..512 families: k_patch, one block per node, working arrays in shared memory."""
    o, d = pc.build_pair(oracle, gpu_lib, synth, 30, 40, 2000, n_occupied=8, hist_cap=96)
    f = o.get_families(history=False, adaptation=False)
    c = np.unique(f.location, return_counts=True)[1]
    assert c.max() > 64 and c.max() <= 512
    pc.run_lockstep(o, d, 3, hist_cap=96)


def test_crowded_nodes_many_hostile_bands(oracle, gpu_lib, synth):
    """Crowded nodes whose families belong to 6 mutually hostile cultures, and all sizes of node with 3: several bands
    per node, newcomers fighting their way through the bands before theirs, members wiped out (lib.rs:1209-1232)."""
    o, d = pc.build_pair(oracle, gpu_lib, synth, 30, 40, 2000, n_occupied=8, hist_cap=96, families_edit=pc.hostile_cultures(6))
    pc.run_lockstep(o, d, 3, hist_cap=96)
    assert np.unique(o.get_storage()[0], return_counts=True)[1].max() >= 4, "scenario must produce several bands on one node"
    o, d = pc.build_pair(oracle, gpu_lib, synth, 20, 24, 1500, n_occupied=80, hist_cap=96, families_edit=pc.hostile_cultures(3))
    pc.run_lockstep(o, d, 3, hist_cap=96)


def test_mid_sized_nodes(oracle, gpu_lib, synth):
    """Nodes with 9..32 families (k_patch_mid: a warp per node, lane per family), hostile neighbours included."""
    o, d = pc.build_pair(oracle, gpu_lib, synth, 20, 24, 1500, n_occupied=80, hist_cap=96)
    f = o.get_families(history=False, adaptation=False)
    counts = np.unique(f.location, return_counts=True)[1]
    assert ((counts > 8) & (counts <= 32)).sum() > 20, "scenario must exercise the 9..32 range"
    pc.run_lockstep(o, d, 5, hist_cap=96)


def test_prefilled_history_exact_capacity(oracle, gpu_lib, synth):
    o, d = pc.build_pair(oracle, gpu_lib, synth, 24, 30, 60, hist_fill=700)
    f = d.get_families()
    cap = int(f.hist_off[1] - f.hist_off[0])
    assert cap >= 625
    pc.run_lockstep(o, d, 4, hist_cap=cap)


def test_two_ecoregions_per_patch(oracle, gpu_lib, synth):
    o, d = pc.build_pair(oracle, gpu_lib, synth, 20, 24, 200, hist_cap=96, graph_edit=pc.two_ecoregion_graph)
    pc.run_lockstep(o, d, 8, hist_cap=96)


def test_ragged_graph(oracle, gpu_lib, synth):
    o, d = pc.build_pair(oracle, gpu_lib, synth, 20, 24, 200, hist_cap=96, graph_edit=pc.sparse_edges_graph)
    pc.run_lockstep(o, d, 8, hist_cap=96)


def test_many_sensed_patches(oracle, gpu_lib, synth):
    """2-hop lists longer than 64 entries: several sensed passes per family in k_move."""
    o, d = pc.build_pair(oracle, gpu_lib, synth, 24, 30, 200, hist_fill=30, hist_cap=96, graph_edit=pc.long_range_graph(24))
    near_off = o.get_reach()[0]
    assert np.diff(near_off.astype(np.int64)).max() > 70, "scenario must exceed one 64-entry pass"
    pc.run_lockstep(o, d, 6, hist_cap=96)


def test_slow_memory_decay(oracle, gpu_lib, synth):
    """A short season (1/24 year) makes the memory decay slowly: ~70 remembered patches are kept per decision
    (several 32-entry passes) and the memory loop reaches back 2500 entries (several 1024-entry chunks)."""
    def edit(p):
        p.season_length_in_years = 1.0 / 24.0
    o, d = pc.build_pair(oracle, gpu_lib, synth, 20, 24, 40, hist_fill=2600, params_edit=edit)
    f = d.get_families()
    cap = int(f.hist_off[1] - f.hist_off[0])
    assert cap >= 2400
    pc.run_lockstep(o, d, 3, hist_cap=cap)


def test_children_with_long_histories(oracle, gpu_lib, synth):
    """A short season makes time_to_grow_adult 361 seasons (lib.rs:1878): a child inherits up to 361 history entries,
    several 96-entry chunks of k_children, and samples its remembered patches from all of them."""
    def edit(p):
        p.season_length_in_years = 1.0 / 24.0
    o, d = pc.build_pair(oracle, gpu_lib, synth, 20, 24, 600, hist_fill=300, params_edit=edit, hist_cap=420)
    pc.run_lockstep(o, d, 1, hist_cap=420)
    f = o.get_families(adaptation=False)
    born = np.flatnonzero(f.parent_id != pc.D.NO_PARENT)
    assert len(born) > 3, "scenario must produce children"
    assert max(int(f.hist_off[i + 1] - f.hist_off[i]) for i in born) > 200, "children must inherit long histories"
    pc.run_lockstep(o, d, 3, hist_cap=420)


def test_parameter_variants(oracle, gpu_lib, synth):
    def edit(p):
        p.cooperation_inclusive = 1
        p.cooperation_threshold = 3
        p.maximum_resources_one_adult_can_harvest = 0.25
        p.fight_deadliness = 1 << 30
        p.culture_mutation_rate = 0.2
        p.minimum_adaptation = 0.3
        p.resource_recovery_per_season = 0.05
    o, d = pc.build_pair(oracle, gpu_lib, synth, 20, 24, 250, params_edit=edit, hist_cap=96)
    pc.run_lockstep(o, d, 10, hist_cap=96)


def test_extinction_and_empty(oracle, gpu_lib, synth):
    def starve(p):
        p.maximum_resources_one_adult_can_harvest = 1e-6
    o, d = pc.build_pair(oracle, gpu_lib, synth, 12, 12, 40, params_edit=starve, hist_cap=96)
    trace = pc.run_lockstep(o, d, 40, by_parts=False, hist_cap=96, check_every=5)
    assert trace[-1] == 0
    pc.run_lockstep(o, d, 2, by_parts=False, hist_cap=96)


def test_config1_reference_cpu_case(oracle, gpu_lib, synth):
    """BASELINE.json configs[0]: ~10k patches, ~1k families, 100 steps — compared at the end and every 10 steps,
    batched stepping (CUDA graph replay, growth checks) on the device side."""
    o, d = pc.build_pair(oracle, gpu_lib, synth, 100, 100, 1000, hist_cap=0)
    for k in range(10):
        so = o.step(10, allow_model_errors=True)
        sd = d.step(10, allow_model_errors=True)
        assert so.as_dict() == sd.as_dict()
        msgs = helpers.compare_sims(o, d, hist_cap=None, what=f"after {10 * (k + 1)} steps")
        assert not msgs, "\n".join(msgs)


def test_mid_size_steady_state(oracle, gpu_lib, synth):
    """A 300x450 grid with 20k families and pre-filled histories (the 625-draw regime of lib.rs:854-866)."""
    o, d = pc.build_pair(oracle, gpu_lib, synth, 300, 450, 20000, hist_fill=640)
    f = d.get_families(adaptation=False)
    cap = int(f.hist_off[1] - f.hist_off[0])
    for k in range(3):
        o.step(2, allow_model_errors=True)
        d.step(2, allow_model_errors=True)
        msgs = helpers.compare_sims(o, d, hist_cap=cap, what=f"after {2 * (k + 1)} steps")
        assert not msgs, "\n".join(msgs)


def test_no_device_fallback_is_absent(gpu_lib):
    """The binding exposes no way to run the step on the CPU."""
    assert not hasattr(helpers.D, "cpu_step")
    assert helpers.D.LIB_PATH.endswith("libdsim_b200.so")


@pytest.mark.parametrize("flags", [pc.D.OPT_PATCH_WARP_ONLY])
def test_kernel_variants(oracle, gpu_lib, synth, flags):
    """k_patch (warp per node) for every node."""
    o, d = pc.build_pair(oracle, gpu_lib, synth, 24, 30, 400, hist_fill=40, hist_cap=96, flags=flags)
    pc.run_lockstep(o, d, 6, hist_cap=96)

    def peaceful(p):
        p.cooperation_threshold = 65
    o, d = pc.build_pair(oracle, gpu_lib, synth, 24, 30, 120, params_edit=peaceful, hist_cap=96, flags=flags)
    pc.run_lockstep(o, d, 40, by_parts=False, hist_cap=96, check_every=4)


def test_full_size_invariants(gpu_lib, synth):
    """BASELINE.json configs[1] at full size (1155x1732 patches, 100k families, 625-entry histories), where the oracle
    would take minutes per step: properties that do not depend on the size.
    * determinism: two handles fed the same state agree bit for bit after the same steps (no race decides anything);
    * batching: 6 steps in one call (CUDA-graph replay) == 6 calls of one step;
    * census: the `storage` map (lib.rs:622-633) holds exactly the families' sizes, per patch;
    * the family vector stays in id order with unique ids (lib.rs:541,551: retain + append), ids of children above
      every id of the start state, parents known;
    * resources stay within [0, max] (lib.rs:2060-2063, 2100-2108)."""
    import bench
    g, res, mx, fam = bench.build_workload(bench.CONFIGS["c2"], 1, synth)

    def make():
        s = pc.D.Sim(gpu_lib, g, seed=7)
        s.set_patches(res, mx)
        s.set_families(fam)
        return s
    a, b = make(), make()
    sa = a.step(6, allow_model_errors=True)
    for _ in range(6):
        sb = b.step(1, allow_model_errors=True)
    assert sa.model_errors == 0 and sb.model_errors == 0
    assert sa.live_families == sb.live_families and sa.t == sb.t == 6
    fa, fb = a.get_families(), b.get_families()
    for name, _ in pc.D._FAM_FIELDS:
        x, y = getattr(fa, name), getattr(fb, name)
        assert x.shape == y.shape and (x.view(np.uint8) == y.view(np.uint8)).all(), name
    assert (fa.hist_off == fb.hist_off).all() and (fa.hist_patch == fb.hist_patch).all()
    assert (fa.hist_harvest.view(np.uint64) == fb.hist_harvest.view(np.uint64)).all()
    ra, ma = a.get_patches()
    rb, _ = b.get_patches()
    assert (ra.view(np.uint64) == rb.view(np.uint64)).all()
    # vector order = id order, unique; children are younger than every initial family
    assert (np.diff(fa.id.astype(np.int64)) > 0).all()
    born = fa.parent_id != pc.D.NO_PARENT
    assert born.sum() > 0 and fa.id[born].min() >= fam.n and (fa.parent_id[born] < fa.id[born]).all()
    assert (fa.size >= 0).all() and (fa.location < g.n_nodes).all()
    # census: storage' sums the sizes of the families of every patch (bands of total size 0 are dropped)
    node, cul, size = a.get_storage()
    per_patch = np.bincount(fa.location.astype(np.int64), weights=fa.size.astype(np.float64), minlength=g.n_nodes)
    from_storage = np.bincount(node.astype(np.int64), weights=size.astype(np.float64), minlength=g.n_nodes)
    assert (per_patch == from_storage).all()
    assert (ra >= 0).all() and (ra <= ma).all()


@pytest.mark.parametrize("case", range(6))
def test_random_scenarios(oracle, gpu_lib, synth, case):
    """Randomly drawn small scenarios (grid, crowding, history length, parameters, graph variants): every state
    column bit-identical to the oracle after every part of every step."""
    rng = np.random.default_rng(1000 + case)
    W, H = int(rng.integers(8, 26)), int(rng.integers(8, 30))
    n_fam = int(rng.integers(20, 500))
    n_occ = int(rng.choice([0, 0, max(1, n_fam // 40), max(1, n_fam // 8)]))
    hist_fill = int(rng.choice([0, 3, 40, 130]))
    graph_edit = [None, None, pc.sparse_edges_graph, pc.two_ecoregion_graph, pc.long_range_graph(W)][int(rng.integers(0, 5))]
    fam_edit = [None, None, pc.hostile_cultures(int(rng.integers(2, 9)), seed=case)][int(rng.integers(0, 3))]
    flags = int(rng.choice([0, 0, pc.D.OPT_PATCH_WARP_ONLY, pc.D.OPT_REACH_ON_DEVICE]))

    def edit(p):
        p.cooperation_threshold = int(rng.integers(2, 12))
        p.cooperation_inclusive = int(rng.integers(0, 2))
        p.fight_deadliness = int(rng.choice([1 << 29, 1 << 31, 3 << 30]))
        p.culture_mutation_rate = float(rng.choice([6e-3, 0.05, 0.3]))
        p.resource_recovery_per_season = float(rng.choice([0.05, 0.10, 0.2]))
        p.minimum_adaptation = float(rng.choice([0.3, 0.5]))
        if rng.random() < 0.3:
            p.maximum_resources_one_adult_can_harvest = 0.25
    o, d = pc.build_pair(oracle, gpu_lib, synth, W, H, n_fam, seed=case + 1, sim_seed=100 + case, n_occupied=n_occ, hist_fill=hist_fill,
                         params_edit=edit, hist_cap=160, graph_edit=graph_edit, families_edit=fam_edit, flags=flags)
    pc.run_lockstep(o, d, int(rng.integers(3, 9)), hist_cap=160)


def _full_size_pair(oracle, gpu_lib, synth, W, H, n_fam, n_occ, hist_fill, flags=0):
    """The named BASELINE.json configuration on both sides.  The oracle runs in its reference-shaped mode (one bounded
    Dijkstra per family per decision, lib.rs:843) with OpenMP on all host cores: 0.4 s per step at C2 on 16 cores."""
    import os
    g, res, mx = pc.D.synth_grid(W, H, 1, lib=synth)
    fam = pc.D.synth_families(W, H, n_fam, 1, n_occupied=n_occ, hist_fill=hist_fill, lib=synth)
    oracle.oracle_set_threads.argtypes = [helpers.C.c_int]
    oracle.oracle_set_threads(len(os.sched_getaffinity(0)))
    o = helpers.oracle_sim(oracle, g, seed=7, mode=0)
    opt = pc.D.Options()
    gpu_lib.dsim_default_options(opt)
    opt.flags = flags
    d = pc.D.Sim(gpu_lib, g, seed=7, options=opt)
    for s in (o, d):
        s.set_patches(res, mx)
        s.set_families(fam)
    return o, d


def _lockstep_bulk(o, d, steps, what):
    for k in range(steps):
        so, sd = o.step(1, allow_model_errors=True), d.step(1, allow_model_errors=True)
        assert so.as_dict() == sd.as_dict(), (what, k, so.as_dict(), sd.as_dict())
        msgs = helpers.compare_sims_bulk(o, d, what=f"{what} step {k}")
        assert not msgs, "\n".join(msgs)
    return so


def test_c2_full_size_vs_oracle(oracle, gpu_lib, synth):
    """BASELINE.json configs[1] at its real size — 1155x1732 patches, 100k families, 625-entry histories — three steps,
    every column of the state bit-identical to the oracle after each (625 + 3 entries still fit the 628-entry ring)."""
    o, d = _full_size_pair(oracle, gpu_lib, synth, 1155, 1732, 100_000, 0, 625)
    st = _lockstep_bulk(o, d, 3, "C2")
    assert st.live_families > 50_000 and st.model_errors == 0


def test_c3_one_step_vs_oracle(oracle, gpu_lib, synth):
    """BASELINE.json configs[2]: the same grid with 1M families; one step against the oracle.  Histories of 128 entries
    (the head of the memory loop and the start of its tail) keep the oracle's array-of-structs state at ~5 GB."""
    o, d = _full_size_pair(oracle, gpu_lib, synth, 1155, 1732, 1_000_000, 0, 128)
    st = _lockstep_bulk(o, d, 1, "C3")
    assert st.live_families > 500_000


def test_c4_shaped_dense_coresidence_vs_oracle(oracle, gpu_lib, synth):
    """BASELINE.json configs[3] in shape: dense co-residence, mean 8 families per occupied patch with a geometric tail
    (200k families on 25k of 135k patches), two steps against the oracle — the harvest-competition-bound regime."""
    o, d = _full_size_pair(oracle, gpu_lib, synth, 300, 450, 200_000, 25_000, 64)
    f = o.get_families(history=False, adaptation=False)
    c = np.unique(f.location, return_counts=True)[1]
    assert 7.0 < c.mean() < 9.0 and c.max() > 32
    _lockstep_bulk(o, d, 2, "C4-shaped")


def test_reach_table_device_vs_host_at_c2(gpu_lib, synth):
    """K0 at the size of BASELINE.json configs[1] (2.0M nodes, 36M edges, 250M row entries): the table the device kernels
    build (the default) and the host builder's are the same bytes in HBM — digests of headers, destinations, costs,
    bucket indices and 2-hop lists (dsim_reach_digest) — and the device build keeps dsim_create under 2.5 s."""
    import time
    g, _, _ = pc.D.synth_grid(1155, 1732, 1, lib=synth)
    digests, times = [], []
    for flags in (0, pc.D.OPT_REACH_ON_HOST):
        opt = pc.D.Options()
        gpu_lib.dsim_default_options(opt)
        opt.flags = flags
        t0 = time.perf_counter()
        s = pc.D.Sim(gpu_lib, g, seed=7, options=opt)
        times.append(time.perf_counter() - t0)
        digests.append(s.reach_digest())
        s.close()
    print(f"dsim_create at C2: device-built reach table {times[0]:.2f} s, host-built {times[1]:.2f} s")
    assert digests[0] == digests[1], (digests, times)
    assert all(x != 0 for x in digests[0])
    assert times[0] < 4.0, times   # 2.1 s measured; the bound leaves room for a cold first CUDA context


@pytest.mark.gpu
def test_walk_shortcuts_with_negative_zero_stores(oracle, gpu_lib, synth):
    """The walk of the hostile newcomers passes bands unscanned and settles band 0's own culture 32 at a time
    (tests/test_emu_parity.py::test_walk_shortcuts_are_taken_and_exact counts the shortcuts on the CPU build); here the
    same scenario on the device, stores of -0.0 included (for those `stores + 0.0` is not a no-op)."""
    def edit(fam):
        pc.hostile_cultures(6)(fam)
        fam.stored[::7] = -0.0
    o, d = pc.build_pair(oracle, gpu_lib, synth, 30, 40, 2000, n_occupied=8, hist_cap=96, families_edit=edit)
    pc.run_lockstep(o, d, 4, hist_cap=96)


@pytest.mark.gpu
def test_cultures_hostile_to_themselves(oracle, gpu_lib, synth):
    """cooperation_threshold 0, exclusive: every member a band of its own; the (band, culture) list of the walk
    overflows and every band is scanned."""
    def edit(p):
        p.cooperation_threshold = 0
        p.cooperation_inclusive = 0
    o, d = pc.build_pair(oracle, gpu_lib, synth, 30, 40, 1600, n_occupied=8, hist_cap=96, params_edit=edit,
                         families_edit=pc.hostile_cultures(5))
    pc.run_lockstep(o, d, 2, hist_cap=96)


@pytest.mark.gpu
def test_long_run_in_batches(oracle, gpu_lib, synth):
    """400 seasons enqueued 25 at a time (no host look at the state inside a batch: the adaptive interval between the
    library's own checks of the live count, capacity growth and the graph re-captures that follow it, the deferred
    end-of-step bookkeeping), a growing population, histories that wrap the exact ring (300 + 400 > 628 entries):
    bit-identical to the oracle after every batch."""
    n0, n1 = _long_run(oracle, gpu_lib, synth, 16)
    assert n1 > 2 * n0 + 1024, "scenario must exercise growth beyond the initial capacity"


def _long_run(oracle, lib, synth, batches, steps_per_batch=25):
    def peaceful(p):
        p.cooperation_threshold = 65
    o, d = pc.build_pair(oracle, lib, synth, 60, 80, 800, params_edit=peaceful, hist_fill=700, hist_cap=0)
    f = d.get_families()
    cap = int(f.hist_off[1] - f.hist_off[0])   # the exact ring: what the memory loop can sample (628 at the default season)
    assert cap >= 625
    n0 = o.family_counts()[0]
    for batch in range(batches):
        so = o.step(steps_per_batch, allow_model_errors=True)
        sd = d.step(steps_per_batch, allow_model_errors=True)
        assert so.as_dict() == sd.as_dict(), (batch, so.as_dict(), sd.as_dict())
        msgs = helpers.compare_sims(o, d, hist_cap=cap, what=f"after batch {batch}")
        assert not msgs, "\n".join(msgs)
    return n0, o.family_counts()[0]
