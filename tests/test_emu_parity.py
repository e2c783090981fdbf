"""The CUDA kernels, compiled for the CPU by the test-only fiber emulator (tests/emu), against the
oracle.  This is how the kernels are debugged in a container without a GPU; the GPU tests
(test_gpu_parity.py) run the same scenarios through the real library on a B200."""
import numpy as np
import pytest

import parity_cases as pc


def test_reach_table(oracle, emu, synth):
    o, d = pc.build_pair(oracle, emu, synth, 16, 20, 10, hist_cap=96)
    pc.check_reach(o, d)


def test_reach_table_ragged(oracle, emu, synth):
    o, d = pc.build_pair(oracle, emu, synth, 16, 20, 10, hist_cap=96, graph_edit=pc.sparse_edges_graph)
    pc.check_reach(o, d)


def test_reach_table_built_on_device(oracle, emu, synth, capfd, monkeypatch):
    """DSIM_OPT_REACH_ON_DEVICE: the label-correcting device build gives the oracle's Dijkstra table bit for bit (regular,
    ragged and long-range graphs), and its rows and bucket indices drive the step like the host-built ones."""
    monkeypatch.setenv("DSIM_DEBUG", "1")
    for edit in (None, pc.sparse_edges_graph, pc.long_range_graph(16)):
        capfd.readouterr()
        o, d = pc.build_pair(oracle, emu, synth, 16, 20, 10, hist_cap=96, graph_edit=edit, flags=pc.D.OPT_REACH_ON_DEVICE)
        assert "reach table filled on the device" in capfd.readouterr().err, "the host builder must not have been used"
        pc.check_reach(o, d)
    o, d = pc.build_pair(oracle, emu, synth, 20, 30, 150, hist_cap=96, hist_fill=40, flags=pc.D.OPT_REACH_ON_DEVICE)
    pc.run_lockstep(o, d, 6, hist_cap=96)


def test_small_grid_by_parts(oracle, emu, synth):
    o, d = pc.build_pair(oracle, emu, synth, 20, 30, 150, hist_cap=96)
    pc.run_lockstep(o, d, 12, hist_cap=96)


def test_growth_and_splits(oracle, emu, synth):
    def peaceful(p):
        p.cooperation_threshold = 65      # everyone cooperates: population grows, families split
    o, d = pc.build_pair(oracle, emu, synth, 24, 30, 120, params_edit=peaceful, hist_cap=96)
    trace = pc.run_lockstep(o, d, 45, by_parts=False, hist_cap=96, check_every=3)
    f = o.get_families(history=False, adaptation=False)
    assert (f.parent_id != pc.D.NO_PARENT).sum() > 5, "scenario must exercise splits"
    assert trace[-1] > 100


def test_dense_coresidence_spill_path(oracle, emu, synth):
    """Nodes with more families than k_patch holds in shared memory (512): k_patch_huge, a block of sixteen warps per node
    (its HBM scratch path, beyond 2048 families, is exercised by the tiny-capacity build below)."""
    o, d = pc.build_pair(oracle, emu, synth, 30, 40, 3600, n_occupied=5, hist_cap=96)
    f = o.get_families(history=False, adaptation=False)
    assert np.unique(f.location, return_counts=True)[1].max() > 512, "scenario must exceed the shared-memory path"
    pc.run_lockstep(o, d, 2, hist_cap=96)


def test_crowded_nodes_in_shared_memory(oracle, emu, synth):
    """Nodes with 33..512 families: k_patch, one block per node, working arrays in shared memory."""
    o, d = pc.build_pair(oracle, emu, synth, 30, 40, 2000, n_occupied=8, hist_cap=96)
    f = o.get_families(history=False, adaptation=False)
    c = np.unique(f.location, return_counts=True)[1]
    assert c.max() > 64 and c.max() <= 512
    pc.run_lockstep(o, d, 3, hist_cap=96)


def test_crowded_nodes_many_hostile_bands(oracle, emu, synth):
    """Crowded nodes whose families belong to 6 mutually hostile cultures, and all sizes of node with 3: several bands
    per node, newcomers fighting their way through the bands before theirs, members wiped out (lib.rs:1209-1232)."""
    o, d = pc.build_pair(oracle, emu, synth, 30, 40, 2000, n_occupied=8, hist_cap=96, families_edit=pc.hostile_cultures(6))
    pc.run_lockstep(o, d, 3, hist_cap=96)
    assert np.unique(o.get_storage()[0], return_counts=True)[1].max() >= 4, "scenario must produce several bands on one node"
    o, d = pc.build_pair(oracle, emu, synth, 20, 24, 1500, n_occupied=80, hist_cap=96, families_edit=pc.hostile_cultures(3))
    pc.run_lockstep(o, d, 3, hist_cap=96)


@pytest.mark.parametrize("n_cultures", [63, 64, 65, 70])
def test_crowded_nodes_around_the_culture_list_limit(oracle, emu, synth, n_cultures):
    """k_patch tests a newcomer against the node's DISTINCT cultures, listed in shared memory up to 64 of them; one more and
    the node takes the quadratic test against every earlier member.  Both sides of the limit, on nodes of ~250 families
    that hold every culture."""
    def round_robin(fam):
        rng = np.random.default_rng(11)
        pool = rng.integers(0, 1 << 63, n_cultures, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, n_cultures, dtype=np.uint64)
        for node in np.unique(fam.location):
            members = np.nonzero(fam.location == node)[0]
            fam.culture[members] = pool[np.arange(len(members)) % n_cultures]
    o, d = pc.build_pair(oracle, emu, synth, 30, 40, 2000, n_occupied=8, hist_cap=96, families_edit=round_robin)
    f = o.get_families(history=False, adaptation=False)
    top = np.bincount(f.location).argmax()
    assert len(np.unique(f.culture[f.location == top])) == n_cultures, "the largest node must hold exactly that many cultures"
    pc.run_lockstep(o, d, 2, by_parts=False, hist_cap=96)


def _walk_stats(lib):
    import ctypes as C
    return list((C.c_ulonglong * 8).in_dll(lib, "dsim_walk_stats"))


def test_walk_shortcuts_are_taken_and_exact(oracle, emu, synth):
    """The walk of the hostile newcomers (patch_node) joins a band without hostile (band, culture) pair unscanned, passes
    a hostile band unscanned once the newcomer is dead and the band holds no dead member with stores, stops a scan where
    the newcomer died, counts dead members with stores per band, and settles the newcomers of a culture that band 0
    already holds 32 at a time.  The CPU build counts each shortcut
    (dsim_walk_stats): the scenario must take every one of them and still equal the oracle bit for bit.  Families with
    stores of -0.0 are in it: for them `stores + 0.0` is not a no-op and the shortcuts must stand back."""
    def edit(fam):
        pc.hostile_cultures(6)(fam)
        fam.stored[::7] = -0.0
    before = _walk_stats(emu)
    o, d = pc.build_pair(oracle, emu, synth, 30, 40, 2000, n_occupied=8, hist_cap=96, families_edit=edit)
    pc.run_lockstep(o, d, 4, hist_cap=96)
    joined, passed, scanned, took_stores, stopped, overflow, died, homed = [b - a for a, b in zip(before, _walk_stats(emu))]
    assert joined > 50 and passed > 50 and scanned > 50 and stopped > 10 and died > 50, (joined, passed, scanned, stopped, died)
    assert homed > 50, "newcomers of a culture band 0 holds must have been settled a round at a time"
    assert took_stores > 0, "a dead member's stores must have been taken by a later newcomer"
    assert overflow == 0


def test_cultures_hostile_to_themselves(oracle, emu, synth):
    """cooperation_threshold 0, exclusive: nobody cooperates, not even with its own culture, so every member is a band
    of its own and a culture sits in as many bands as it has members: the (band, culture) list of the walk outgrows its
    64 pairs on the crowded nodes and the walk falls back to scanning every band."""
    def edit(p):
        p.cooperation_threshold = 0
        p.cooperation_inclusive = 0
    before = _walk_stats(emu)[5]
    o, d = pc.build_pair(oracle, emu, synth, 30, 40, 1600, n_occupied=8, hist_cap=96, params_edit=edit, families_edit=pc.hostile_cultures(5))
    pc.run_lockstep(o, d, 2, hist_cap=96)
    assert _walk_stats(emu)[5] > before, "the pair list must have overflowed"


@pytest.mark.parametrize("bits,threshold,inclusive", [(5, 2, 0), (6, 3, 0), (4, 1, 0), (5, 1, 1), (3, 2, 1)])
def test_partially_cooperating_cultures(oracle, emu, synth, bits, threshold, inclusive):
    """Cultures of a few bits with a threshold inside their range: cooperation is not transitive, bands are cliques of
    cultures close to one another, band 0 holds several cultures and turns hostile to some newcomers and not to others
    — the (band, culture) list of the walk decides joins and passes on more than "same culture or not"."""
    def cultures(fam):
        fam.culture[:] = np.random.default_rng(bits * 10 + threshold).integers(0, 1 << bits, len(fam.culture)).astype(np.uint64)
    def params(p):
        p.cooperation_threshold = threshold
        p.cooperation_inclusive = inclusive
    before = _walk_stats(emu)
    o, d = pc.build_pair(oracle, emu, synth, 30, 40, 2000, n_occupied=8, hist_cap=96, families_edit=cultures, params_edit=params)
    pc.run_lockstep(o, d, 3, by_parts=False, hist_cap=96)
    o, d = pc.build_pair(oracle, emu, synth, 20, 24, 1500, n_occupied=60, hist_cap=96, families_edit=cultures, params_edit=params)
    pc.run_lockstep(o, d, 3, by_parts=False, hist_cap=96)
    joined, passed, scanned = [b - a for a, b in zip(before, _walk_stats(emu))][:3]
    assert joined > 0 and scanned > 0, (joined, passed, scanned)


def test_mid_sized_nodes(oracle, emu, synth):
    """Nodes with 9..32 families (k_patch_mid: a warp per node, lane per family), hostile neighbours included."""
    o, d = pc.build_pair(oracle, emu, synth, 20, 24, 1500, n_occupied=80, hist_cap=96)
    f = o.get_families(history=False, adaptation=False)
    counts = np.unique(f.location, return_counts=True)[1]
    assert ((counts > 8) & (counts <= 32)).sum() > 20, "scenario must exercise the 9..32 range"
    pc.run_lockstep(o, d, 5, hist_cap=96)


def test_prefilled_history_exact_capacity(oracle, emu, synth):
    # default history capacity (every entry the memory loop can sample); oracle history is unbounded
    o, d = pc.build_pair(oracle, emu, synth, 24, 30, 60, hist_fill=700)
    cap = d.get_families().hist_off[1] - d.get_families().hist_off[0]
    assert cap >= 625
    pc.run_lockstep(o, d, 4, hist_cap=int(cap))


def test_two_ecoregions_per_patch(oracle, emu, synth):
    o, d = pc.build_pair(oracle, emu, synth, 20, 24, 200, hist_cap=96, graph_edit=pc.two_ecoregion_graph)
    pc.run_lockstep(o, d, 8, hist_cap=96)


def test_ragged_graph(oracle, emu, synth):
    o, d = pc.build_pair(oracle, emu, synth, 20, 24, 200, hist_cap=96, graph_edit=pc.sparse_edges_graph)
    pc.run_lockstep(o, d, 8, hist_cap=96)


def test_many_sensed_patches(oracle, emu, synth):
    """2-hop lists longer than 64 entries: several sensed passes per family in k_move."""
    o, d = pc.build_pair(oracle, emu, synth, 24, 30, 200, hist_fill=30, hist_cap=96, graph_edit=pc.long_range_graph(24))
    near_off = o.get_reach()[0]
    assert np.diff(near_off.astype(np.int64)).max() > 70, "scenario must exceed one 64-entry pass"
    pc.run_lockstep(o, d, 6, hist_cap=96)


def test_slow_memory_decay(oracle, emu, synth):
    """A short season (1/24 year) makes the memory decay slowly: ~70 remembered patches are kept per decision
    (several 32-entry passes) and the memory loop reaches back 2500 entries (several 1024-entry chunks)."""
    def edit(p):
        p.season_length_in_years = 1.0 / 24.0
    o, d = pc.build_pair(oracle, emu, synth, 20, 24, 40, hist_fill=2600, params_edit=edit)
    f = d.get_families()
    cap = int(f.hist_off[1] - f.hist_off[0])
    assert cap >= 2400
    pc.run_lockstep(o, d, 3, hist_cap=cap)


def test_children_with_long_histories(oracle, emu, synth):
    """A short season makes time_to_grow_adult 361 seasons (lib.rs:1878): a child inherits up to 361 history entries,
    several 96-entry chunks of k_children, and samples its remembered patches from all of them."""
    def edit(p):
        p.season_length_in_years = 1.0 / 24.0
    o, d = pc.build_pair(oracle, emu, synth, 20, 24, 600, hist_fill=300, params_edit=edit, hist_cap=420)
    pc.run_lockstep(o, d, 1, hist_cap=420)
    f = o.get_families(adaptation=False)
    born = np.flatnonzero(f.parent_id != pc.D.NO_PARENT)
    assert len(born) > 3, "scenario must produce children"
    assert max(int(f.hist_off[i + 1] - f.hist_off[i]) for i in born) > 200, "children must inherit long histories"
    pc.run_lockstep(o, d, 3, hist_cap=420)


def test_parameter_variants(oracle, emu, synth):
    def edit(p):
        p.cooperation_inclusive = 1
        p.cooperation_threshold = 3
        p.maximum_resources_one_adult_can_harvest = 0.25   # plotting/plot.py:259 treats this as typical
        p.fight_deadliness = 1 << 30
        p.culture_mutation_rate = 0.2                      # exercise the mutation clock hitting zero
        p.minimum_adaptation = 0.3
        p.resource_recovery_per_season = 0.05
    o, d = pc.build_pair(oracle, emu, synth, 20, 24, 250, params_edit=edit, hist_cap=96)
    pc.run_lockstep(o, d, 10, hist_cap=96)


def test_extinction_and_empty(oracle, emu, synth):
    def starve(p):
        p.maximum_resources_one_adult_can_harvest = 1e-6
    o, d = pc.build_pair(oracle, emu, synth, 12, 12, 40, params_edit=starve, hist_cap=96)
    trace = pc.run_lockstep(o, d, 40, by_parts=False, hist_cap=96, check_every=5)
    assert trace[-1] == 0, "everyone starves (extinction check of src/cli.rs:138)"
    pc.run_lockstep(o, d, 2, by_parts=False, hist_cap=96)   # stepping an empty world is a no-op


@pytest.mark.parametrize("flags", [pc.D.OPT_PATCH_WARP_ONLY])
def test_kernel_variants(oracle, emu, synth, flags):
    """k_patch (warp per node) for every node."""
    o, d = pc.build_pair(oracle, emu, synth, 24, 30, 400, hist_fill=40, hist_cap=96, flags=flags)
    pc.run_lockstep(o, d, 6, hist_cap=96)

    def peaceful(p):
        p.cooperation_threshold = 65
    o, d = pc.build_pair(oracle, emu, synth, 24, 30, 120, params_edit=peaceful, hist_cap=96, flags=flags)
    pc.run_lockstep(o, d, 40, by_parts=False, hist_cap=96, check_every=4)


def test_move_rounds_with_tiny_capacities():
    """k_move compiled with 32-entry history chunks, a 1-entry queue of undecided draws and 2 families per warp tile
    (tests/emu/libdsim_emu_small.so): the history takes several chunks, the queue overflows.  Runs in its own interpreter: two builds of the same sources do not share a process."""
    import os
    import subprocess
    import sys
    code = """
import sys
sys.path.insert(0, %r)
import helpers, parity_cases as pc
oracle, synth, emu_small = helpers.load_oracle(), helpers.load_synth(), helpers.load_emu_small()
o, d = pc.build_pair(oracle, emu_small, synth, 24, 30, 300, hist_fill=200, hist_cap=256)
pc.run_lockstep(o, d, 5, hist_cap=256)
o, d = pc.build_pair(oracle, emu_small, synth, 20, 24, 200, hist_fill=40, hist_cap=96, graph_edit=pc.sparse_edges_graph)
pc.run_lockstep(o, d, 6, hist_cap=96)
def peaceful(p):
    p.cooperation_threshold = 65
o, d = pc.build_pair(oracle, emu_small, synth, 24, 30, 120, params_edit=peaceful, hist_cap=96)
pc.run_lockstep(o, d, 30, by_parts=False, hist_cap=96, check_every=5)
# nodes of ~720 families: beyond this build's shared-memory capacity of k_patch_huge (600): its HBM scratch path
o, d = pc.build_pair(oracle, emu_small, synth, 30, 40, 3600, n_occupied=5, hist_cap=96)
pc.run_lockstep(o, d, 2, hist_cap=96)
# the same with four mutually hostile cultures: the walk of the newcomers over working arrays in HBM scratch
o, d = pc.build_pair(oracle, emu_small, synth, 30, 40, 3600, n_occupied=5, hist_cap=96, families_edit=pc.hostile_cultures(4))
pc.run_lockstep(o, d, 2, hist_cap=96)
print("tiny capacities ok")
""" % os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "tiny capacities ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_more_families_than_one_scan_pass(oracle, emu, synth):
    """70 000 families: the count scan at the end of k_lifecycle handles several per-256 blocks per thread, ids and
    positions of survivors and children come out of multi-level prefixes (lib.rs:541, 551)."""
    o, d = pc.build_pair(oracle, emu, synth, 120, 160, 70000, hist_cap=32, hist_fill=4)
    pc.run_lockstep(o, d, 2, hist_cap=32)


@pytest.mark.parametrize("case", range(10))
def test_random_scenarios(oracle, emu, synth, case):
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
    o, d = pc.build_pair(oracle, emu, synth, W, H, n_fam, seed=case + 1, sim_seed=100 + case, n_occupied=n_occ, hist_fill=hist_fill,
                         params_edit=edit, hist_cap=160, graph_edit=graph_edit, families_edit=fam_edit, flags=flags)
    pc.run_lockstep(o, d, int(rng.integers(3, 9)), hist_cap=160)
