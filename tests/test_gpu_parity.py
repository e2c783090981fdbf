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
    o, d = pc.build_pair(oracle, gpu_lib, synth, 30, 40, 2500, n_occupied=6, hist_cap=96)
    pc.run_lockstep(o, d, 3, hist_cap=96)


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
