"""gavin2017processbased: the oracle against the reference's own functions (golden vectors made
by tests/golden/make_gavin_golden.py from supplement/gavin2017processbased/language.py), its invariants, and the
batched CUDA path against the oracle."""
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "oracle"))
import gavin as G  # noqa: E402  (test infrastructure)

GOLD = np.load(os.path.join(REPO, "tests", "golden", "gavin_functions.npz"))


def test_dhondt_matches_reference():
    for k in range(len(GOLD["seats"])):
        v = GOLD["votes"][k]
        n = int(np.sum(~np.isnan(v)))
        assert G.dhondt(float(GOLD["seats"][k]), [float(x) for x in v[:n]]) == GOLD["dhondt"][k, :n].tolist(), k


def test_distribute_matches_reference():
    for k in range(len(GOLD["growth"])):
        pc = GOLD["popcap"][k]
        n = int(np.sum(~np.isnan(pc)))
        space = [float(pc[i]) - int(GOLD["population"][k, i]) for i in range(n)]
        keys = sorted((int(GOLD["keys"][k, i]), i) for i in range(1, n))
        order = [0] + [i for _, i in keys]
        got, left = G.distribute(float(GOLD["growth"][k]), space, order)
        want = [(int(i), int(GOLD["got"][k, i])) for i in GOLD["visit"][k] if i >= 0]
        assert got == want, (k, got, want)          # values AND the insertion order of the reference's Counter
        assert left == GOLD["left"][k], k


def test_grid_topology_matches_reference():
    """`Grid.gridcell` / `Grid.neighbors` of the reference (model.py:193-262), executed from model.py's syntax tree
    (tests/golden/make_gavin_grid_golden.py): the six neighbours in the reference's order with the cell itself where
    the grid ends, and the default filter (same or no language, capacity of at least 1) on random states."""
    gold = np.load(os.path.join(REPO, "tests", "golden", "gavin_grid.npz"))
    assert G.AREA == float(gold["area"])
    for R, C in gold["shapes"]:
        R, C = int(R), int(C)
        assert (G.neighbors_table(R, C) == gold[f"all_{R}x{C}"]).all(), (R, C)
        popcap, language = gold[f"popcap_{R}x{C}"], gold[f"language_{R}x{C}"]
        m = G.Member(R, C, popcap, 1.1, 1, [100.0])
        m.language[:] = language
        off, flat = gold[f"default_off_{R}x{C}"], gold[f"default_{R}x{C}"]
        for c in range(2 * R * C):
            assert m.neighbors(c, int(language[c])) == [int(q) for q in flat[off[c]:off[c + 1]]], (R, C, c)
    # the published capacity model (model.py:265-267) evaluated the way the callers do
    p = np.array([10.0, 250.0, 1800.0])
    want = float(gold["alpha"]) * p ** float(gold["beta"]) * float(gold["area"]) / 1000000
    assert (G.population_capacity(p, float(gold["alpha"]), float(gold["beta"])) == want).all()


def test_philox_kat():
    # Random123 known-answer vector for Philox4x32-10 (same generator as the dispersal model's streams)
    assert G.philox4x32_10((0xFFFFFFFF,) * 4, (0xFFFFFFFF, 0xFFFFFFFF)) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    assert G.philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)


def small_member(seed=5, R=10, C=12, alpha=10 ** -8.07, beta=2.64, gf=1.1):
    P = G.synthetic_precipitation(R, C)
    return G.Member(R, C, G.population_capacity(P, alpha, beta), gf, seed, G.synthetic_lang_caps())


def test_oracle_invariants():
    m = small_member()
    n_lang = m.run()
    assert n_lang >= 3
    owned = m.language != 0
    assert (m.population[~owned] == 0).all() and (m.population[owned] >= 1).all()
    assert (m.popcap[owned] >= 1).all(), "only cells that can hold an individual are settled (model.py:262, 346-349)"
    assert set(np.unique(m.language[owned])) == set(range(1, n_lang + 1))
    # the run ended because no free neighbour is left (model.py:364-366)
    for c in np.flatnonzero(owned):
        for q in m.neighbors(int(c), m.language[c]):
            assert owned[q] or m.popcap[q] <= 0
    # every language is connected through its own cells
    for lid in range(1, n_lang + 1):
        cells = set(np.flatnonzero(m.language == lid).tolist())
        seen, todo = set(), [next(iter(cells))]
        while todo:
            c = todo.pop()
            if c in seen:
                continue
            seen.add(c)
            todo += [int(q) for q in m.nb[c] if int(q) in cells]
        assert seen == cells, lid
    # determinism
    m2 = small_member()
    m2.run()
    assert (m2.language == m.language).all() and (m2.population == m.population).all()


# ---- the batched CUDA path against the oracle --------------------------------------------------------------------
def _sweep_case(S=6, R=8, C=10):
    sys.path.insert(0, os.path.join(REPO, "dispersal-simulation_b200", "python"))
    from dsim_b200 import gavin as GB
    P = G.synthetic_precipitation(R, C)
    alpha = 10 ** np.linspace(-8.3, -7.8, S)[:, None]
    beta = np.linspace(2.5, 2.75, S)[:, None]
    popcap = GB.population_capacity(P[None, :], alpha, beta)
    assert (popcap == G.population_capacity(P[None, :], alpha, beta)).all()
    gf = np.linspace(1.05, 1.3, S)
    seed = np.arange(S, dtype=np.uint64) * 7919 + 11
    return GB, R, C, popcap, gf, seed, G.synthetic_lang_caps()


def _check_sweep(lib):
    GB, R, C, popcap, gf, seed, caps = _sweep_case()
    got = GB.run_sweep(lib, R, C, popcap, gf, seed, caps)
    calls = cells = 0
    for k in range(len(gf)):
        m = G.Member(R, C, popcap[k], gf[k], int(seed[k]), caps)
        n_lang = m.run()
        assert got.n_languages[k] == n_lang, (k, got.n_languages[k], n_lang)
        assert (got.language[k] == m.language).all(), f"member {k}: language map"
        assert (got.population[k] == m.population).all(), f"member {k}: populations"
        calls += m.grow_calls
    assert got.grow_calls == calls
    assert len(set(got.n_languages.tolist())) > 1, "the sweep members must differ"
    # a capped run: stop after three languages
    got3 = GB.run_sweep(lib, R, C, popcap[:2], gf[:2], seed[:2], caps, max_languages=3)
    for k in range(2):
        m = G.Member(R, C, popcap[k], gf[k], int(seed[k]), caps, max_languages=3)
        m.run()
        assert (got3.language[k] == m.language).all() and (got3.population[k] == m.population).all()


def test_sweep_matches_oracle_emu():
    import helpers
    _check_sweep(helpers.load_emu())


@pytest.mark.gpu
def test_sweep_matches_oracle_gpu():
    import helpers
    _check_sweep(helpers.D.load())


@pytest.mark.gpu
def test_sweep_members_are_independent_gpu():
    """A member's result does not depend on the other members of the batch (one warp each, no shared state)."""
    import helpers
    GB, R, C, popcap, gf, seed, caps = _sweep_case(S=40, R=12, C=12)
    lib = helpers.D.load()
    whole = GB.run_sweep(lib, R, C, popcap, gf, seed, caps)
    part = GB.run_sweep(lib, R, C, popcap[17:23], gf[17:23], seed[17:23], caps)
    assert (whole.language[17:23] == part.language).all() and (whole.population[17:23] == part.population).all()
