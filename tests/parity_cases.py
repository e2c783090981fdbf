"""Parity scenarios shared by the CPU (emulated kernels) and GPU test modules.

Every scenario builds the same synthetic state for the oracle and for the device path, advances
both and requires: integer state bit-exact (location, size, culture, clocks, ids, lineage, history
patches, storage cultures), floats bit-exact as well (the kernels use the oracle's operation order
with FMA contraction off) — the north-star tolerance for floats is 1e-5 relative, which
`rtol=0.0` tightens to zero.
"""
from __future__ import annotations

import numpy as np

import helpers
from helpers import D


def build_pair(oracle_lib, dev_lib, synth, W, H, n_fam, seed=1, sim_seed=7, n_occupied=0, hist_fill=0,
               params_edit=None, hist_cap=0, graph_edit=None, dev_prefix="dsim_", use_graphs=1, flags=0, families_edit=None):
    g, res, mx = D.synth_grid(W, H, seed, lib=synth)
    if graph_edit is not None:
        g, res, mx = graph_edit(g, res, mx)
    fam = D.synth_families(W, H, n_fam, seed, n_occupied=n_occupied, hist_fill=hist_fill, lib=synth)
    if families_edit is not None:
        families_edit(fam)
    p = helpers.default_params(oracle_lib)
    if params_edit is not None:
        params_edit(p)
    o = helpers.oracle_sim(oracle_lib, g, params=p, seed=sim_seed, mode=1)
    opt = D.Options()
    getattr(dev_lib, dev_prefix + "default_options")(opt)
    opt.history_capacity = hist_cap
    opt.use_graphs = use_graphs
    opt.flags = flags | D.OPT_SHORT_HISTORY
    d = D.Sim(dev_lib, g, params=p, seed=sim_seed, options=opt, prefix=dev_prefix)
    for s in (o, d):
        s.set_patches(res, mx)
        s.set_families(fam)
    return o, d


def run_lockstep(o, d, steps, by_parts=True, hist_cap=None, check_every=1):
    """Advance both sims; compare after every part (or step).  Returns the population trace."""
    trace = []
    for k in range(steps):
        if by_parts:
            for part in (1, 2):
                o.step_part(part)
                d.step_part(part)
                if k % check_every == 0:
                    msgs = helpers.compare_sims(o, d, hist_cap=hist_cap, what=f"step {k} part {part}")
                    assert not msgs, "\n".join(msgs)
        else:
            so = o.step(1, allow_model_errors=True)
            sd = d.step(1, allow_model_errors=True)
            assert so.as_dict() == sd.as_dict(), (so.as_dict(), sd.as_dict())
            if k % check_every == 0:
                msgs = helpers.compare_sims(o, d, hist_cap=hist_cap, what=f"step {k}")
                assert not msgs, "\n".join(msgs)
        trace.append(o.family_counts()[0])
    return trace


def two_ecoregion_graph(g, res, mx):
    """Give every third node a second ecoregion (the real graph has 1-3 per patch, lib.rs:142-152)."""
    n = g.n_nodes
    counts = np.where(np.arange(n) % 3 == 0, 2, 1).astype(np.uint64)
    eco_off = np.zeros(n + 1, np.uint64)
    eco_off[1:] = np.cumsum(counts)
    total = int(eco_off[n])
    eco_id = np.zeros(total, np.uint32)
    res2 = np.zeros(total)
    mx2 = np.zeros(total)
    pop2 = np.zeros(total)
    for v in range(n):
        a = int(eco_off[v])
        first = int(g.eco_id[v])
        eco_id[a] = first
        res2[a], mx2[a], pop2[a] = res[v], mx[v], g.popcap[v]
        if counts[v] == 2:
            eco_id[a + 1] = first + 100          # ascending within the node
            res2[a + 1], mx2[a + 1], pop2[a + 1] = 0.5 * res[v], 0.7 * mx[v], 0.7 * g.popcap[v]
    g2 = D.Graph(n_nodes=n, eco_off=eco_off, eco_id=eco_id, edge_off=g.edge_off, edge_dst=g.edge_dst,
                 edge_sec=g.edge_sec, popcap=pop2, h3=g.h3, lon=g.lon, lat=g.lat)
    return g2, res2, mx2


def sparse_edges_graph(g, res, mx):
    """Drop 40 % of the edges and stretch some beyond the season range: nearby is no longer a subset of
    reach, rows have ragged lengths, a few nodes become sinks."""
    rng = np.random.default_rng(5)
    n = g.n_nodes
    keep = rng.random(len(g.edge_dst)) > 0.4
    sec = g.edge_sec.copy()
    far = rng.random(len(sec)) < 0.1
    sec[far] *= 6.0
    counts = np.zeros(n, np.uint64)
    src = np.repeat(np.arange(n), np.diff(g.edge_off.astype(np.int64)))
    kill_nodes = rng.choice(n, size=max(1, n // 50), replace=False)
    keep &= ~np.isin(src, kill_nodes)
    np.add.at(counts, src[keep], 1)
    edge_off = np.zeros(n + 1, np.uint64)
    edge_off[1:] = np.cumsum(counts)
    g2 = D.Graph(n_nodes=n, eco_off=g.eco_off, eco_id=g.eco_id, edge_off=edge_off,
                 edge_dst=np.ascontiguousarray(g.edge_dst[keep]), edge_sec=np.ascontiguousarray(sec[keep]),
                 popcap=g.popcap, h3=g.h3, lon=g.lon, lat=g.lat)
    return g2, res, mx


def long_range_graph(W):
    """Add four fast long-range edges per node (3 columns / 3 rows away): the 2-hop `nearby` lists grow beyond 64
    entries, so the sensed patches of a family take more than one 64-entry pass of k_move."""
    def edit(g, res, mx):
        n = g.n_nodes
        off = g.edge_off.astype(np.int64)
        src = np.repeat(np.arange(n, dtype=np.int64), np.diff(off))
        extra_src, extra_dst = [], []
        for delta in (3, -3, 3 * W, -3 * W):
            v = np.arange(n, dtype=np.int64)
            w = v + delta
            ok = (w >= 0) & (w < n)
            if abs(delta) == 3:                      # stay inside the grid row
                ok &= (v // W) == (w // W)
            extra_src.append(v[ok])
            extra_dst.append(w[ok])
        all_src = np.concatenate([src] + extra_src)
        all_dst = np.concatenate([g.edge_dst.astype(np.int64)] + extra_dst)
        all_sec = np.concatenate([g.edge_sec] + [np.full(len(x), 9000.0) for x in extra_src])
        order = np.lexsort((all_dst, all_src))
        all_src, all_dst, all_sec = all_src[order], all_dst[order], all_sec[order]
        keep = np.ones(len(all_src), bool)           # drop duplicates of existing edges (keep the first)
        keep[1:] = (all_src[1:] != all_src[:-1]) | (all_dst[1:] != all_dst[:-1])
        all_src, all_dst, all_sec = all_src[keep], all_dst[keep], all_sec[keep]
        edge_off = np.zeros(n + 1, np.uint64)
        edge_off[1:] = np.cumsum(np.bincount(all_src, minlength=n))
        g2 = D.Graph(n_nodes=n, eco_off=g.eco_off, eco_id=g.eco_id, edge_off=edge_off,
                     edge_dst=np.ascontiguousarray(all_dst.astype(g.edge_dst.dtype)), edge_sec=np.ascontiguousarray(all_sec),
                     popcap=g.popcap, h3=g.h3, lon=g.lon, lat=g.lat)
        return g2, res, mx
    return edit


def check_reach(o, d):
    a, b = o.get_reach(), d.get_reach()
    for x, y, name in zip(a, b, ["near_off", "near_dst", "reach_off", "reach_dst", "reach_cost"]):
        assert len(x) == len(y), name
        if x.dtype == np.float64:
            assert (x.view(np.uint64) == y.view(np.uint64)).all(), name
        else:
            assert (x == y).all(), name


def hostile_cultures(n_cultures, seed=5):
    """families_edit: every family gets one of `n_cultures` random 64-bit cultures (pairwise about 32 bits apart, so
    members of different cultures never cooperate): crowded nodes then hold several bands, and most newcomers fight
    their way past the bands before theirs (lib.rs:1209-1232, 1380-1399), wiping out members on the way."""
    def edit(fam):
        rng = np.random.default_rng(seed)
        pool = rng.integers(0, 1 << 63, n_cultures, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, n_cultures, dtype=np.uint64)
        fam.culture[:] = pool[rng.integers(0, n_cultures, len(fam.culture))]
    return edit
