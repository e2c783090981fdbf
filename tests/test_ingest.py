"""Real-graph ingestion and the reference's start state: sqlite -> CSR graph
(src/serialize_graph.rs:9-116), contraction of unpopulated nodes (src/serialize_graph.rs:161-197) and `initialization`
(lib.rs:1501-1635), then the step loop on that state against the oracle."""
import ctypes as C
import os
import sqlite3

import numpy as np
import pytest

import helpers
import parity_cases as pc
from helpers import D
from dsim_b200 import ingest


def make_db(path, W=14, H=18, seed=3):
    """A small network.sqlite with the schema of supplement/distances/database.py:24-74: a hex-ish grid around the
    Alaskan start points, a belt of unpopulated nodes, nodes outside the bounding box, an ecoregion 999 row and
    edges with several sources."""
    rng = np.random.default_rng(seed)
    conn = sqlite3.connect(path)
    conn.executescript("""
        CREATE TABLE nodes (node_id INTEGER PRIMARY KEY, longitude FLOAT NOT NULL, latitude FLOAT NOT NULL, h3longitude FLOAT,
                            h3latitude FLOAT, coastal BOOLEAN, popdensity FLOAT, short INTEGER UNIQUE);
        CREATE TABLE edges (node1 INTEGER, node2 INTEGER, travel_time FLOAT, flat_distance FLOAT, source VARCHAR,
                            PRIMARY KEY (node1, node2, source));
        CREATE TABLE ecology (node INTEGER, ecoregion INTEGER, area FLOAT, population_capacity FLOAT, PRIMARY KEY (node, ecoregion));
    """)
    ids = {}
    nid = 600_000_000_000
    for r in range(H):
        for c in range(W):
            nid += int(rng.integers(1, 50))
            ids[(r, c)] = nid
            lon, lat = -162.0 + 0.35 * c + 0.17 * (r % 2), 58.5 + 0.45 * r
            conn.execute("INSERT INTO nodes (node_id, longitude, latitude) VALUES (?, ?, ?)", (nid, lon, lat))
            unpopulated = (r == H // 2 and c % 3 != 1)           # a belt that the contraction removes
            eco = 400 + (r * 3) // H                              # three ecoregion codes, renumbered 0.. in order of appearance
            cap = 0.0 if unpopulated else float(np.clip(20.0 * np.exp(rng.normal()), 0.5, 400.0))
            conn.execute("INSERT INTO ecology VALUES (?, ?, 1.0, ?)", (nid, eco, cap))
            if (r + c) % 7 == 0 and not unpopulated:
                conn.execute("INSERT INTO ecology VALUES (?, ?, 1.0, ?)", (nid, eco + 20, 0.3 * cap))   # a second ecoregion
            if (r + c) % 5 == 0:
                conn.execute("INSERT INTO ecology VALUES (?, 999, 1.0, 77.0)", (nid,))                  # skipped
    outside = nid + 10
    conn.execute("INSERT INTO nodes (node_id, longitude, latitude) VALUES (?, 10.0, 50.0)", (outside,))     # Europe
    conn.execute("INSERT INTO ecology VALUES (?, 400, 1.0, 5.0)", (outside,))
    for (r, c), a in ids.items():
        for dr, dc in ((0, 1), (0, -1), (1, 0), (-1, 0), (1, 1 if r % 2 else -1), (-1, 1 if r % 2 else -1), (0, 2), (2, 0), (-2, 0), (0, -2)):
            b = ids.get((r + dr, c + dc))
            if b is None:
                continue
            t = 14400.0 * max(abs(dr), abs(dc)) * (1.0 + 0.25 * rng.random())
            conn.execute("INSERT INTO edges VALUES (?, ?, ?, 0, 'grid')", (a, b, t))
            if (r + c) % 4 == 0:
                conn.execute("INSERT INTO edges VALUES (?, ?, ?, 0, 'river')", (a, b, 0.8 * t))   # the faster one counts
        if c == 0:
            conn.execute("INSERT INTO edges VALUES (?, ?, 1000.0, 0, 'sea')", (a, outside))        # leaves the box: dropped
    conn.commit()
    conn.close()
    return ids


def test_read_graph(tmp_path):
    path = str(tmp_path / "network.sqlite")
    ids = make_db(path)
    g, expand = ingest.read_graph_from_db(path)
    assert len(g.nodes) == len(ids), "the node outside the bounding box is not read"
    assert [x[0] for x in g.nodes] == sorted(ids.values()), "nodes in ascending node_id"
    assert sorted(expand.values()) == list(range(len(expand))) and 999 not in expand
    first = g.nodes[0][3]
    assert set(first) <= set(expand.values())
    # parallel edges collapse to the fastest; no edge leaves the box
    conn = sqlite3.connect(path)
    a, b, t = conn.execute("SELECT node1, node2, min(travel_time) FROM edges WHERE source IN ('grid','river') GROUP BY node1, node2 "
                           "HAVING count(*) > 1 LIMIT 1").fetchone()
    order = {h3: k for k, (h3, *_rest) in enumerate(g.nodes)}
    assert g.out[order[a]][order[b]] == t
    assert all(dst < len(g.nodes) for o in g.out for dst in o)


def test_contraction_petgraph_semantics():
    """a -> z -> b with z unpopulated: a direct edge a -> b of the summed time appears (if under 24 h), z goes, and the
    LAST node takes z's index (petgraph `remove_node`), unexamined even if unpopulated itself."""
    g = ingest.MutableGraph()
    a = g.add_node((1, 0.0, 0.0, {0: 5.0}))
    z = g.add_node((2, 0.0, 0.0, {0: 0.0}))
    b = g.add_node((3, 0.0, 0.0, {0: 5.0}))
    far = g.add_node((4, 0.0, 0.0, {0: 5.0}))
    last = g.add_node((5, 0.0, 0.0, {}))           # unpopulated, but it will sit at index 1 when the loop is past it
    g.add_edge(a, z, 10_000.0); g.add_edge(z, b, 20_000.0); g.add_edge(z, a, 1.0)
    g.add_edge(far, z, 80_000.0)                    # far -> z -> b takes 100 000 s > 86 400 s: no shortcut
    g.add_edge(a, b, 50_000.0)                      # an existing slower edge is shortened
    g.add_edge(last, a, 7.0)
    assert ingest.contract_unpopulated(g) == 1
    assert [x[0] for x in g.nodes] == [1, 5, 3, 4]
    assert g.out[0] == {2: 30_000.0} and 2 not in g.out[3]
    assert g.out[3] == {0: 80_001.0}, "far -> z -> a takes 80 001 s: a shortcut"
    assert g.out[1] == {0: 7.0} and g.inc[0] == {1: 7.0, 3: 80_001.0}, "the moved node keeps its edges under its new index"


def _real_pair(oracle, dev, synth, tmp_path, prefix="dsim_"):
    path = str(tmp_path / "network.sqlite")
    make_db(path)
    mg, _ = ingest.read_graph_from_db(path)
    assert ingest.contract_unpopulated(mg) > 0
    g = mg.to_graph()
    p = helpers.default_params(oracle)
    p.cooperation_threshold = 10                     # cultures 0xffff and 0 are 16 bits apart: the two lineages stay hostile
    res, mx, fam, starts = ingest.initialization(g, p, scale=1.0)
    assert starts[0] != starts[1] and (res > 0).sum() > 0.9 * len(res)
    o = helpers.oracle_sim(oracle, g, params=p, seed=11, mode=1)
    opt = D.Options()
    getattr(dev, prefix + "default_options")(C.byref(opt))
    d = D.Sim(dev, g, params=p, seed=11, options=opt, prefix=prefix)
    for s in (o, d):
        s.set_patches(res, mx)
        s.set_families(fam)
    return o, d


def test_initialization_and_steps_emu(oracle, emu, synth, tmp_path):
    o, d = _real_pair(oracle, emu, synth, tmp_path)
    trace = pc.run_lockstep(o, d, 200, by_parts=False, check_every=10)
    assert trace[-1] > 2, "the founding families have multiplied (the first split needs ~90 seasons)"


@pytest.mark.gpu
def test_initialization_and_steps_gpu(oracle, gpu_lib, synth, tmp_path):
    o, d = _real_pair(oracle, gpu_lib, synth, tmp_path)
    trace = pc.run_lockstep(o, d, 260, by_parts=False, check_every=20)
    assert trace[-1] > 4


def test_database_has_the_reference_schema(tmp_path):
    """The test database has the tables and columns that supplement/distances/database.py declares (fixture generated
    from its syntax tree, tests/golden/db_schema.json), and the loader queries only columns that exist there."""
    import json
    import re
    ref = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "db_schema.json")))["tables"]
    path = str(tmp_path / "network.sqlite")
    make_db(path)
    conn = sqlite3.connect(path)
    sql_type = {"Integer": "INTEGER", "Float": "FLOAT", "Boolean": "BOOLEAN", "String": "VARCHAR"}
    for table, decl in ref.items():
        have = conn.execute(f"PRAGMA table_info({table})").fetchall()
        assert [(c[1], c[2]) for c in have] == [(c["name"], sql_type[c["type"]]) for c in decl["columns"]], table
        pk = [c[1] for c in sorted(have, key=lambda c: c[5]) if c[5] > 0]
        want_pk = decl["primary_key_constraint"] or [c["name"] for c in decl["columns"] if c["primary_key"]]
        assert pk == want_pk, table
    import inspect
    src = inspect.getsource(ingest.read_graph_from_db)
    used = set(re.findall(r"\b(node_id|longitude|latitude|node1|node2|travel_time|ecoregion|population_capacity|node)\b", src))
    declared = {c["name"] for t in ref.values() for c in t["columns"]}
    assert used <= declared, used - declared
