"""Real-graph ingestion and the reference's start state: what feeds the step loop when it is
not driven by synthetic workloads.

* `read_graph_from_db`   `read_graph_from_db` of the reference's `make-graph` binary (src/serialize_graph.rs:9-116) over
                         the sqlite schema of supplement/distances/database.py:24-74 (tables nodes, edges, ecology);
* `contract_unpopulated` the node contraction of its `main` (src/serialize_graph.rs:161-197), including petgraph's
                         `Graph::remove_node` semantics (the last node takes the removed index), which decide the node
                         numbering and which nodes the loop ever looks at;
* `initialization`       `initialization` (lib.rs:1501-1635): two families in Alaska, resources from the population
                         capacities of the patches reachable from them.

The reference serialises the graph with `bincode` into `graph.bincode` (un-vendored crate) between the two programs;
here the graph goes straight into the CSR arrays of the C ABI (`dsim_b200.Graph`).
"""
from __future__ import annotations

import sqlite3
from typing import Dict, List, Tuple

import numpy as np

from . import Families, Graph, NO_PARENT, Params

AMERICAS = (-168.571_541, -34.535_395, -56.028_198, 74.52671)  # west, east, south, north: src/serialize_graph.rs:146-149
ATTESTED_ECOREGIONS = 303                                      # src/ecology.rs:137


class MutableGraph:
    """The part of `petgraph::graph::Graph` the contraction relies on: dense node indices, directed weighted edges,
    `remove_node` = swap-remove."""

    def __init__(self):
        self.nodes: List[Tuple[int, float, float, Dict[int, float]]] = []  # NodeData = (h3, lon, lat, {ecoregion: popcap})
        self.out: List[Dict[int, float]] = []
        self.inc: List[Dict[int, float]] = []

    def add_node(self, data) -> int:
        self.nodes.append(data)
        self.out.append({})
        self.inc.append({})
        return len(self.nodes) - 1

    def add_edge(self, a: int, b: int, w: float):
        self.out[a][b] = w
        self.inc[b][a] = w

    def remove_node(self, a: int):
        """`Graph::remove_node`: drop the node and its edges; the last node moves into index `a`."""
        for b in list(self.out[a]):
            del self.inc[b][a]
        for b in list(self.inc[a]):
            del self.out[b][a]
        last = len(self.nodes) - 1
        if a != last:
            self.nodes[a] = self.nodes[last]
            self.out[a] = self.out[last]
            self.inc[a] = self.inc[last]
            if last in self.out[a]:  # a self-loop of the moved node
                w = self.out[a].pop(last)
                self.inc[a].pop(last)
                self.out[a][a] = w
                self.inc[a][a] = w
            for b, w in self.out[a].items():
                if b != a:
                    del self.inc[b][last]
                    self.inc[b][a] = w
            for b, w in self.inc[a].items():
                if b != a:
                    del self.out[b][last]
                    self.out[b][a] = w
        self.nodes.pop()
        self.out.pop()
        self.inc.pop()

    def to_graph(self) -> Graph:
        n = len(self.nodes)
        eco_off = np.zeros(n + 1, np.uint64)
        edge_off = np.zeros(n + 1, np.uint64)
        eco_id, popcap, edge_dst, edge_sec = [], [], [], []
        for v, (h3, lon, lat, ecos) in enumerate(self.nodes):
            for e in sorted(ecos):  # canonical ecoregion order within a patch: ascending id
                eco_id.append(e)
                popcap.append(ecos[e])
            eco_off[v + 1] = len(eco_id)
            for b in sorted(self.out[v]):
                edge_dst.append(b)
                edge_sec.append(self.out[v][b])
            edge_off[v + 1] = len(edge_dst)
        return Graph(n_nodes=n, eco_off=eco_off, eco_id=np.asarray(eco_id, np.uint32), edge_off=edge_off,
                     edge_dst=np.asarray(edge_dst, np.uint32), edge_sec=np.asarray(edge_sec, np.float64),
                     popcap=np.asarray(popcap, np.float64), h3=np.asarray([x[0] for x in self.nodes], np.uint64),
                     lon=np.asarray([x[1] for x in self.nodes], np.float64), lat=np.asarray([x[2] for x in self.nodes], np.float64))


def read_graph_from_db(path: str, west: float = AMERICAS[0], east: float = AMERICAS[1], south: float = AMERICAS[2],
                       north: float = AMERICAS[3]):
    """src/serialize_graph.rs:9-116.  Returns (MutableGraph, {ecoregion code: dense index}).  Nodes in ascending
    node_id (the order sqlite returns an INTEGER PRIMARY KEY scan in); ecoregion 999 is skipped; ecoregion codes are
    renumbered densely in order of first appearance (`expand_attested`); of several edges between two nodes the
    fastest counts; edges to nodes outside the box are dropped."""
    conn = sqlite3.connect(f"file:{path}?mode=ro", uri=True)
    g = MutableGraph()
    h3_to_graph: Dict[int, int] = {}
    expand: Dict[int, int] = {}
    box = (west, east, south, north)
    rows = conn.execute("SELECT node_id, longitude, latitude FROM nodes "
                        "WHERE ? < longitude AND longitude < ? AND ? < latitude AND latitude < ? ORDER BY node_id", box).fetchall()
    for node_id, lon, lat in rows:
        ecos: Dict[int, float] = {}
        for eco, cap in conn.execute("SELECT ecoregion, population_capacity FROM ecology WHERE node = ? AND ecoregion != 999 "
                                     "ORDER BY ecoregion", (node_id,)):
            ecos[expand.setdefault(int(eco), len(expand))] = float(cap)
        h3_to_graph[node_id] = g.add_node((int(node_id) & 0xFFFFFFFFFFFFFFFF, float(lon), float(lat), ecos))
    if len(expand) > ATTESTED_ECOREGIONS:
        raise ValueError(f"{len(expand)} ecoregions, the model holds {ATTESTED_ECOREGIONS}")  # the reference asserts
    for n1, n2, t in conn.execute("SELECT node1, node2, min(travel_time) FROM edges JOIN nodes ON node1 = node_id "
                                  "WHERE ? < longitude AND longitude < ? AND ? < latitude AND latitude < ? "
                                  "GROUP BY node1, node2 ORDER BY node1, node2", box):
        a, b = h3_to_graph.get(n1), h3_to_graph.get(n2)
        if a is None or b is None:
            continue
        g.add_edge(a, b, float(t))
    conn.close()
    return g, expand


def contract_unpopulated(g: MutableGraph) -> int:
    """src/serialize_graph.rs:161-197: nodes whose population capacities sum to at most 0.01 are removed; every pair
    (in-edge, out-edge) through such a node whose times sum to less than 3 x 8 hours becomes a direct edge (or
    shortens the existing one).  The loop runs over the node indices that existed at the start; `remove_node` moves
    the last node into the freed index, where the loop has already been — as in the reference, that node is not
    examined.  Returns the number of nodes removed."""
    removed = 0
    for node in range(len(g.nodes)):
        if node >= len(g.nodes):  # node_weight(node) == None
            continue
        if sum(g.nodes[node][3].values()) > 0.01:
            continue
        new_edges = []
        for src, w_in in g.inc[node].items():
            for dst, w_out in g.out[node].items():
                w = w_in + w_out
                if src != dst and w < 3.0 * 8.0 * 3600.0:
                    new_edges.append((src, dst, w))
        for src, dst, w in new_edges:
            old = g.out[src].get(dst)
            if old is None or old > w:
                g.add_edge(src, dst, w)
        g.remove_node(node)
        removed += 1
    return removed


def _very_coarse_dist(x1, y1, x2, y2):  # lib.rs:1497-1499
    return abs(x1 - x2) + abs(y1 - y2)


def initialization(graph: Graph, p: Params, scale: float = 1.0):
    """lib.rs:1501-1635 -> (res, max_res, Families).  Start patches: the nodes closest (|dlon| + |dlat|, first of
    equals) to (-159.873, 65.613) and (-158.2718, 60.8071); every patch reachable from them along out-edges is
    filled with `popcap * scale / resource_recovery_per_season` per ecoregion (current = maximum); patches that
    cannot be reached get no resources (the reference creates no Patch for them).  Two families of 5 with 10 stored
    resources, cultures 0xffff and 0, 4 seasons to the next adult."""
    n = graph.n_nodes
    d1 = np.abs(graph.lon - (-159.873)) + np.abs(graph.lat - 65.613)
    d2 = np.abs(graph.lon - (-158.2718)) + np.abs(graph.lat - 60.8071)
    start1, start2 = int(np.argmin(d1)), int(np.argmin(d2))  # argmin = first minimum = the strict `<` of the reference
    seen = np.zeros(n, bool)
    stack = [start1, start2]
    off, dst = graph.edge_off.astype(np.int64), graph.edge_dst
    while stack:
        v = stack.pop()
        if seen[v]:
            continue
        seen[v] = True
        for q in dst[off[v]:off[v + 1]]:
            if not seen[q]:
                stack.append(int(q))
    n_eco = int(graph.eco_off[n])
    node_of = np.repeat(np.arange(n), np.diff(graph.eco_off.astype(np.int64)))
    q = graph.popcap * scale / p.resource_recovery_per_season
    res = np.where(seen[node_of], q, 0.0).astype(np.float64)
    assert len(res) == n_eco
    f = Families.empty(2)
    f.id[:] = [0, 1]
    f.parent_id[:] = NO_PARENT
    f.location[:] = [start1, start2]
    f.culture[:] = [0xFFFF, 0]
    f.size[:] = 5
    f.till_adult[:] = 4
    f.stored[:] = 10.0
    return res, res.copy(), f, (start1, start2)
