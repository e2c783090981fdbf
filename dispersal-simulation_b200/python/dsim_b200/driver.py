"""The callers either side of the step loop: the `run` driver of the reference
(src/cli.rs:123-156), its stdout log protocol (lib.rs:649-653, 1443-1464, src/debug.rs:10-21), which the
reference's analysis scripts parse (supplement/plotting/plot.py:290-381), and state files for resuming
(`State::store` / `State::load_from_file`, lib.rs:324-346; src/resume.rs).

The step itself runs behind the C ABI (`Sim`); nothing here touches family state except through it.

Where the reference's text depends on things it leaves undefined (hash-map iteration order) the order is
canonical: patches by ascending node id, cultures of a patch in band creation order.  The reference's
state file is `bincode` of its in-memory `State` (un-vendored crate, unpinned version, forgets the
`storage` map that `step` returns); the state file here is a versioned `.npz` of the structure-of-arrays
state the ABI exchanges, storage and clock included, so a resumed run continues bit-identically.
"""
from __future__ import annotations

import ctypes as C
import io
import sys
from dataclasses import dataclass
from typing import Optional, TextIO

import numpy as np

from . import Families, Graph, Options, Params, Sim, _FAM_FIELDS, NO_PARENT

STATE_FORMAT_VERSION = 1


@dataclass
class Settings:
    """`observation::Settings` (lib.rs:1472-1478); periods in time steps, 0 = never."""
    log_every: int = 1
    log_patch_resources: int = 0
    store_every: int = 0
    statefile: str = "state.npz"


def _rust_f64(x: float) -> str:
    """`{:?}` of an f64: shortest representation that round-trips, always with a decimal point or exponent."""
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "inf" if x > 0 else "-inf"
    r = repr(float(x))
    if "e" in r:  # Python: 1e+21 / 1e-07, Rust: 1e21 / 1e-7
        mant, exp = r.split("e")
        return f"{mant}e{int(exp)}"
    return r


def format_population(graph: Graph, node: np.ndarray, culture: np.ndarray, size: np.ndarray) -> str:
    """`print_population_by_location` (lib.rs:1443-1464): `POPULATION: [(lon, lat, {"<culture hex>": count, ..}), ..]`
    from the `storage` map that `step` returns (lib.rs:622-633), as `dsim_get_storage` delivers it: one entry per
    (patch, band culture), grouped by patch."""
    parts = []
    k, n = 0, len(node)
    while k < n:
        v = int(node[k])
        cults = []
        while k < n and int(node[k]) == v:
            cults.append(f'"{int(culture[k]):x}": {int(size[k])}')
            k += 1
        lon = float(graph.lon[v]) if graph.lon is not None else 0.0
        lat = float(graph.lat[v]) if graph.lat is not None else 0.0
        parts.append(f"({_rust_f64(lon)}, {_rust_f64(lat)}, {{{', '.join(cults)}}})")
    return "POPULATION: [" + ", ".join(parts) + "]"


def _oyr_debug(c: float) -> str:
    """`Debug for OneYearResources` (src/ecology.rs:117-121): `{:.2} ({:.2} kCal)`."""
    return f"{c:.2f} ({c * 2000.0 * 365.0:.2f} kCal)"


def format_first_family(f: Families) -> str:
    """`println!("F: {:?}", families.get(0))` (lib.rs:652) with `Debug for Family` (src/debug.rs:10-21).
    The `descendence` string of the reference ("A:1:3": the whole lineage, lib.rs:1880) is carried by the ABI as
    (id, parent id, birth index); the line names a founder by its id and a descendant as "<parent id>:<birth index>" —
    one level of the reference's name, unique per family, which is all plot.py:367-371 uses it for (a dictionary key)."""
    if f.n == 0:
        return "F: None"
    parent = int(f.parent_id[0])
    name = f"{int(f.id[0])}" if parent == NO_PARENT else f"{parent}:{int(f.birth_index[0])}"
    # `Binary for Culture` (lib.rs:260-272) writes the joined `{:b}` of the words through `write!(f, "{}", ..)`, so the
    # `{:020b}` of src/debug.rs:17 neither pads nor brackets: the culture prints as its plain binary digits
    return ("F: Some(Family { descendence: \"%s\", size: %d, location: NodeIndex(%d), culture: %s, "
            "stored_resources: %s, last_harvest: %s })") % (
        name, int(f.size[0]), int(f.location[0]), format(int(f.culture[0]), "b"),
        _oyr_debug(float(f.stored[0])), _oyr_debug(float(f.last_harvest[0])))


def format_observed_population(node, lon, lat, off, culture, population) -> str:
    """The same line from what `dsim_observe_population` delivers (grouped per node, coordinates attached)."""
    parts = []
    for k in range(len(node)):
        cults = ", ".join(f'"{int(culture[e]):x}": {int(population[e])}' for e in range(int(off[k]), int(off[k + 1])))
        parts.append(f"({_rust_f64(float(lon[k]))}, {_rust_f64(float(lat[k]))}, {{{cults}}})")
    return "POPULATION: [" + ", ".join(parts) + "]"


def log_step(sim: Sim, t: int, out: TextIO):
    """The three lines `step` prints when `t % log_every == 0` (lib.rs:649-653)."""
    out.write(f"t: {t}\n")
    if sim.prefix == "dsim_":   # the product: the observer export of the C ABI
        out.write(format_observed_population(*sim.observe_population()) + "\n")
    else:                       # other libraries driven through the same binding (the test oracle): from the storage map
        node, culture, size = sim.get_storage()
        out.write(format_population(sim.graph, node, culture, size) + "\n")
    out.write(format_first_family(sim.get_families(history=False, adaptation=False)) + "\n")


def store_state(sim: Sim, path: str):
    """`State::store` (lib.rs:325-334): everything needed to continue the run bit-identically."""
    f = sim.get_families()
    res, max_res = sim.get_patches()
    node, culture, size = sim.get_storage()
    arrays = {name: getattr(f, name) for name, _ in _FAM_FIELDS}
    arrays.update(hist_off=f.hist_off, hist_patch=f.hist_patch[:int(f.hist_off[f.n])], hist_harvest=f.hist_harvest[:int(f.hist_off[f.n])],
                  adapt_off=f.adapt_off, adapt_eco=f.adapt_eco[:int(f.adapt_off[f.n])], adapt_val=f.adapt_val[:int(f.adapt_off[f.n])],
                  res=res, max_res=max_res, st_node=node, st_culture=culture, st_size=size,
                  t=np.array([sim.get_time()], np.uint64), seed=np.array([sim.seed], np.uint64),
                  params=np.frombuffer(bytes(sim.params), np.uint8), version=np.array([STATE_FORMAT_VERSION], np.uint32),
                  n_nodes=np.array([sim.graph.n_nodes], np.uint64))
    with open(path, "wb") as fh:  # np.savez appends ".npz" to plain names; a file object is taken as is
        np.savez(fh, **arrays)


def load_state(lib: C.CDLL, graph: Graph, path: str, options: Optional[Options] = None, prefix: str = "dsim_") -> Sim:
    """`State::load_from_file` (lib.rs:336-345) + the set-up of src/resume.rs: a `Sim` that continues the stored run.
    The graph is not part of the file (the reference stores it inside `Parameters`); it must be the one the run
    started with."""
    z = np.load(path)
    if int(z["version"][0]) != STATE_FORMAT_VERSION:
        raise ValueError(f"{path}: state format version {int(z['version'][0])}, this build reads {STATE_FORMAT_VERSION}")
    if int(z["n_nodes"][0]) != graph.n_nodes:
        raise ValueError(f"{path}: stored for a graph of {int(z['n_nodes'][0])} nodes, got {graph.n_nodes}")
    params = Params.from_buffer_copy(z["params"].tobytes())
    sim = Sim(lib, graph, params=params, seed=int(z["seed"][0]), options=options, prefix=prefix)
    n = len(z["id"])
    f = Families(n)
    for name, dt in _FAM_FIELDS:
        setattr(f, name, np.ascontiguousarray(z[name], dtype=dt))
    f.hist_off = np.ascontiguousarray(z["hist_off"], np.uint64)
    f.hist_patch = np.ascontiguousarray(np.concatenate([z["hist_patch"], np.zeros(1, np.uint32)]), np.uint32)
    f.hist_harvest = np.ascontiguousarray(np.concatenate([z["hist_harvest"], np.zeros(1)]), np.float64)
    f.adapt_off = np.ascontiguousarray(z["adapt_off"], np.uint64)
    f.adapt_eco = np.ascontiguousarray(np.concatenate([z["adapt_eco"], np.zeros(1, np.uint32)]), np.uint32)
    f.adapt_val = np.ascontiguousarray(np.concatenate([z["adapt_val"], np.zeros(1)]), np.float64)
    sim.set_patches(np.ascontiguousarray(z["res"]), np.ascontiguousarray(z["max_res"]))
    sim.set_families(f)
    sim.set_storage(np.ascontiguousarray(z["st_node"]), np.ascontiguousarray(z["st_culture"]), np.ascontiguousarray(z["st_size"]))
    sim.set_time(int(z["t"][0]))
    return sim


def run(sim: Sim, max_t: int, o: Settings, out: TextIO = sys.stdout) -> str:
    """`run` (src/cli.rs:123-156): step until the families die out or `t` reaches `max_t`; log and store state at
    the configured periods.  Returns "Died out" or "Ended" (the line the reference prints last).

    Between observations the steps are enqueued in one call (CUDA-graph replay); a read-back of the live count per
    batch serves the extinction check of src/cli.rs:138 (an extinct population stays extinct, so checking at the
    observation points loses nothing but the exact step number in the message)."""
    while True:
        t = sim.get_time()
        # the next step whose result the host must look at: a logging or storing step, or the last one
        nxt = max(max_t, t)
        for period in (o.log_every, o.store_every):
            if period > 0:
                nxt = min(nxt, ((t + period - 1) // period) * period)
        if nxt > t:
            st = sim.step(nxt - t, allow_model_errors=True)
            if st.live_families == 0:
                out.write("Died out\n")
                return "Died out"
        st = sim.step(1, allow_model_errors=True)  # the step the reference runs with State.t == nxt
        if o.log_every > 0 and nxt % o.log_every == 0:
            log_step(sim, nxt, out)
        if st.live_families == 0:
            out.write("Died out\n")
            return "Died out"
        if nxt == max_t or (o.store_every > 0 and nxt % o.store_every == 0):
            out.write(f"{nxt} % {o.store_every}\n")
            store_state(sim, o.statefile)
        if nxt >= max_t:
            out.write("Ended\n")
            return "Ended"


def parse_log(text: str):
    """What supplement/plotting/plot.py:336-369 extracts from a log: [(t, [(lon, lat, count, culture), ..])] and the
    path of the first family [(name, node)].  Same expressions as the reference's parser (literal_eval of the
    POPULATION line, the regular expression of the F line)."""
    import re
    from ast import literal_eval
    steps, path, t = [], [], None
    for line in io.StringIO(text):
        if line.startswith("t: "):
            t = int(line[3:])
        elif line.startswith("POPULATION: ["):
            content = [(x, y, n, int(c, 16)) for x, y, p in literal_eval(line[len("POPULATION: "):]) for c, n in p.items()]
            steps.append((t, content))
        elif line.startswith("F:"):
            m = re.search(r'escendence: "([^"]*)".*ocation: NodeIndex\(([0-9]*)\),.*culture: ([01]*)', line)
            if m:
                path.append((m.group(1), int(m.group(2))))
    return steps, path
