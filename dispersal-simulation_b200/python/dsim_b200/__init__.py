"""ctypes binding of the C ABI in include/dsim.h.

Host glue only: the product is the CUDA library `libdsim_b200.so` built from
dispersal-simulation_b200/csrc.  There is no CPU fallback — `load()` raises when the library is
missing and `Sim` raises when no GPU is present.

The same `Sim` class can be pointed at any library that exports the ABI under another prefix; the
parity tests use that to drive the CPU oracle (prefix ``oracle_``) with the same code.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.normpath(os.path.join(_HERE, "..", ".."))
REPO_ROOT = os.path.normpath(os.path.join(PKG_ROOT, ".."))
LIB_PATH = os.environ.get("DSIM_B200_LIB") or os.path.join(PKG_ROOT, "libdsim_b200.so")  # env: developer A/B builds
SYNTH_LIB_PATH = os.path.join(PKG_ROOT, "libdsim_synth.so")

N_ECOREGIONS = 304
OPT_PATCH_WARP_ONLY = 0x2
OPT_REACH_ON_DEVICE = 0x4
OPT_REACH_ON_HOST = 0x8
OPT_NO_PDL = 0x10
OPT_SHORT_HISTORY = 0x20  # accept a history ring shorter than the exact length (tests whose histories never outgrow it)
NO_PARENT = 2**64 - 1
MG_IPC_BYTES = 384

u8p = C.POINTER(C.c_uint8)
u16p = C.POINTER(C.c_uint16)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
f64p = C.POINTER(C.c_double)


class DsimError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"dsim error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    """`Parameters` of src/parameters.rs:5-20 (without the graph)."""
    _fields_ = [
        ("resource_recovery_per_season", C.c_double),
        ("culture_mutation_rate", C.c_double),
        ("culture_dimensionality", C.c_uint64),
        ("cooperation_threshold", C.c_uint32),
        ("fight_deadliness", C.c_uint32),
        ("maximum_resources_one_adult_can_harvest", C.c_double),
        ("evidence_needed", C.c_double),
        ("payoff_std", C.c_double),
        ("minimum_adaptation", C.c_double),
        ("enemy_discount", C.c_double),
        ("season_length_in_years", C.c_double),
        ("cooperation_inclusive", C.c_uint32),
        ("reserved_", C.c_uint32),
    ]


class GraphC(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_uint32),
        ("h3", u64p), ("lon", f64p), ("lat", f64p),
        ("eco_off", u64p), ("eco_id", u32p), ("popcap", f64p),
        ("edge_off", u64p), ("edge_dst", u32p), ("edge_sec", f64p),
    ]


class FamiliesC(C.Structure):
    _fields_ = [
        ("n", C.c_uint64),
        ("id", u64p), ("parent_id", u64p), ("birth_index", u16p), ("culture", u64p),
        ("location", u32p), ("size", u32p), ("stored", f64p), ("last_harvest", f64p),
        ("till_adult", u32p), ("till_mutation", u32p), ("has_mutation_clock", u8p),
        ("n_offspring", u16p),
        ("hist_off", u64p), ("hist_patch", u32p), ("hist_harvest", f64p),
        ("adapt_off", u64p), ("adapt_eco", u32p), ("adapt_val", f64p),
    ]


class Options(C.Structure):
    _fields_ = [
        ("family_capacity", C.c_uint64),
        ("history_capacity", C.c_uint32),
        ("device", C.c_uint32),
        ("use_graphs", C.c_uint32),
        ("flags", C.c_uint32),
    ]


class StepStats(C.Structure):
    _fields_ = [
        ("t", C.c_uint64), ("live_families", C.c_uint64), ("births", C.c_uint64),
        ("deaths", C.c_uint64), ("family_updates", C.c_uint64), ("occupied_patches", C.c_uint64),
        ("model_errors", C.c_uint32), ("reserved_", C.c_uint32),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved_"}


MAX_KERNELS = 16


class KernelTimes(C.Structure):
    _fields_ = [
        ("n", C.c_uint32),
        ("name", C.c_char_p * MAX_KERNELS),
        ("ms", C.c_double * MAX_KERNELS),
        ("launches", C.c_uint64 * MAX_KERNELS),
    ]


def _ptr(a: Optional[np.ndarray], ctype):
    if a is None:
        return C.cast(None, C.POINTER(ctype))
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(ctype))


@dataclass
class Graph:
    """`MovementGraph` (src/movementgraph.rs:8-9) in CSR form, numpy arrays."""
    n_nodes: int
    eco_off: np.ndarray
    eco_id: np.ndarray
    edge_off: np.ndarray
    edge_dst: np.ndarray
    edge_sec: np.ndarray
    popcap: Optional[np.ndarray] = None
    h3: Optional[np.ndarray] = None
    lon: Optional[np.ndarray] = None
    lat: Optional[np.ndarray] = None

    def as_c(self) -> GraphC:
        return GraphC(
            self.n_nodes, _ptr(self.h3, C.c_uint64), _ptr(self.lon, C.c_double), _ptr(self.lat, C.c_double),
            _ptr(self.eco_off, C.c_uint64), _ptr(self.eco_id, C.c_uint32), _ptr(self.popcap, C.c_double),
            _ptr(self.edge_off, C.c_uint64), _ptr(self.edge_dst, C.c_uint32), _ptr(self.edge_sec, C.c_double))


_FAM_FIELDS = [
    ("id", np.uint64), ("parent_id", np.uint64), ("birth_index", np.uint16), ("culture", np.uint64),
    ("location", np.uint32), ("size", np.uint32), ("stored", np.float64), ("last_harvest", np.float64),
    ("till_adult", np.uint32), ("till_mutation", np.uint32), ("has_mutation_clock", np.uint8),
    ("n_offspring", np.uint16),
]
_CT = {np.uint64: C.c_uint64, np.uint32: C.c_uint32, np.uint16: C.c_uint16, np.uint8: C.c_uint8,
       np.float64: C.c_double}


@dataclass
class Families:
    """`Vec<Family>` (lib.rs:178-229) as structure-of-arrays; see include/dsim.h."""
    n: int
    id: np.ndarray = None
    parent_id: np.ndarray = None
    birth_index: np.ndarray = None
    culture: np.ndarray = None
    location: np.ndarray = None
    size: np.ndarray = None
    stored: np.ndarray = None
    last_harvest: np.ndarray = None
    till_adult: np.ndarray = None
    till_mutation: np.ndarray = None
    has_mutation_clock: np.ndarray = None
    n_offspring: np.ndarray = None
    hist_off: np.ndarray = None
    hist_patch: np.ndarray = None
    hist_harvest: np.ndarray = None
    adapt_off: np.ndarray = None
    adapt_eco: np.ndarray = None
    adapt_val: np.ndarray = None

    @staticmethod
    def empty(n: int, n_hist: int = 0, n_adapt: int = 0) -> "Families":
        f = Families(n)
        for name, dt in _FAM_FIELDS:
            setattr(f, name, np.zeros(n, dtype=dt))
        f.parent_id[:] = NO_PARENT
        f.hist_off = np.zeros(n + 1, dtype=np.uint64)
        f.hist_patch = np.zeros(max(n_hist, 1), dtype=np.uint32)
        f.hist_harvest = np.zeros(max(n_hist, 1), dtype=np.float64)
        f.adapt_off = np.zeros(n + 1, dtype=np.uint64)
        f.adapt_eco = np.zeros(max(n_adapt, 1), dtype=np.uint32)
        f.adapt_val = np.zeros(max(n_adapt, 1), dtype=np.float64)
        return f

    def as_c(self) -> FamiliesC:
        c = FamiliesC()
        c.n = self.n
        for name, dt in _FAM_FIELDS:
            setattr(c, name, _ptr(getattr(self, name), _CT[dt]))
        c.hist_off = _ptr(self.hist_off, C.c_uint64)
        c.hist_patch = _ptr(self.hist_patch, C.c_uint32)
        c.hist_harvest = _ptr(self.hist_harvest, C.c_double)
        c.adapt_off = _ptr(self.adapt_off, C.c_uint64)
        c.adapt_eco = _ptr(self.adapt_eco, C.c_uint32)
        c.adapt_val = _ptr(self.adapt_val, C.c_double)
        return c

    def history_of(self, i: int):
        a, b = int(self.hist_off[i]), int(self.hist_off[i + 1])
        return self.hist_patch[a:b], self.hist_harvest[a:b]

    def adaptation_of(self, i: int):
        a, b = int(self.adapt_off[i]), int(self.adapt_off[i + 1])
        return self.adapt_eco[a:b], self.adapt_val[a:b]


def load(path: str = LIB_PATH) -> C.CDLL:
    """Load the CUDA library.  Fails loudly when it has not been built (no fallback)."""
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} not found: build it with `make -C {PKG_ROOT}` (or __graft_entry__.build()); "
            "there is no CPU fallback for the step loop")
    return C.CDLL(path, mode=C.RTLD_GLOBAL)


def load_synth(path: str = SYNTH_LIB_PATH) -> C.CDLL:
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} not found: build it with `make -C {PKG_ROOT} synth`")
    return C.CDLL(path)


class Sim:
    """One simulation state behind the C ABI (State of lib.rs:292-318 + `run` of src/cli.rs:123-156)."""

    def __init__(self, lib: C.CDLL, graph: Graph, params: Optional[Params] = None, seed: int = 1,
                 options: Optional[Options] = None, prefix: str = "dsim_"):
        self.lib, self.prefix, self.graph, self.seed = lib, prefix, graph, int(seed)
        self._declare()
        self.params = params if params is not None else self.default_params()
        self.options = options if options is not None else self.default_options()
        self._g = graph.as_c()
        self.h = C.c_void_p()
        rc = self._f("create")(C.byref(self.params), C.byref(self._g), C.c_uint64(seed), C.byref(self.options),
                               C.byref(self.h))
        if rc != 0:
            msg = self._f("last_error")(None)
            raise DsimError(rc, msg.decode() if msg else "create failed")
        self.n_eco = int(graph.eco_off[graph.n_nodes])

    # -- plumbing ---------------------------------------------------------------------------
    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    def _declare(self):
        vp = C.c_void_p
        sig = {
            "create": (C.c_int, [C.POINTER(Params), C.POINTER(GraphC), C.c_uint64, C.POINTER(Options), C.POINTER(vp)]),
            "destroy": (None, [vp]),
            "last_error": (C.c_char_p, [vp]),
            "default_params": (None, [C.POINTER(Params)]),
            "default_options": (None, [C.POINTER(Options)]),
            "set_patches": (C.c_int, [vp, f64p, f64p]),
            "get_patches": (C.c_int, [vp, f64p, f64p]),
            "set_families": (C.c_int, [vp, C.POINTER(FamiliesC)]),
            "family_counts": (C.c_int, [vp, u64p, u64p, u64p]),
            "get_families": (C.c_int, [vp, C.POINTER(FamiliesC)]),
            "set_time": (C.c_int, [vp, C.c_uint64]),
            "get_time": (C.c_int, [vp, u64p]),
            "storage_count": (C.c_int, [vp, u64p]),
            "get_storage": (C.c_int, [vp, u32p, u64p, u64p]),
            "set_storage": (C.c_int, [vp, C.c_uint64, u32p, u64p, u64p]),
            "step": (C.c_int, [vp, C.c_uint32, C.POINTER(StepStats)]),
            "sync": (C.c_int, [vp]),
            "step_part": (C.c_int, [vp, C.c_int]),
            "reach_counts": (C.c_int, [vp, u64p, u64p]),
            "get_reach": (C.c_int, [vp, u64p, u32p, u64p, u32p, f64p]),
        }
        if self.prefix == "dsim_":
            sig["reach_digest"] = (C.c_int, [vp, u64p])
            sig["observe_population_counts"] = (C.c_int, [vp, u64p, u64p])
            sig["observe_population"] = (C.c_int, [vp, u32p, f64p, f64p, u64p, u64p, u64p])
        for name, (res, args) in sig.items():
            fn = self._f(name)
            fn.restype, fn.argtypes = res, args
        for name, (res, args) in {
            "profile_enable": (C.c_int, [vp, C.c_int]),
            "profile_read": (C.c_int, [vp, C.POINTER(KernelTimes), C.c_int]),
            "launch_count": (C.c_int, [vp, u64p]),
            "step_timed": (C.c_int, [vp, C.c_uint32, C.POINTER(StepStats), f64p]),
            "mg_configure": (C.c_int, [vp, C.c_uint32, C.c_uint32, u32p]),
            "mg_halo_width": (C.c_int, [vp, u32p]),
            "mg_phase": (C.c_int, [vp, C.c_int]),
            "mg_buffer": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), u64p]),
            "mg_memcpy": (C.c_int, [vp, C.c_void_p, C.c_void_p, C.c_uint64]),
            "mg_unique_id": (C.c_int, [C.c_void_p]),
            "mg_init_nccl": (C.c_int, [vp, C.c_void_p]),
            "mg_ipc_export": (C.c_int, [vp, C.c_void_p]),
            "mg_ipc_connect": (C.c_int, [vp, C.c_void_p]),
            "mg_halo_sync": (C.c_int, [vp]),
            "mg_reduce_stats": (C.c_int, [vp, C.POINTER(StepStats)]),
        }.items():
            if hasattr(self.lib, self.prefix + name):
                fn = self._f(name)
                fn.restype, fn.argtypes = res, args

    def _check(self, rc, allow_model=False):
        if rc != 0 and not (allow_model and rc == -5):
            msg = self._f("last_error")(self.h)
            raise DsimError(rc, msg.decode() if msg else "")
        return rc

    def default_params(self) -> Params:
        p = Params()
        self._f("default_params")(C.byref(p))
        return p

    def default_options(self) -> Options:
        o = Options()
        self._f("default_options")(C.byref(o))
        return o

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self._f("destroy")(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- state ------------------------------------------------------------------------------
    def set_patches(self, res: np.ndarray, max_res: np.ndarray):
        res = np.ascontiguousarray(res, dtype=np.float64)
        max_res = np.ascontiguousarray(max_res, dtype=np.float64)
        assert res.shape[0] == self.n_eco and max_res.shape[0] == self.n_eco
        self._check(self._f("set_patches")(self.h, _ptr(res, C.c_double), _ptr(max_res, C.c_double)))

    def get_patches(self, maximum: bool = True, out=None):
        """(current resources, maxima) per (patch, ecoregion) entry.  `maximum=False` skips the maxima (they never
        change; None is returned for them); `out` = a caller-owned float64 array for the resources (e.g. page-locked)."""
        res = np.zeros(self.n_eco, dtype=np.float64) if out is None else out
        if res.dtype != np.float64 or res.shape != (self.n_eco,) or not res.flags.c_contiguous:
            raise ValueError("out must be a contiguous float64 array with one entry per (patch, ecoregion)")
        mx = np.zeros(self.n_eco, dtype=np.float64) if maximum else None
        self._check(self._f("get_patches")(self.h, _ptr(res, C.c_double), _ptr(mx, C.c_double) if maximum else None))
        return res, mx

    def set_families(self, f: Families):
        c = f.as_c()
        self._check(self._f("set_families")(self.h, C.byref(c)))

    def family_counts(self, history: bool = True, adaptation: bool = True):
        """(families, history entries, adaptation entries); a count that is not asked for is not computed (0)."""
        n, nh, na = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self._f("family_counts")(self.h, C.byref(n), C.byref(nh) if history else None,
                                             C.byref(na) if adaptation else None))
        return n.value, nh.value, na.value

    def get_families(self, history: bool = True, adaptation: bool = True, out: "Families | None" = None) -> Families:
        """`out` = caller-owned column arrays to fill (core columns only: history=False, adaptation=False), e.g. the
        buffers of a checkpoint writer that downloads the state repeatedly; they must hold at least the live families."""
        n, nh, na = self.family_counts(history, adaptation)
        if out is not None:
            if history or adaptation:
                raise ValueError("out= takes the core columns only")
            if len(out.id) < n:
                raise ValueError("out is too small")
            f = Families(n)
            for name, _ in _FAM_FIELDS:
                setattr(f, name, getattr(out, name))
            f.hist_off = f.hist_patch = f.hist_harvest = f.adapt_off = f.adapt_eco = f.adapt_val = None
            c = f.as_c()
            self._check(self._f("get_families")(self.h, C.byref(c)))
            for name, _ in _FAM_FIELDS:
                setattr(f, name, getattr(out, name)[:n])
            return f
        f = Families.empty(n, nh if history else 0, na if adaptation else 0)
        if not history:
            f.hist_off = f.hist_patch = f.hist_harvest = None
        if not adaptation:
            f.adapt_off = f.adapt_eco = f.adapt_val = None
        c = f.as_c()
        self._check(self._f("get_families")(self.h, C.byref(c)))
        if history:
            f.hist_patch, f.hist_harvest = f.hist_patch[:nh], f.hist_harvest[:nh]
        if adaptation:
            f.adapt_eco, f.adapt_val = f.adapt_eco[:na], f.adapt_val[:na]
        return f

    def set_time(self, t: int):
        self._check(self._f("set_time")(self.h, t))

    def get_time(self) -> int:
        t = C.c_uint64()
        self._check(self._f("get_time")(self.h, C.byref(t)))
        return t.value

    def get_storage(self):
        n = C.c_uint64()
        self._check(self._f("storage_count")(self.h, C.byref(n)))
        node = np.zeros(max(n.value, 1), dtype=np.uint32)
        cul = np.zeros(max(n.value, 1), dtype=np.uint64)
        size = np.zeros(max(n.value, 1), dtype=np.uint64)
        self._check(self._f("get_storage")(self.h, _ptr(node, C.c_uint32), _ptr(cul, C.c_uint64), _ptr(size, C.c_uint64)))
        return node[:n.value], cul[:n.value], size[:n.value]

    def set_storage(self, node, culture, size):
        node = np.ascontiguousarray(node, dtype=np.uint32)
        culture = np.ascontiguousarray(culture, dtype=np.uint64)
        size = np.ascontiguousarray(size, dtype=np.uint64)
        self._check(self._f("set_storage")(self.h, len(node), _ptr(node, C.c_uint32), _ptr(culture, C.c_uint64),
                                           _ptr(size, C.c_uint64)))

    # -- stepping ---------------------------------------------------------------------------
    def step(self, n_steps: int = 1, want_stats: bool = True, allow_model_errors: bool = False):
        if want_stats:
            st = StepStats()
            self._check(self._f("step")(self.h, n_steps, C.byref(st)), allow_model_errors)
            return st
        self._check(self._f("step")(self.h, n_steps, None), allow_model_errors)
        return None

    def step_timed(self, n_steps: int):
        """n_steps steps timed with CUDA events on the library's stream -> (stats, device milliseconds)."""
        st, ms = StepStats(), C.c_double()
        self._check(self._f("step_timed")(self.h, n_steps, C.byref(st), C.byref(ms)))
        return st, ms.value

    def sync(self):
        self._check(self._f("sync")(self.h))

    def step_part(self, part: int):
        self._check(self._f("step_part")(self.h, part))

    def reach_counts(self):
        """(entries of the 2-hop lists, entries of the reach rows) this handle holds: everything, or on a multi-GPU handle
        the rows of the own band and its halo."""
        a, b = C.c_uint64(), C.c_uint64()
        self._check(self._f("reach_counts")(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def get_reach(self):
        a, b = C.c_uint64(), C.c_uint64()
        self._check(self._f("reach_counts")(self.h, C.byref(a), C.byref(b)))
        n = self.graph.n_nodes
        near_off = np.zeros(n + 1, dtype=np.uint64)
        near_dst = np.zeros(max(a.value, 1), dtype=np.uint32)
        reach_off = np.zeros(n + 1, dtype=np.uint64)
        reach_dst = np.zeros(max(b.value, 1), dtype=np.uint32)
        reach_cost = np.zeros(max(b.value, 1), dtype=np.float64)
        self._check(self._f("get_reach")(self.h, _ptr(near_off, C.c_uint64), _ptr(near_dst, C.c_uint32),
                                         _ptr(reach_off, C.c_uint64), _ptr(reach_dst, C.c_uint32),
                                         _ptr(reach_cost, C.c_double)))
        return near_off, near_dst[:a.value], reach_off, reach_dst[:b.value], reach_cost[:b.value]

    def observe_population(self):
        """`print_population_by_location` (lib.rs:1443-1464) as arrays: (node, lon, lat, off, culture, population);
        entries [off[k], off[k+1]) are the cultures of node[k]."""
        a, b = C.c_uint64(), C.c_uint64()
        self._check(self._f("observe_population_counts")(self.h, C.byref(a), C.byref(b)))
        node = np.zeros(max(a.value, 1), np.uint32)
        lon, lat = np.zeros(max(a.value, 1)), np.zeros(max(a.value, 1))
        off = np.zeros(a.value + 1, np.uint64)
        cult, pop = np.zeros(max(b.value, 1), np.uint64), np.zeros(max(b.value, 1), np.uint64)
        self._check(self._f("observe_population")(self.h, _ptr(node, C.c_uint32), _ptr(lon, C.c_double), _ptr(lat, C.c_double),
                                                  _ptr(off, C.c_uint64), _ptr(cult, C.c_uint64), _ptr(pop, C.c_uint64)))
        return node[:a.value], lon[:a.value], lat[:a.value], off, cult[:b.value], pop[:b.value]

    def reach_digest(self):
        """Five order-independent digests of the reach table in HBM (headers, destinations, costs, bucket indices,
        2-hop lists)."""
        out = (C.c_uint64 * 5)()
        self._check(self._f("reach_digest")(self.h, out))
        return tuple(int(x) for x in out)

    # -- multi-GPU (band partition; include/dsim.h) ---------------------------------------------
    def mg_configure(self, rank: int, world: int, bounds):
        b = np.ascontiguousarray(bounds, dtype=np.uint32)
        assert len(b) == world + 1
        self._check(self._f("mg_configure")(self.h, rank, world, _ptr(b, C.c_uint32)))
        self.mg_rank, self.mg_world, self.mg_bounds = rank, world, b

    def mg_halo_width(self) -> int:
        w = C.c_uint32()
        self._check(self._f("mg_halo_width")(self.h, C.byref(w)))
        return w.value

    def mg_phase(self, phase: int):
        self._check(self._f("mg_phase")(self.h, phase))

    def mg_buffer(self, channel: int, peer: int, recv: bool):
        p, n = C.c_void_p(), C.c_uint64()
        self._check(self._f("mg_buffer")(self.h, channel, peer, 1 if recv else 0, C.byref(p), C.byref(n)))
        return p.value, n.value

    def mg_memcpy(self, dst, src, nbytes):
        self._check(self._f("mg_memcpy")(self.h, dst, src, nbytes))

    def mg_init_nccl(self, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        self._check(self._f("mg_init_nccl")(self.h, buf))
        self._check(self._f("mg_halo_sync")(self.h))

    def mg_halo_sync(self):
        """Initial halo exchange over the built-in transport (every rank, after the state was set everywhere)."""
        self._check(self._f("mg_halo_sync")(self.h))

    def mg_ipc_export(self) -> bytes:
        """CUDA IPC handles of this rank's receive buffers (peer-memory transport, include/dsim.h)."""
        buf = C.create_string_buffer(MG_IPC_BYTES)
        self._check(self._f("mg_ipc_export")(self.h, buf))
        return buf.raw

    def mg_ipc_connect(self, handles_of_all_ranks):
        """`handles_of_all_ranks[r]` = mg_ipc_export() of rank r; from now on the steps exchange through peer memory."""
        blob = b"".join(handles_of_all_ranks)
        assert len(blob) == MG_IPC_BYTES * len(handles_of_all_ranks)
        self._check(self._f("mg_ipc_connect")(self.h, C.create_string_buffer(blob, len(blob))))

    def mg_reduce_stats(self, st: "StepStats"):
        self._check(self._f("mg_reduce_stats")(self.h, C.byref(st)))
        return st

    # -- measurement ------------------------------------------------------------------------
    def profile_enable(self, on: bool):
        self._check(self._f("profile_enable")(self.h, 1 if on else 0))

    def profile_read(self, reset: bool = True):
        kt = KernelTimes()
        self._check(self._f("profile_read")(self.h, C.byref(kt), 1 if reset else 0))
        return {kt.name[i].decode(): (kt.ms[i], kt.launches[i]) for i in range(kt.n)}

    def launch_count(self) -> int:
        n = C.c_uint64()
        self._check(self._f("launch_count")(self.h, C.byref(n)))
        return n.value


# -- synthetic workloads (include/dsim_synth.h) ---------------------------------------------------
def synth_grid(W: int, H: int, seed: int = 1, recovery: float = 0.10, lib: Optional[C.CDLL] = None):
    lib = lib or load_synth()
    nn, ne = C.c_uint64(), C.c_uint64()
    lib.dsim_synth_grid_sizes.argtypes = [C.c_uint32, C.c_uint32, u64p, u64p]
    rc = lib.dsim_synth_grid_sizes(W, H, C.byref(nn), C.byref(ne))
    assert rc == 0, rc
    n, e = nn.value, ne.value
    g = Graph(n_nodes=n, eco_off=np.zeros(n + 1, np.uint64), eco_id=np.zeros(n, np.uint32),
              edge_off=np.zeros(n + 1, np.uint64), edge_dst=np.zeros(e, np.uint32),
              edge_sec=np.zeros(e, np.float64), popcap=np.zeros(n, np.float64), h3=np.zeros(n, np.uint64),
              lon=np.zeros(n, np.float64), lat=np.zeros(n, np.float64))
    res = np.zeros(n, np.float64)
    mx = np.zeros(n, np.float64)
    lib.dsim_synth_grid.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64, C.c_double, u64p, f64p, f64p, u64p, u32p,
                                    f64p, u64p, u32p, f64p, f64p, f64p]
    rc = lib.dsim_synth_grid(W, H, seed, recovery, _ptr(g.h3, C.c_uint64), _ptr(g.lon, C.c_double),
                             _ptr(g.lat, C.c_double), _ptr(g.eco_off, C.c_uint64), _ptr(g.eco_id, C.c_uint32),
                             _ptr(g.popcap, C.c_double), _ptr(g.edge_off, C.c_uint64), _ptr(g.edge_dst, C.c_uint32),
                             _ptr(g.edge_sec, C.c_double), _ptr(res, C.c_double), _ptr(mx, C.c_double))
    assert rc == 0, rc
    return g, res, mx


def synth_families(W: int, H: int, n: int, seed: int = 1, n_occupied: int = 0, hist_fill: int = 0,
                   lib: Optional[C.CDLL] = None) -> Families:
    lib = lib or load_synth()
    f = Families.empty(n, n * hist_fill, 0)
    c = f.as_c()
    lib.dsim_synth_families.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32,
                                        C.POINTER(FamiliesC)]
    rc = lib.dsim_synth_families(W, H, seed, n, n_occupied, hist_fill, C.byref(c))
    assert rc == 0, rc
    f.hist_patch, f.hist_harvest = f.hist_patch[:n * hist_fill], f.hist_harvest[:n * hist_fill]
    f.adapt_eco, f.adapt_val = f.adapt_eco[:0], f.adapt_val[:0]
    return f


MG_PHASE_LIFECYCLE, MG_PHASE_MOVE, MG_PHASE_PATCH, MG_PHASE_FINISH, MG_PHASE_HALO_PACK, MG_PHASE_HALO_UNPACK = range(6)
MG_CH_SPLIT, MG_CH_EMIGRANTS, MG_CH_HALO = range(3)


def mg_unique_id(lib: C.CDLL) -> bytes:
    buf = C.create_string_buffer(128)
    rc = lib.dsim_mg_unique_id(buf)
    if rc != 0:
        raise DsimError(rc, "dsim_mg_unique_id failed")
    return buf.raw


def mg_band_bounds(location: np.ndarray, n_nodes: int, world: int, halo_width: int) -> np.ndarray:
    """Node boundaries of `world` contiguous bands holding about the same number of families each
    (equal family counts rather than equal node counts), every band >= halo_width nodes."""
    loc = np.sort(np.asarray(location, dtype=np.int64))
    b = [0]
    for r in range(1, world):
        q = int(loc[min(len(loc) - 1, (len(loc) * r) // world)]) if len(loc) else (n_nodes * r) // world
        q = max(q, b[-1] + halo_width)
        q = min(q, n_nodes - (world - r) * halo_width)
        b.append(q)
    b.append(n_nodes)
    out = np.array(b, dtype=np.uint32)
    if (np.diff(out.astype(np.int64)) < halo_width).any():
        raise ValueError("grid too small for that many bands")
    return out


def mg_exchange_in_process(sims, channel):
    """Move the messages of one channel between handles that live in this process ("virtual ranks")."""
    world = len(sims)
    if channel == MG_CH_SPLIT:
        for src in sims:
            sp, nb = src.mg_buffer(channel, 0, False)
            for dst in sims:
                dp, _ = dst.mg_buffer(channel, src.mg_rank, True)
                dst.mg_memcpy(dp, sp, nb)
        return
    for r, s in enumerate(sims):
        if r + 1 < world:   # up: my upper send buffer -> the upper neighbour's lower recv buffer
            sp, nb = s.mg_buffer(channel, 1, False)
            dp, _ = sims[r + 1].mg_buffer(channel, 0, True)
            sims[r + 1].mg_memcpy(dp, sp, nb)
        if r > 0:           # down
            sp, nb = s.mg_buffer(channel, 0, False)
            dp, _ = sims[r - 1].mg_buffer(channel, 1, True)
            sims[r - 1].mg_memcpy(dp, sp, nb)


def mg_step_in_process(sims):
    """One season on virtual ranks: the four phases with the three exchanges in between."""
    for s in sims:
        s.mg_phase(MG_PHASE_LIFECYCLE)
    mg_exchange_in_process(sims, MG_CH_SPLIT)
    for s in sims:
        s.mg_phase(MG_PHASE_MOVE)
    mg_exchange_in_process(sims, MG_CH_EMIGRANTS)
    for s in sims:
        s.mg_phase(MG_PHASE_PATCH)
    mg_exchange_in_process(sims, MG_CH_HALO)
    for s in sims:
        s.mg_phase(MG_PHASE_FINISH)


def mg_halo_sync_in_process(sims):
    for s in sims:
        s.mg_phase(MG_PHASE_HALO_PACK)
    mg_exchange_in_process(sims, MG_CH_HALO)
    for s in sims:
        s.mg_phase(MG_PHASE_HALO_UNPACK)
