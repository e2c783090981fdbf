"""Shared helpers of the test-suite: library loading and state comparison.

Only the tests (and __graft_entry__.smoke / bench.py's CPU legs) load oracle/liboracle.so.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "dispersal-simulation_b200", "python"))
import dsim_b200 as D  # noqa: E402

ORACLE_PATH = os.path.join(REPO, "oracle", "liboracle.so")
EMU_PATH = os.path.join(REPO, "tests", "emu", "libdsim_emu.so")


def _make(directory, target=None):
    cmd = ["make", "-s", "-C", directory] + ([target] if target else [])
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)


def load_oracle() -> C.CDLL:
    _make(os.path.join(REPO, "oracle"))
    lib = C.CDLL(ORACLE_PATH)
    lib.oracle_set_mode.argtypes = [C.c_void_p, C.c_int]
    return lib


def load_emu() -> C.CDLL:
    """The CPU debugging build of the CUDA sources (tests/emu); never used by the product."""
    _make(os.path.join(REPO, "tests", "emu"))
    return C.CDLL(EMU_PATH)


def load_emu_small() -> C.CDLL:
    """Same, compiled with tiny capacities in k_move_blk (tests/emu/Makefile) so that the multi-round paths run."""
    _make(os.path.join(REPO, "tests", "emu"))
    return C.CDLL(os.path.join(REPO, "tests", "emu", "libdsim_emu_small.so"))


def load_synth() -> C.CDLL:
    _make(os.path.join(REPO, "dispersal-simulation_b200"), "synth")
    return D.load_synth()


def oracle_sim(lib, graph, params=None, seed=1, mode=1):
    s = D.Sim(lib, graph, params=params, seed=seed, prefix="oracle_")
    lib.oracle_set_mode(s.h, mode)
    return s


def default_params(lib, prefix="oracle_") -> D.Params:
    p = D.Params()
    getattr(lib, prefix + "default_params")(C.byref(p))
    return p


INT_FIELDS = ["id", "parent_id", "birth_index", "culture", "location", "size", "till_adult", "till_mutation",
              "has_mutation_clock", "n_offspring"]
FLOAT_FIELDS = ["stored", "last_harvest"]


def compare_families(a: D.Families, b: D.Families, rtol=0.0, hist_cap=None, what=""):
    """a = oracle, b = device.  Integer state bit-exact; floats within rtol (0 = bit-exact)."""
    msgs = []
    if a.n != b.n:
        return [f"{what}: family count {a.n} vs {b.n}"]
    for name in INT_FIELDS:
        x, y = getattr(a, name), getattr(b, name)
        bad = np.nonzero(x != y)[0]
        if len(bad):
            i = int(bad[0])
            msgs.append(f"{what}: {name} differs at {len(bad)} families, first i={i} id={int(a.id[i])}: {x[i]} vs {y[i]}")
    for name in FLOAT_FIELDS:
        x, y = getattr(a, name), getattr(b, name)
        if rtol == 0.0:
            bad = np.nonzero(x.view(np.uint64) != y.view(np.uint64))[0]
        else:
            bad = np.nonzero(~np.isclose(x, y, rtol=rtol, atol=0.0, equal_nan=True))[0]
        if len(bad):
            i = int(bad[0])
            msgs.append(f"{what}: {name} differs at {len(bad)} families, first i={i} id={int(a.id[i])}: {x[i]!r} vs {y[i]!r}")
    if a.hist_off is not None and b.hist_off is not None:
        for i in range(a.n):
            pa, ha = a.history_of(i)
            pb, hb = b.history_of(i)
            if hist_cap is not None:
                pa, ha = pa[-hist_cap:], ha[-hist_cap:]
            if len(pa) != len(pb) or (pa != pb).any() or not np.array_equal(ha, hb, equal_nan=True) and not (
                    rtol > 0 and np.allclose(ha, hb, rtol=rtol, atol=0, equal_nan=True)):
                msgs.append(f"{what}: history of family i={i} id={int(a.id[i])} differs: len {len(pa)} vs {len(pb)}")
                break
    if a.adapt_off is not None and b.adapt_off is not None:
        for i in range(a.n):
            ea, va = a.adaptation_of(i)
            eb, vb = b.adaptation_of(i)
            if len(ea) != len(eb) or (ea != eb).any() or not (np.array_equal(va, vb) or (rtol > 0 and np.allclose(va, vb, rtol=rtol, atol=0))):
                msgs.append(f"{what}: adaptation of family i={i} id={int(a.id[i])} differs: {list(zip(ea, va))} vs {list(zip(eb, vb))}")
                break
    return msgs


def compare_sims(oracle: D.Sim, dev: D.Sim, rtol=0.0, hist_cap=None, what=""):
    msgs = compare_families(oracle.get_families(), dev.get_families(), rtol=rtol, hist_cap=hist_cap, what=what)
    ra, _ = oracle.get_patches()
    rb, _ = dev.get_patches()
    if rtol == 0.0:
        bad = np.nonzero(ra.view(np.uint64) != rb.view(np.uint64))[0]
    else:
        bad = np.nonzero(~np.isclose(ra, rb, rtol=rtol, atol=0.0))[0]
    if len(bad):
        k = int(bad[0])
        msgs.append(f"{what}: patch resources differ at {len(bad)} entries, first {k}: {ra[k]!r} vs {rb[k]!r}")
    na, ca, sa = oracle.get_storage()
    nb, cb, sb = dev.get_storage()
    if len(na) != len(nb) or (na != nb).any() or (ca != cb).any() or (sa != sb).any():
        msgs.append(f"{what}: storage differs: {len(na)} vs {len(nb)} entries")
    if oracle.get_time() != dev.get_time():
        msgs.append(f"{what}: t differs: {oracle.get_time()} vs {dev.get_time()}")
    return msgs


def compare_sims_bulk(oracle: D.Sim, dev: D.Sim, what=""):
    """compare_sims for states of 10^5..10^6 families: whole-array comparisons only (no per-family Python loop), all
    bit-exact.  Requires that the device's history ring has not wrapped (history_capacity >= the longest history),
    so that both sides report the same entries."""
    msgs = []
    a, b = oracle.get_families(), dev.get_families()
    if a.n != b.n:
        return [f"{what}: family count {a.n} vs {b.n}"]
    for name in INT_FIELDS + FLOAT_FIELDS:
        x, y = getattr(a, name), getattr(b, name)
        bad = np.nonzero(x.view(np.uint8).reshape(len(x), -1) != y.view(np.uint8).reshape(len(y), -1))[0]
        if len(bad):
            i = int(bad[0])
            msgs.append(f"{what}: {name} differs at {len(np.unique(bad))} families, first i={i} id={int(a.id[i])}: {x[i]!r} vs {y[i]!r}")
    for off, cols in (("hist_off", ("hist_patch", "hist_harvest")), ("adapt_off", ("adapt_eco", "adapt_val"))):
        xo, yo = getattr(a, off), getattr(b, off)
        if len(xo) != len(yo) or (xo != yo).any():
            msgs.append(f"{what}: {off} differs")
            continue
        n = int(xo[-1])
        for c in cols:
            x, y = getattr(a, c)[:n], getattr(b, c)[:n]
            if x.dtype == np.float64:
                bad = np.nonzero(x.view(np.uint64) != y.view(np.uint64))[0]
            else:
                bad = np.nonzero(x != y)[0]
            if len(bad):
                k = int(bad[0])
                i = int(np.searchsorted(xo, k, side="right") - 1)
                msgs.append(f"{what}: {c} differs at {len(bad)} entries, first in family i={i} id={int(a.id[i])}: {x[k]!r} vs {y[k]!r}")
    ra, _ = oracle.get_patches()
    rb, _ = dev.get_patches()
    bad = np.nonzero(ra.view(np.uint64) != rb.view(np.uint64))[0]
    if len(bad):
        k = int(bad[0])
        msgs.append(f"{what}: patch resources differ at {len(bad)} entries, first {k}: {ra[k]!r} vs {rb[k]!r}")
    na, ca, sa = oracle.get_storage()
    nb, cb, sb = dev.get_storage()
    if len(na) != len(nb) or (na != nb).any() or (ca != cb).any() or (sa != sb).any():
        msgs.append(f"{what}: storage differs: {len(na)} vs {len(nb)} entries")
    if oracle.get_time() != dev.get_time():
        msgs.append(f"{what}: t differs: {oracle.get_time()} vs {dev.get_time()}")
    return msgs
