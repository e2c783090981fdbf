"""Batched gavin2017processbased model behind the C ABI of include/gavin.h.

`population_capacity` restates model.py:271-297 / Grid.__init__ (model.py:209-214) on the host (numpy), so that the
`pow` of the capacity formula is the caller's; `run_sweep` hands the whole sweep to `gavin_run` (one warp per member).
There is no CPU fallback: without a device `gavin_run` fails with DSIM_ERR_NO_DEVICE.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

AREA = 450_000_000.0  # m^2, model.py:127


class Config(C.Structure):
    _fields_ = [("rows", C.c_uint32), ("cols", C.c_uint32), ("n_members", C.c_uint32), ("n_lang_caps", C.c_uint32),
                ("max_languages", C.c_uint32), ("device", C.c_uint32)]


class Result(C.Structure):
    _fields_ = [("device_ms", C.c_double), ("grow_calls", C.c_uint64), ("cell_updates", C.c_uint64)]


@dataclass
class Sweep:
    language: np.ndarray      # int32 [members, cells], 0 = no language
    population: np.ndarray    # int32 [members, cells]
    n_languages: np.ndarray   # uint32 [members]
    device_ms: float
    grow_calls: int
    cell_updates: int


def population_capacity(precipitation: np.ndarray, alpha, beta) -> np.ndarray:
    """K = alpha * P ** beta individuals / km^2 (model.py:271-297) times the cell area in km^2 (model.py:209-214);
    `alpha`, `beta` scalars or [members, 1] arrays -> [cells] or [members, cells]."""
    return alpha * np.power(precipitation, beta) * AREA / 1000000


def run_sweep(lib: C.CDLL, rows: int, cols: int, popcap: np.ndarray, growth_factor: np.ndarray, seed: np.ndarray,
              lang_caps: np.ndarray, max_languages: int = 0, device: int = 0) -> Sweep:
    popcap = np.ascontiguousarray(popcap, np.float64)
    S, n = popcap.shape
    assert n == 2 * rows * cols
    gf = np.ascontiguousarray(growth_factor, np.float64)
    sd = np.ascontiguousarray(seed, np.uint64)
    caps = np.ascontiguousarray(lang_caps, np.float64)
    assert gf.shape == (S,) and sd.shape == (S,)
    lang = np.zeros((S, n), np.int32)
    pop = np.zeros((S, n), np.int32)
    nl = np.zeros(S, np.uint32)
    cfg = Config(rows, cols, S, len(caps), max_languages, device)
    res = Result()
    f = lib.gavin_run
    f.restype = C.c_int
    f.argtypes = [C.POINTER(Config)] + [C.c_void_p] * 7 + [C.POINTER(Result)]
    rc = f(C.byref(cfg), popcap.ctypes.data, gf.ctypes.data, sd.ctypes.data, caps.ctypes.data, lang.ctypes.data, pop.ctypes.data,
           nl.ctypes.data, C.byref(res))
    if rc != 0:
        lib.gavin_last_error.restype = C.c_char_p
        raise RuntimeError(f"gavin_run failed ({rc}): {lib.gavin_last_error().decode()}")
    return Sweep(lang, pop, nl, res.device_ms, int(res.grow_calls), int(res.cell_updates))
