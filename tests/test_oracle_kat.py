"""Known-answer tests that pin the oracle (the reference ships none: SURVEY.md §4, §8c).

The micro-vectors are derived by hand from the cited lines of
supplement/dispersal_model_rust/src/lib.rs with the default parameters (season = 1/6).
"""
import ctypes as C
import math

import numpy as np
import pytest

SEASON = 1.0 / 6.0


def _philox(lib, ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib.oracle_philox4x32_10(c, k, o)
    return list(o)


def test_philox4x32_10_random123_vectors(oracle):
    # Random123 kat_vectors, "philox4x32 10"
    assert _philox(oracle, [0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert _philox(oracle, [0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _philox(oracle, [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == [
        0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_constants(oracle):
    adult, grow, n_mem = C.c_uint32(), C.c_uint32(), C.c_uint32()
    rng, cps, decay = C.c_double(), C.c_double(), C.c_double()
    oracle.oracle_constants.argtypes = [C.c_double] + [C.c_void_p] * 6
    oracle.oracle_constants(SEASON, C.byref(adult), C.byref(grow), C.byref(rng), C.byref(cps), C.byref(decay), C.byref(n_mem))
    assert adult.value == 91          # (15/season) as u32 + 1, lib.rs:535,1878
    assert grow.value == 17           # (2.9296875*6) as u32, lib.rs:1910-1916
    assert rng.value == pytest.approx(68899.8098, rel=1e-8)       # lib.rs:1183
    assert cps.value == pytest.approx(3.0 / 31556925.216, rel=1e-12)  # lib.rs:100,1139-1140
    assert decay.value == pytest.approx(0.5 ** (1 / 12), rel=1e-15)   # lib.rs:847
    # memory >= f64::EPSILON = 2^-52 while n <= 624 (0.5^(n/12) = 2^-52 at n = 624)
    assert n_mem.value in (624, 625, 626)


def test_use_resources_and_maybe_shrink(oracle):
    # lib.rs:1923-1937: size 5 needs 5/6 = 0.8333 > 0.5 -> shrinks to 3 (3/6 = 0.5), stored 0.0
    oracle.oracle_use_resources_and_maybe_shrink.argtypes = [C.c_void_p, C.c_void_p, C.c_double]
    size, stored = C.c_uint64(5), C.c_double(0.5)
    shrunk = oracle.oracle_use_resources_and_maybe_shrink(C.byref(size), C.byref(stored), SEASON)
    assert shrunk == 1 and size.value == 3 and stored.value == 0.0
    size, stored = C.c_uint64(4), C.c_double(10.0)
    shrunk = oracle.oracle_use_resources_and_maybe_shrink(C.byref(size), C.byref(stored), SEASON)
    assert shrunk == 0 and size.value == 4 and stored.value == 10.0 - 4 * SEASON
    # nothing stored at all: the family starves to size 0
    size, stored = C.c_uint64(3), C.c_double(0.0)
    shrunk = oracle.oracle_use_resources_and_maybe_shrink(C.byref(size), C.byref(stored), SEASON)
    assert shrunk == 1 and size.value == 0 and stored.value == 0.0


def test_maybe_grow(oracle):
    oracle.oracle_maybe_grow.argtypes = [C.c_void_p, C.c_void_p, C.c_double]
    size, till = C.c_uint32(4), C.c_uint32(0)
    oracle.oracle_maybe_grow(C.byref(size), C.byref(till), SEASON)
    assert (size.value, till.value) == (5, 17)       # lib.rs:1913-1916
    size, till = C.c_uint32(4), C.c_uint32(3)
    oracle.oracle_maybe_grow(C.byref(size), C.byref(till), SEASON)
    assert (size.value, till.value) == (4, 2)        # lib.rs:1918


def test_recover(oracle):
    oracle.oracle_recover.restype = C.c_double
    oracle.oracle_recover.argtypes = [C.c_double] * 4
    # lib.rs:2100-2108 with U = 0.25: min(2, 2*0.5*(2*0.25) + 1) = 1.5
    assert oracle.oracle_recover(1.0, 2.0, 0.5, 0.25) == 1.5
    # the stale doctest of lib.rs:2087-2092 (1.4375) describes an older four-argument function
    assert oracle.oracle_recover(1.0, 2.0, 0.5, 0.25) != 1.4375
    assert oracle.oracle_recover(1.9, 2.0, 0.5, 0.9) == 2.0     # clipped at the maximum
    assert oracle.oracle_recover(2.0, 2.0, 0.5, 0.9) == 2.0     # res == max: untouched
    assert oracle.oracle_recover(3.0, 2.0, 0.5, 0.9) == 3.0     # res > max: untouched


def test_best_location(oracle):
    oracle.oracle_best_location.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    # (gain, cost, U) = (4, .01, .5), (10, .02, .1): g1 = 1.99 accepted, m = 3.99; g2 = .98 rejected
    gain = (C.c_double * 2)(4.0, 10.0)
    cost = (C.c_double * 2)(0.01, 0.02)
    u = (C.c_double * 2)(0.5, 0.1)
    tc = C.c_double()
    assert oracle.oracle_best_location(2, gain, cost, u, C.byref(tc)) == 0 and tc.value == 0.01
    # threshold moves to the un-noised gain: a later, better-noised but smaller gain cannot win
    gain = (C.c_double * 3)(4.0, 3.9, 8.0)
    cost = (C.c_double * 3)(0.0, 0.0, 0.5)
    u = (C.c_double * 3)(0.5, 0.99, 0.6)
    assert oracle.oracle_best_location(3, gain, cost, u, C.byref(tc)) == 2 and tc.value == 0.5
    # nothing beats staying: target = null (-1), cost 0
    gain = (C.c_double * 1)(0.0)
    cost = (C.c_double * 1)(0.1)
    u = (C.c_double * 1)(0.9)
    assert oracle.oracle_best_location(1, gain, cost, u, C.byref(tc)) == -1 and tc.value == 0.0


def test_exploit_patch_single_family(oracle):
    # one band, one family of 5, fresh adaptation 0.5, res 100, cap 1e50
    oracle.oracle_exploit_one.restype = C.c_double
    oracle.oracle_exploit_one.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double,
                                          C.c_void_p]
    size = (C.c_uint32 * 1)(5)
    adapt = (C.c_double * 1)(0.5)
    band = (C.c_int * 1)(0)
    lh = (C.c_double * 1)()
    left = oracle.oracle_exploit_one(1, size, adapt, band, 100.0, 1e50, SEASON, lh)
    a = max(0.5, 0.5 * 0.95) + 0.05                     # 0.55, lib.rs:1953-1954
    effort = 5 * a                                      # 2.75
    total = 1.0 + effort * effort                       # 8.5625
    harvest = 100.0 * (effort * effort) / total         # 88.32...
    coef = min(harvest / effort, 2.5)                   # 2.5
    assert lh[0] == coef * SEASON * effort              # 1.1458333
    assert lh[0] == pytest.approx(1.1458333, rel=1e-7)
    assert left == 100.0 - coef * effort                # 93.125
    assert left == 93.125


def test_exploit_patch_two_bands_share_by_squared_effort(oracle):
    size = (C.c_uint32 * 3)(4, 2, 3)
    adapt = (C.c_double * 3)(1.0, 1.0, 1.0)      # stays 1.0: max(.5, .95) + .05
    band = (C.c_int * 3)(0, 0, 1)
    lh = (C.c_double * 3)()
    res = 3.0                                      # scarce: coefficient below the 2.5 cap
    left = oracle.oracle_exploit_one(3, size, adapt, band, res, 1e50, SEASON, lh)
    e0, e1 = 6.0, 3.0
    total = 1.0 + e0 * e0 + e1 * e1
    h0 = res * (e0 * e0) / total
    c0 = min(h0 / e0, 2.5)
    res1 = res - c0 * e0
    h1 = res1 * (e1 * e1) / total                  # the second band sees what the first left (lib.rs:2062)
    c1 = min(h1 / e1, 2.5)
    assert lh[0] == c0 * SEASON * 4.0 and lh[1] == c0 * SEASON * 2.0 and lh[2] == c1 * SEASON * 3.0
    assert left == res1 - c1 * e1


def test_adaptation_entry(oracle):
    oracle.oracle_adapt_touch.restype = C.c_double
    oracle.oracle_adapt_touch.argtypes = [C.c_double, C.c_double]
    assert oracle.oracle_adapt_touch(0.5, 0.5) == 0.55
    assert oracle.oracle_adapt_touch(1.0, 0.5) == 1.0 * 0.95 + 0.05
    v = 0.5
    for _ in range(200):                           # converges to the fixed point 1.0 from below
        v = oracle.oracle_adapt_touch(v, 0.5)
    assert 0.999 < v <= 1.0


def test_log_and_mutation_clock(oracle):
    oracle.oracle_log.restype = C.c_double
    oracle.oracle_log.argtypes = [C.c_double]
    rng = np.random.default_rng(1)
    xs = list(rng.random(5000)) + list(np.exp(rng.uniform(-700, 700, 5000))) + [0.5, 2.0, 1 - 6e-3, 5e-324, 1e-310]
    for x in xs:
        assert oracle.oracle_log(x) == pytest.approx(math.log(x), rel=2e-15, abs=1e-300)
    assert oracle.oracle_log(1.0) == 0.0
    assert oracle.oracle_log(0.0) == -math.inf
    assert math.isnan(oracle.oracle_log(-1.0))
    oracle.oracle_time_till_next_mutation.restype = C.c_uint32
    oracle.oracle_time_till_next_mutation.argtypes = [C.c_double, C.c_double]
    # lib.rs:1270-1272: ln(U)/ln(1-rate) truncated; U = 0 saturates (ln 0 = -inf)
    assert oracle.oracle_time_till_next_mutation(0.0, 6e-3) == 0xFFFFFFFF
    assert oracle.oracle_time_till_next_mutation(0.999999, 6e-3) == 0
    assert oracle.oracle_time_till_next_mutation(0.5, 6e-3) == int(math.log(0.5) / math.log(1 - 6e-3))
    oracle.oracle_f64_as_u32.restype = C.c_uint32
    oracle.oracle_f64_as_u32.argtypes = [C.c_double]
    assert oracle.oracle_f64_as_u32(float("nan")) == 0
    assert oracle.oracle_f64_as_u32(-3.5) == 0
    assert oracle.oracle_f64_as_u32(17.578125) == 17
    assert oracle.oracle_f64_as_u32(1e20) == 0xFFFFFFFF


def test_will_cooperate_against_the_reference_analysis_code(oracle):
    """`emergence::will_cooperate` is missing from lib.rs; the reference's plotting code evaluates the predicate as
    `cultural_distance(c1, c2) < cooperation_threshold` (plot.py:100-101, 138).  tests/golden/cooperation.npz holds the
    outputs of those two expressions, executed from plot.py's syntax tree (make_cooperation_golden.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cooperation.npz"))
    oracle.oracle_will_cooperate.restype = C.c_int
    oracle.oracle_will_cooperate.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int]
    c1, c2, dist, coop = g["c1"], g["c2"], g["distance"], g["cooperate"]
    assert dist.max() > 20 and (np.bincount(dist)[:12] > 100).all()
    for k, t in enumerate(g["thresholds"]):
        strict = np.array([oracle.oracle_will_cooperate(int(a), int(b), int(t), 0) for a, b in zip(c1, c2)], bool)
        assert (strict == coop[:, k]).all(), f"threshold {t}"
        incl = np.array([oracle.oracle_will_cooperate(int(a), int(b), int(t), 1) for a, b in zip(c1, c2)], bool)
        assert (incl == (dist <= t)).all(), f"threshold {t}, inclusive"
