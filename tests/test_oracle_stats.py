"""Statistical pins of the canonical draws against the reference's formulas.

The reference draws from an unseeded thread-local `fastrand` and cannot be run here, so its trajectories cannot be
compared sample for sample; what CAN be checked is that every canonical draw of the oracle (and therefore of the
kernels, which are bit-identical to it) has the distribution the reference's code prescribes:

    lib.rs:847-866   history entry n is remembered with probability 0.5^(n * season / 2), entries independently
    lib.rs:1333-1335 shuffle: a uniformly random permutation
    lib.rs:1270-1272 mutation clock: floor(ln U / ln(1 - rate)), i.e. geometric with P(clock >= k) = (1 - rate)^k
    lib.rs:1277-1285 the flipped bit: uniform on 0..63
    lib.rs:1295-1298 pivot family: uniform on 0..len-1
    lib.rs:1309-1311 regrowth factor 2 U: uniform on [0, 2), mean 1
    lib.rs:1325-1327 fight_is_deadly: Bernoulli(fight_deadliness / 2^32), the two draws of an encounter independent
    lib.rs:1317-1319 resource_uncertainty: uniform on [0, 1)

Every test draws >= 10^5 samples per seed (seeds 1, 2, 3 of BASELINE.md) through the SAME functions the oracle's step
calls (oracle_draw_* exports) and states its bound: |z| < 4.5 for means and frequencies (false alarm 7e-6 per test
statistic), chi-square below the 1 - 1e-5 quantile.  CPU only, seconds."""
import ctypes as C
import itertools

import numpy as np
import pytest
from scipy import stats

SEEDS = [1, 2, 3]
Z = 4.5


@pytest.fixture(scope="module")
def lib(oracle):
    o = oracle
    o.oracle_draw_remembered.restype = C.c_uint32
    o.oracle_draw_remembered.argtypes = [C.c_double, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]
    o.oracle_draw_noise.restype = C.c_double
    o.oracle_draw_noise.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32]
    o.oracle_draw_shuffle_key.restype = C.c_uint64
    o.oracle_draw_shuffle_key.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
    o.oracle_draw_fight.restype = None
    o.oracle_draw_fight.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_int)]
    o.oracle_draw_pivot.restype = C.c_uint64
    o.oracle_draw_pivot.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64]
    o.oracle_draw_regrowth.restype = C.c_double
    o.oracle_draw_regrowth.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32]
    o.oracle_draw_mutation.restype = None
    o.oracle_draw_mutation.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    return o


def chi2_bound(dof):
    return stats.chi2.ppf(1 - 1e-5, dof)


def remembered_matrix(lib, seed, season, n_calls, length):
    kept = np.zeros((n_calls, length), np.uint8)
    n_mem = 0
    for k in range(n_calls):
        # (step, family id) pairs as the step uses them: many families in a few steps
        n_mem = lib.oracle_draw_remembered(season, seed, k % 50, 1000 + k // 50, length, kept[k].ctypes.data_as(C.c_void_p))
    return kept, n_mem


@pytest.mark.parametrize("seed", SEEDS)
def test_memory_keep_probability(lib, seed):
    """lib.rs:847-866 at season = 1/6: entry n is kept with probability 0.5^(n/12) (memory by repeated multiplication),
    nothing beyond the 625 entries with memory >= f64::EPSILON.  The head (n < 112) is checked entry by entry, the
    skip-sampled tail in bins (0.07 kept entries expected per family and step) and as a whole."""
    n_calls, length = 20000, 700      # 1.25e7 entry draws
    kept, n_mem = remembered_matrix(lib, seed, 1.0 / 6.0, n_calls, length)
    assert n_mem == 625
    decay = 0.5 ** (1.0 / 12.0)
    p = np.ones(n_mem)
    for n in range(1, n_mem):
        p[n] = p[n - 1] * decay
    k = kept.sum(axis=0).astype(np.float64)
    assert k[0] == n_calls                                  # memory_0 = 1: always remembered
    assert k[n_mem:].sum() == 0                             # lib.rs:856: the loop has stopped
    head = np.arange(1, 112)
    z = (k[head] - n_calls * p[head]) / np.sqrt(n_calls * p[head] * (1 - p[head]))
    assert np.abs(z).max() < Z, (int(head[np.abs(z).argmax()]), float(np.abs(z).max()))
    # tail: totals per bin against the sum of the probabilities (Poisson-binomial variance)
    edges = [112, 124, 136, 160, 200, 625]
    for a, b in zip(edges[:-1], edges[1:]):
        e, v = n_calls * p[a:b].sum(), n_calls * (p[a:b] * (1 - p[a:b])).sum()
        zz = (k[a:b].sum() - e) / np.sqrt(v)
        assert abs(zz) < Z, (a, b, zz, k[a:b].sum(), e)
    # the whole vector: mean number of remembered entries = sum of memory_n = 17.8 (SURVEY.md section 8d's n_h)
    per_call = kept.sum(axis=1)
    se = np.sqrt((p * (1 - p)).sum() / n_calls)
    assert abs(per_call.mean() - p.sum()) < Z * se
    assert abs(p.sum() - 17.817) < 0.01


@pytest.mark.parametrize("seed", SEEDS)
def test_memory_entries_are_independent(lib, seed):
    """The draws of lib.rs:859 are independent per entry: neighbouring entries of the head (same Philox block, adjacent
    bytes), entries across the head/tail boundary, and pairs inside the skip-sampled tail show no correlation."""
    n_calls = 20000
    kept, n_mem = remembered_matrix(lib, seed, 1.0 / 6.0, n_calls, 625)
    kept = kept.astype(np.float64)
    p = kept.mean(axis=0)
    for a, b in [(5, 6), (11, 12), (15, 16), (30, 31), (47, 48), (20, 40), (111, 112), (100, 113)]:
        cov = (kept[:, a] * kept[:, b]).mean() - p[a] * p[b]
        se = np.sqrt(p[a] * (1 - p[a]) * p[b] * (1 - p[b]) / n_calls)
        assert abs(cov) < Z * se, (a, b, cov, se)
    # inside the tail: number of families with two or more kept tail entries against the Poisson-binomial expectation
    decay = 0.5 ** (1.0 / 12.0)
    pt = decay ** np.arange(112, 625)
    p0 = np.prod(1 - pt)
    p1 = p0 * (pt / (1 - pt)).sum()
    p2 = 1 - p0 - p1
    tail_counts = kept[:, 112:].sum(axis=1)
    for observed, prob in [((tail_counts == 0).sum(), p0), ((tail_counts == 1).sum(), p1), ((tail_counts >= 2).sum(), p2)]:
        zz = (observed - n_calls * prob) / np.sqrt(n_calls * prob * (1 - prob))
        assert abs(zz) < Z, (observed, n_calls * prob)


def test_memory_other_seasons(lib):
    """The same law at other season lengths: a long season (head shorter than one Philox block per 16 entries allows)
    and a short one (1/24 year: the head is capped at 256 entries and the tail holds entries with up to 2.5 %)."""
    for season, n_calls in [(0.5, 20000), (1.0 / 24.0, 4000)]:
        decay = 0.5 ** (season / 2.0)
        mem, m = [], 1.0
        while not m < np.finfo(np.float64).eps:
            mem.append(m)
            m *= decay
        p = np.array(mem)
        kept, n_mem = remembered_matrix(lib, 3, season, n_calls, len(p) + 20)
        assert n_mem == len(p)
        k = kept.sum(axis=0).astype(np.float64)
        assert k[n_mem:].sum() == 0
        step = max(1, len(p) // 40)
        for a in range(1, len(p), step):      # bins of consecutive entries
            b = min(len(p), a + step)
            e, v = n_calls * p[a:b].sum(), n_calls * (p[a:b] * (1 - p[a:b])).sum()
            if e < 5:
                continue
            assert abs(k[a:b].sum() - e) < Z * np.sqrt(v), (season, a, b)
        assert abs(kept.sum() - n_calls * p.sum()) < Z * np.sqrt(n_calls * (p * (1 - p)).sum())


@pytest.mark.parametrize("seed", SEEDS)
def test_shuffle_is_a_uniform_permutation(lib, seed):
    """lib.rs:1333-1335 (Fisher-Yates) = a uniformly random permutation; canonically: order by a keyed 64-bit number.
    All 24 orders of four families are equally likely, and so is every rank for each of eight families."""
    n = 100_000
    ids = [17, 18, 4000, 123456789]
    counts = {}
    for t in range(n):
        keys = [lib.oracle_draw_shuffle_key(seed, t, i) for i in ids]
        order = tuple(np.argsort(keys))
        counts[order] = counts.get(order, 0) + 1
    assert len(counts) == 24
    c = np.array([counts[p] for p in itertools.permutations(range(4))], dtype=np.float64)
    chi = ((c - n / 24) ** 2 / (n / 24)).sum()
    assert chi < chi2_bound(23), chi
    ids8 = list(range(100, 108))
    rank = np.zeros((8, 8))
    m = 30_000
    for t in range(m):
        keys = np.array([lib.oracle_draw_shuffle_key(seed, t, i) for i in ids8], dtype=np.uint64)
        for pos, who in enumerate(np.argsort(keys)):
            rank[who, pos] += 1
    chi = ((rank - m / 8) ** 2 / (m / 8)).sum()
    assert chi < chi2_bound(49), chi          # 8x8 table with fixed margins: 49 degrees of freedom


@pytest.mark.parametrize("seed", SEEDS)
def test_mutation_clock_and_bit(lib, seed):
    """lib.rs:1270-1272: clock = floor(ln U / ln(1 - rate)) is geometric, P(clock >= k) = (1 - rate)^k, mean
    (1 - rate) / rate; both clocks of a step (first clock, clock after a flip) follow it and are independent;
    lib.rs:1277-1285: the flipped bit is uniform on 0..63."""
    n, rate = 100_000, 6e-3
    c = (C.c_uint32 * 2)()
    b = C.c_uint32()
    clocks = np.zeros((n, 2))
    bits = np.zeros(64)
    for k in range(n):
        lib.oracle_draw_mutation(seed, k % 100, 5000 + k // 100, rate, c, C.byref(b))
        clocks[k] = (c[0], c[1])
        bits[b.value] += 1
    mean, sd = (1 - rate) / rate, np.sqrt(1 - rate) / rate
    for j in range(2):
        assert abs(clocks[:, j].mean() - mean) < Z * sd / np.sqrt(n), (j, clocks[:, j].mean(), mean)
        for kk in (1, 50, 166, 500):          # survival function at a few points
            ps = (1 - rate) ** kk
            zz = ((clocks[:, j] >= kk).sum() - n * ps) / np.sqrt(n * ps * (1 - ps))
            assert abs(zz) < Z, (j, kk, zz)
    r = np.corrcoef(clocks[:, 0], clocks[:, 1])[0, 1]
    assert abs(r) < Z / np.sqrt(n)
    chi = ((bits - n / 64) ** 2 / (n / 64)).sum()
    assert chi < chi2_bound(63), chi


@pytest.mark.parametrize("seed", SEEDS)
def test_pivot_regrowth_fight_noise(lib, seed):
    n = 100_000
    # select_pivot_family, lib.rs:1295-1298: uniform on 0..len-1
    for length in (2, 7):
        c = np.bincount([lib.oracle_draw_pivot(seed, t % 200, 10 + t // 200, t % 3, length) for t in range(n)], minlength=length)
        chi = ((c - n / length) ** 2 / (n / length)).sum()
        assert chi < chi2_bound(length - 1), (length, chi)
    # recover_resources_proportion, lib.rs:1309-1311: 2 U, uniform on [0, 2), mean 1
    x = np.array([lib.oracle_draw_regrowth(seed, t % 200, 10 + t // 200, t % 2) for t in range(n)])
    assert x.min() >= 0.0 and x.max() < 2.0
    assert abs(x.mean() - 1.0) < Z * (2 / np.sqrt(12)) / np.sqrt(n)
    c = np.histogram(x, bins=20, range=(0, 2))[0]
    assert ((c - n / 20) ** 2 / (n / 20)).sum() < chi2_bound(19)
    # fight_is_deadly, lib.rs:1325-1327, twice per hostile encounter (lib.rs:1219,1226): independent Bernoulli(p)
    out = (C.c_int * 2)()
    for deadliness, p in ((1 << 31, 0.5), (1 << 29, 0.125)):
        t4 = np.zeros(4)
        for t in range(n):
            lib.oracle_draw_fight(seed, t % 200, 77 + t // 200, t % 5, deadliness, out)
            t4[2 * out[0] + out[1]] += 1
        e = n * np.array([(1 - p) ** 2, (1 - p) * p, p * (1 - p), p * p])
        assert ((t4 - e) ** 2 / e).sum() < chi2_bound(3), (deadliness, t4, e)
    # resource_uncertainty, lib.rs:1317-1319: uniform on [0, 1); candidates i and i + 32 share a Philox block
    u = np.array([[lib.oracle_draw_noise(seed, t % 100, 900 + t // 100, i) for i in (0, 32, 33, 64, 700)] for t in range(n // 2)])
    assert u.min() >= 0.0 and u.max() < 1.0
    for j in range(u.shape[1]):
        assert abs(u[:, j].mean() - 0.5) < Z / np.sqrt(12 * len(u))
        c = np.histogram(u[:, j], bins=16, range=(0, 1))[0]
        assert ((c - len(u) / 16) ** 2 / (len(u) / 16)).sum() < chi2_bound(15)
    assert abs(np.corrcoef(u[:, 0], u[:, 1])[0, 1]) < Z / np.sqrt(len(u))
    assert abs(np.corrcoef(u[:, 1], u[:, 2])[0, 1]) < Z / np.sqrt(len(u))
