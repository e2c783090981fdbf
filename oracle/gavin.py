"""CPU oracle of the gavin2017processbased model — TEST INFRASTRUCTURE ONLY.

A restatement of `supplement/gavin2017processbased` of the reference (model.py, language.py, popcaps.py): languages
are seeded one after the other on a hexagonal grid; a language grows cell by cell — logistic growth of every owned
cell over its neighbourhood, the integer growth distributed over the neighbourhood by D'Hondt's method — until the
population cap drawn for it is used up or nothing grows any more; then the next language is seeded in a free
neighbouring cell.  Only tests/, tools/bench_gavin.py's CPU leg and __graft_entry__.smoke() may import this module.

PARITY: `dhondt` and `distribute` are pinned against the reference's own functions (tests/golden/make_gavin_golden.py
imports language.py with a stub `popcaps` module; tests/golden/gavin_*.npz), `neighbors_table` and `Member.neighbors`
against `Grid.gridcell` / `Grid.neighbors` taken out of model.py's syntax tree (make_gavin_grid_golden.py).  Above that level the reference is not
reproducible (iteration order of Python sets of cell objects, the global numpy generator, inputs that are not in
the repository: a WorldClim raster at an absolute path, a Binford table outside the tree), so — as for the
dispersal model — the orders and draws are canonical:

* cells are numbered `m * R * C + i * C + j` (two interleaved rectangular sub-grids m = 0, 1 of R x C cells,
  model.py:62-124); `neighbors` in the order of model.py:236-262, out-of-range neighbours replaced by the cell itself;
* `self.cells` of a language iterates in insertion order; the candidate list of `distribute` is [cell, neighbours in
  list order, without repeats] (the reference: a set); `free` of the seeding step ascends by cell number;
* draws: Philox4x32-10 keyed by the sweep member's seed, counter = (tag, language id, grow call, 8 * cell ordinal +
  candidate), tags START / LANGCAP / PERM / SEED; a permutation sorts by (draw, candidate index).
"""
from __future__ import annotations

import numpy as np

AREA = 450_000_000.0  # m^2, model.py:127
TAG_START, TAG_LANGCAP, TAG_PERM, TAG_SEED = 1, 2, 3, 4
M32 = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c3 ^ k1) & M32, p0 & M32
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return c0, c1, c2, c3


def draw(seed, tag, a, b, c):
    """four 32-bit words of the member's stream"""
    return philox4x32_10((tag & M32, a & M32, b & M32, c & M32), (seed & M32, (seed >> 32) & M32))


def bounded(word64, n):
    """uniform integer in [0, n): multiply-high of a 64-bit draw (numpy.random.randint in the reference)"""
    return (word64 * n) >> 64


def neighbors_table(R, C):
    """model.py:236-262: the six neighbours of every cell, `self` where the grid ends -> int32 [2*R*C, 6]"""
    nb = np.zeros((2 * R * C, 6), np.int32)

    def cell(m, i, j, me):
        if i < 0 or i >= R or j < 0 or j >= C:
            return me
        return m * R * C + i * C + j
    for i in range(R):
        for j in range(C):
            me = i * C + j
            nb[me] = [cell(0, i, j + 1, me), cell(1, i, j, me), cell(1, i, j - 1, me), cell(0, i, j - 1, me),
                      cell(1, i - 1, j - 1, me), cell(1, i - 1, j, me)]
            me = R * C + i * C + j
            nb[me] = [cell(1, i, j + 1, me), cell(0, i, j + 1, me), cell(0, i, j, me), cell(1, i, j - 1, me),
                      cell(0, i + 1, j, me), cell(0, i + 1, j + 1, me)]
    return nb


def population_capacity(precipitation, alpha, beta):
    """model.py:271-297 and Grid.__init__ (model.py:209-214): K = alpha * P ** beta per km^2, times the cell area"""
    return alpha * np.power(precipitation, beta) * AREA / 1000000


def dhondt(seats_to_allocate, votes):
    """language.py:86-99.  Ties go to the earlier entry."""
    d = [1] * len(votes)
    result = [0] * len(votes)
    while sum(result) < seats_to_allocate:
        best, best_q = 0, votes[0] / d[0]
        for i in range(1, len(votes)):
            q = votes[i] / d[i]
            if q > best_q:
                best, best_q = i, q
        result[best] += 1
        d[best] += 1
    return result


def distribute(growth, space, order):
    """language.py:46-84 on a candidate list: `space[i]` = popcap - population of candidate i, `order` = the visiting
    order [0, permutation of 1..n-1].  Returns ([(candidate, its growth)] in visiting order — the insertion order of
    the reference's Counter — up to where the loop stops, the growth carried over)."""
    m = dhondt(growth, space)
    out = []
    for i in order:
        out.append((i, m[i]))
        growth -= m[i]
        if growth < 1:
            break
    return out, growth


class Member:
    """One sweep member: a grid state and the sequence of languages (model.py:333-378)."""

    def __init__(self, R, C, popcap, growth_factor, seed, lang_caps, max_languages=1 << 30):
        self.R, self.C, self.n = R, C, 2 * R * C
        self.nb = neighbors_table(R, C)
        self.popcap = np.asarray(popcap, np.float64)
        self.gf = float(growth_factor)
        self.seed = int(seed)
        self.lang_caps = np.asarray(lang_caps, np.float64)  # the Binford table of popcaps.py:4-17
        self.population = np.zeros(self.n, np.int64)
        self.language = np.zeros(self.n, np.int32)           # 0 = no language
        self.max_languages = max_languages
        self.n_languages = 0
        self.grow_calls = 0

    # --- Grid.neighbors(include_unlivable=False, include_foreign=False), model.py:236-262
    def neighbors(self, c, lang):
        return [int(q) for q in self.nb[c] if (self.language[q] == lang or self.language[q] == 0) and self.popcap[q] >= 1]

    def _lang_cap(self, lid):
        w = draw(self.seed, TAG_LANGCAP, lid, 0, 0)
        return float(self.lang_caps[bounded((w[1] << 32) | w[0], len(self.lang_caps))])  # random_popcap, popcaps.py:16-17

    def _start_language(self, lid, c):  # Language.__init__, language.py:9-16
        self.population[c] = max(1, min(10, int(self.popcap[c])))
        self.language[c] = lid
        return [c], self._lang_cap(lid)

    def _grow(self, lid, cells, cap, call):
        """Language.grow, language.py:18-44 -> (cap left, stop reason or None)"""
        grow_into = {}
        growth = 0
        for ordinal, cell in enumerate(cells):
            nbs = self.neighbors(cell, self.language[cell])
            region = [cell]
            region_popcap = self.popcap[cell]
            region_population = int(self.population[cell])
            for q in nbs:                                    # sums count repeated neighbours, the region does not
                if q not in region:
                    region.append(q)
                region_popcap = region_popcap + self.popcap[q]
                region_population += int(self.population[q])
            growth = growth + self.gf * float(self.population[cell]) * (1 - float(region_population) / region_popcap)
            space = [self.popcap[q] - float(self.population[q]) for q in region]
            keys = []
            for i in range(1, len(region)):
                w = draw(self.seed, TAG_PERM, lid, call, 8 * ordinal + i)
                keys.append(((w[1] << 32) | w[0], i))
            order = [0] + [i for _, i in sorted(keys)]
            got, growth = distribute(growth, space, order)
            for i, g in got:                                 # grow_into += more_grow_into (Counter: positive counts only,
                if g > 0:                                    # new keys in the visiting order)
                    grow_into[region[i]] = grow_into.get(region[i], 0) + g
        if not grow_into:
            return cap, "No more growth"
        for cell, g in grow_into.items():
            self.population[cell] += g
            if self.language[cell] != lid:
                self.language[cell] = lid
                cells.append(cell)
            cap -= g
            if cap <= 0:
                return cap, "popcap reached"
        return cap, None

    def run(self):
        """model.py:333-378 without the plotting: returns the number of languages."""
        k = 0
        while True:  # Grid.random_cell until the cell can hold one individual, model.py:346-349
            w = draw(self.seed, TAG_START, k, 0, 0)
            m, i, j = bounded(w[0] << 32, 2), bounded(w[1] << 32, self.R), bounded(w[2] << 32, self.C)
            c = m * self.R * self.C + i * self.C + j
            k += 1
            if self.popcap[c] >= 1:
                break
            if k > 1 << 20:
                return 0
        lid = 1
        cells, cap = self._start_language(lid, c)
        while True:
            self.n_languages = lid
            call = 0
            while True:
                cap, stop = self._grow(lid, cells, cap, call)
                call += 1
                self.grow_calls += 1
                if stop is not None:
                    break
            if lid >= self.max_languages:
                break
            filled = self.language != 0
            zone = set()
            for g in np.flatnonzero(filled):
                for q in self.neighbors(int(g), self.language[g]):
                    if self.popcap[q] > 0 and not filled[q]:
                        zone.add(q)
            if not zone:
                break
            free = sorted(zone)
            w = draw(self.seed, TAG_SEED, lid, 0, 0)
            lid += 1
            cells, cap = self._start_language(lid, free[bounded((w[1] << 32) | w[0], len(free))])
        return lid


def synthetic_precipitation(R, C, seed=1):
    """A smooth precipitation field (mm / year) with a dry belt, standing in for the WorldClim raster of model.py:277."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:R, 0:C]
    base = 900 + 700 * np.sin(2.2 * y / R + 0.3) * np.cos(1.7 * x / C) - 500 * np.exp(-((y - 0.55 * R) ** 2) / (0.02 * R * R + 1))
    p0 = np.clip(base + rng.normal(0, 60, (R, C)), 20, None)
    p1 = np.clip(base + rng.normal(0, 60, (R, C)) + 15, 20, None)
    return np.concatenate([p0.ravel(), p1.ravel()])


def synthetic_lang_caps(n=339, seed=2):
    """Stand-in for the TLPOP column of the Binford table (popcaps.py:4-17): log-normal group sizes."""
    rng = np.random.default_rng(seed)
    return np.round(np.clip(np.exp(rng.normal(6.3, 0.9, n)), 50, 20000))
