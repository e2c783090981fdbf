#!/usr/bin/env python
"""Golden vectors for the gavin2017processbased oracle from the REFERENCE's own functions.

Runs in the build container only (needs /root/reference): imports supplement/gavin2017processbased/language.py with a
stub `popcaps` module (the real one reads a Binford table that is not in the repository) and records
* `dhondt(seats, votes)`                                       (language.py:86-99)
* `Language.distribute(growth, candidates)`                    (language.py:46-84) with the candidate set given as a
  list (the function indexes `list(candidates)`) and numpy.random.random replaced by a recorded sequence of sort keys
on random inputs.  Output: tests/golden/gavin_functions.npz, checked by tests/test_gavin.py against oracle/gavin.py.
    python tests/golden/make_gavin_golden.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/supplement/gavin2017processbased"
stub = types.ModuleType("popcaps")
stub.random_popcap = lambda: 100.0
sys.modules["popcaps"] = stub
sys.path.insert(0, REF)
import language  # noqa: E402  (the reference)


class Cell:
    def __init__(self, popcap, population):
        self.popcap, self.population = popcap, population


def main():
    rng = np.random.default_rng(2017)
    n_cases, width = 400, 7
    seats = np.zeros(n_cases)
    votes = np.full((n_cases, width), np.nan)
    dh = np.full((n_cases, width), -1, np.int64)
    growth = np.zeros(n_cases)
    popcap = np.full((n_cases, width), np.nan)
    population = np.zeros((n_cases, width), np.int64)
    keys = np.zeros((n_cases, width), np.uint64)
    got = np.full((n_cases, width), -1, np.int64)     # -1 = the loop did not reach the candidate
    left = np.zeros(n_cases)
    visit = np.full((n_cases, width), -1, np.int64)   # candidates in the insertion order of the returned Counter
    lang = object.__new__(language.Language)
    for k in range(n_cases):
        n = int(rng.integers(1, width + 1))
        pc = np.round(rng.uniform(0.2, 60, n), 3)
        pop = rng.integers(0, 50, n)
        if k % 5 == 0:
            pop = np.minimum(pop, pc.astype(np.int64))             # mostly room to grow
        g = float(np.round(rng.uniform(-2, 25), 4)) if k % 7 else float(rng.integers(0, 12))
        # dhondt on the growth space of the case
        space = [float(a) - int(b) for a, b in zip(pc, pop)]
        seats[k] = g
        votes[k, :n] = space
        dh[k, :n] = language.dhondt(g, space)
        # distribute with recorded sort keys
        ks = rng.integers(0, 2 ** 63, n, dtype=np.uint64)
        if k % 11 == 0 and n > 2:
            ks[2] = ks[1]                                          # a tie: the earlier candidate comes first
        seq = iter(int(x) for x in ks[1:])
        language.numpy.random.random = lambda: next(seq)
        cells = [Cell(float(a), int(b)) for a, b in zip(pc, pop)]
        counter, rest = lang.distribute(g, cells)
        growth[k], left[k] = g, rest
        popcap[k, :n], population[k, :n], keys[k, :n] = pc, pop, ks
        for i, c in enumerate(cells):
            if c in counter:
                got[k, i] = counter[c]
        for pos, c in enumerate(counter.keys()):
            visit[k, pos] = cells.index(c)
    np.savez_compressed(os.path.join(HERE, "gavin_functions.npz"), seats=seats, votes=votes, dhondt=dh, growth=growth,
                        popcap=popcap, population=population, keys=keys, got=got, left=left, visit=visit)
    print("gavin_functions.npz:", n_cases, "cases")


if __name__ == "__main__":
    main()
