#!/usr/bin/env python
"""Golden vectors for `will_cooperate` from the reference's own code.

`emergence::will_cooperate` is missing from lib.rs at this revision; the reference's analysis code evaluates the same
predicate as `cultural_distance(c1, c2) < params["cooperation_threshold"]` (supplement/plotting/plot.py:100-101, 138).
plot.py cannot be imported (matplotlib, cartopy, an OSM extract), so the two definitions are taken out of its syntax
tree and executed on their own: `cultural_distance` and the comparison used in `count_cultures`.

usage (in the build container, where /root/reference exists): python tests/golden/make_cooperation_golden.py
"""
import ast
import os
import warnings

import numpy as np

REF = "/root/reference/supplement/plotting/plot.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cooperation.npz")

with warnings.catch_warnings():
    warnings.simplefilter("ignore", SyntaxWarning)  # regular expressions without raw strings elsewhere in plot.py
    tree = ast.parse(open(REF).read())
fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "cultural_distance")
ns = {}
exec(compile(ast.Module(body=[fn], type_ignores=[]), REF, "exec"), ns)
cultural_distance = ns["cultural_distance"]
# the comparison of count_cultures (plot.py:138), taken from the tree as well
cc = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "count_cultures")
cmp = next(n for n in ast.walk(cc) if isinstance(n, ast.Compare) and isinstance(n.left, ast.Call)
           and getattr(n.left.func, "id", "") == "cultural_distance")
assert isinstance(cmp.ops[0], ast.Lt), "the reference compares with <"
expr = ast.Expression(ast.Lambda(
    args=ast.arguments(posonlyargs=[], args=[ast.arg("c1"), ast.arg("c2"), ast.arg("params")], kwonlyargs=[], kw_defaults=[], defaults=[]),
    body=cmp))
cooperate = eval(compile(ast.fix_missing_locations(expr), REF, "eval"), ns)

rng = np.random.default_rng(11)
n = 4096
c1 = rng.integers(0, 1 << 63, n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, n, dtype=np.uint64)
flips = rng.integers(0, 14, n)                      # most pairs close to the thresholds in use (6 by default)
c2 = c1.copy()
for i in range(n):
    for b in rng.choice(64, int(flips[i]), replace=False):
        c2[i] ^= np.uint64(1) << np.uint64(b)
c2[:64] = rng.integers(0, 1 << 63, 64, dtype=np.uint64)  # and some far apart
dist = np.array([cultural_distance(int(a), int(b)) for a, b in zip(c1, c2)], np.uint32)
thresholds = np.arange(0, 14, dtype=np.uint32)
coop = np.array([[bool(cooperate(int(a), int(b), {"cooperation_threshold": int(t)})) for t in thresholds] for a, b in zip(c1, c2)])
np.savez_compressed(OUT, c1=c1, c2=c2, distance=dist, thresholds=thresholds, cooperate=coop)
print("wrote", OUT, "distances", np.bincount(dist)[:16])
