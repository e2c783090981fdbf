#!/usr/bin/env python
"""Golden vectors for the grid topology of the gavin2017processbased oracle from the REFERENCE's own methods.

supplement/gavin2017processbased/model.py cannot be imported (cartopy, shapely, a Natural Earth download, a WorldClim
raster at an absolute path), so `Grid.gridcell` and `Grid.neighbors` (model.py:193-262) are taken out of its syntax
tree and attached to a minimal cell class with the attributes they use (`grid[m].shape`, `all_gridcells`, `m`, `ij`,
`popcap`, `language`).  Recorded for several grid shapes:
* neighbors(include_unlivable=True, include_foreign=True): the six neighbours in the reference's order, the cell itself
  where the grid ends;
* neighbors() with the defaults on random states: only neighbours of the same language or without one, and only those
  with a population capacity of at least 1 — what `Language.grow` and the seeding step iterate over.
Also the constants of PopulationCapModel (model.py:265-267) and `area` (model.py:127).
Output: tests/golden/gavin_grid.npz, checked by tests/test_gavin.py against oracle/gavin.py.
    python tests/golden/make_gavin_grid_golden.py
"""
import ast
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/supplement/gavin2017processbased/model.py"
tree = ast.parse(open(REF).read())
grid_cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "Grid")
methods = [n for n in grid_cls.body if isinstance(n, ast.FunctionDef) and n.name in ("gridcell", "neighbors")]
assert len(methods) == 2
popcap_cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "PopulationCapModel")
consts = {n.targets[0].id: eval(compile(ast.Expression(n.value), REF, "eval")) for n in popcap_cls.body if isinstance(n, ast.Assign)}
area = next(eval(compile(ast.Expression(n.value), REF, "eval")) for n in tree.body
            if isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") == "area")


def make_class(R, C):
    body = "class Cell:\n    pass\n"
    ns = {}
    exec(body, ns)
    cls = ns["Cell"]
    mod = ast.Module(body=[ast.ClassDef(name="Ref", bases=[], keywords=[], body=methods, decorator_list=[])], type_ignores=[])
    ns2 = {}
    exec(compile(ast.fix_missing_locations(mod), REF, "exec"), ns2)
    ref = ns2["Ref"]

    def init(self, m, i, j):
        self.m, self.ij, self.population, self.popcap, self.language = m, (i, j), 0, 5.0, None
    ref.__init__ = init
    ref.grid = {0: np.zeros((R, C, 2)), 1: np.zeros((R, C, 2))}
    ref.all_gridcells = {}
    return ref


def ident(cell, R, C):
    return cell.m * R * C + cell.ij[0] * C + cell.ij[1]


out = {"alpha": consts["alpha"], "beta": consts["beta"], "area": float(area)}
shapes = [(1, 1), (1, 5), (4, 1), (3, 4), (5, 5), (2, 7), (6, 9)]
out["shapes"] = np.array(shapes)
rng = np.random.default_rng(77)
for R, C in shapes:
    ref = make_class(R, C)
    n = 2 * R * C
    cells = [ref.gridcell(m, i, j) for m in range(2) for i in range(R) for j in range(C)]
    assert [ident(c, R, C) for c in cells] == list(range(n))
    out[f"all_{R}x{C}"] = np.array([[ident(q, R, C) for q in c.neighbors(True, True)] for c in cells], np.int32)
    # random states: capacities around 1, three languages and unowned cells
    popcap = np.round(rng.uniform(0.0, 3.0, n), 3)
    language = rng.integers(0, 4, n)              # 0 = None
    for c, pc, lg in zip(cells, popcap, language):
        c.popcap, c.language = float(pc), (None if lg == 0 else int(lg))
    lists = [[ident(q, R, C) for q in c.neighbors()] for c in cells]
    flat = np.array([q for l in lists for q in l], np.int32)
    out[f"popcap_{R}x{C}"] = popcap
    out[f"language_{R}x{C}"] = language.astype(np.int32)
    out[f"default_off_{R}x{C}"] = np.cumsum([0] + [len(l) for l in lists]).astype(np.int32)
    out[f"default_{R}x{C}"] = flat
np.savez_compressed(os.path.join(HERE, "gavin_grid.npz"), **out)
print("alpha", out["alpha"], "beta", out["beta"], "area", out["area"], "shapes", shapes)
