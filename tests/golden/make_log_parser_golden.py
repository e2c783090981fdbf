#!/usr/bin/env python
"""The reference's own log-parsing expressions, as a fixture for the stdout protocol test.

supplement/plotting/plot.py:290-381 is the consumer of the model's stdout.  The module cannot be imported (matplotlib,
cartopy, an OSM extract, a command line), so the expressions that touch the `t:`, `POPULATION:` and `F:` lines are taken
out of its syntax tree: the `bitvec` pattern and its replacement, the comprehension that builds `content`, the regular
expression of the `F:` line, and the expression that reads the time stamp.  tests/test_driver.py evaluates exactly these
on the log our driver writes.

usage (in the build container, where /root/reference exists): python tests/golden/make_log_parser_golden.py
"""
import ast
import json
import os
import warnings

REF = "/root/reference/supplement/plotting/plot.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "log_parser.json")

with warnings.catch_warnings():
    warnings.simplefilter("ignore", SyntaxWarning)
    tree = ast.parse(open(REF).read())

out = {"source": "supplement/plotting/plot.py"}
for node in ast.walk(tree):
    # bitvec = re.compile('"\[([01]*)\]"')
    if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", "") == "bitvec":
        out["bitvec_pattern"] = node.value.args[0].value
        out["bitvec_line"] = node.lineno
    # line = bitvec.sub("0b\\1", line)
    if (isinstance(node, ast.Assign) and getattr(node.targets[0], "id", "") == "line" and isinstance(node.value, ast.Call)
            and isinstance(node.value.func, ast.Attribute) and getattr(node.value.func.value, "id", "") == "bitvec"):
        out["bitvec_repl"] = node.value.args[0].value
    # content = [(x, y, n, int(c, 16)) for x, y, p in literal_eval(line[len("POPULATION: "):]) for c, n in p.items()]
    if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", "") == "content" and isinstance(node.value, ast.ListComp):
        out["content_expr"] = ast.unparse(node.value)
        out["content_line"] = node.lineno
    # match = re.search(r'escendence: "([^"]*)".*ocation: NodeIndex\(([0-9]*)\),.*culture: ([01]*)', line)
    if (isinstance(node, ast.Assign) and getattr(node.targets[0], "id", "") == "match" and isinstance(node.value, ast.Call)
            and getattr(node.value.func, "attr", "") == "search"):
        out["f_regex"] = node.value.args[0].value
        out["f_line"] = node.lineno
    # timestamp = int(line[3:]) * season_length_in_years
    if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", "") == "timestamp" and isinstance(node.value, ast.BinOp):
        out["timestamp_expr"] = ast.unparse(node.value)
# the line prefixes the parser dispatches on
out["prefixes"] = sorted({n.args[0].value for n in ast.walk(tree)
                          if isinstance(n, ast.Call) and getattr(n.func, "attr", "") == "startswith"
                          and getattr(n.func.value, "id", "") == "line" and isinstance(n.args[0], ast.Constant)})
assert {"bitvec_pattern", "bitvec_repl", "content_expr", "f_regex", "timestamp_expr"} <= set(out), sorted(out)
json.dump(out, open(OUT, "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1, sort_keys=True))
