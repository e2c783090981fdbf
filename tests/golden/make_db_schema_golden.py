#!/usr/bin/env python
"""Tables and columns of the reference's network database, as a fixture for the ingestion test.

supplement/distances/database.py:18-76 declares the schema with sqlalchemy (not installed here), so the table and
column names and types are read from its syntax tree.

usage (in the build container, where /root/reference exists): python tests/golden/make_db_schema_golden.py
"""
import ast
import json
import os

REF = "/root/reference/supplement/distances/database.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "db_schema.json")

tree = ast.parse(open(REF).read())
tables = {}
for node in ast.walk(tree):
    if isinstance(node, ast.Call) and getattr(node.func, "attr", "") == "Table" and isinstance(node.args[0], ast.Constant):
        cols = []
        for a in node.args[2:]:
            if isinstance(a, ast.Call) and getattr(a.func, "attr", "") == "Column":
                typ = a.args[1]
                typ = typ.func if isinstance(typ, ast.Call) else typ
                cols.append({"name": a.args[0].value, "type": typ.attr,
                             "primary_key": any(k.arg == "primary_key" and getattr(k.value, "value", False) for k in a.keywords)})
        pk = [[c.value for c in a.args] for a in node.args[2:]
              if isinstance(a, ast.Call) and getattr(a.func, "id", "") == "PrimaryKeyConstraint"]
        tables[node.args[0].value] = {"columns": cols, "primary_key_constraint": pk[0] if pk else None, "line": node.lineno}
json.dump({"source": "supplement/distances/database.py", "tables": tables}, open(OUT, "w"), indent=1, sort_keys=True)
print(json.dumps(tables, indent=1)[:1500])
