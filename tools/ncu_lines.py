#!/usr/bin/env python
"""Aggregate an ncu report's source page per CUDA source line: samples, instructions, main stalls.
usage: ncu_lines.py report.ncu-rep [kernel-substring] [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; kern = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks = []; cur = None; path = ""
for r in rows:
    if r and r[0] == "File Path": path = r[1].split("/")[-1]
    if r and r[0] == "Function Name": cur = {"name": r[1], "rows": [], "hdr": None, "file": path}; blocks.append(cur)
    elif r and r[0] == "Line No" and cur is not None: cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and len(r) >= len(cur["hdr"]): cur["rows"].append(r)
# one block per (function, source file): merge the files of a function
merged = {}
for b in blocks:
    if kern not in b["name"] or b["hdr"] is None: continue
    m = merged.setdefault(b["name"], {"name": b["name"], "hdr": b["hdr"], "rows": []})
    for r in b["rows"]:
        r = list(r); r[1] = b["file"] + ": " + r[1]; m["rows"].append(r)
for b in merged.values():
    h = b["hdr"]; isamp = h.index("# Samples"); iinst = h.index("Instructions Executed")
    cols = {k: h.index(k) for k in ("stall_long_sb", "stall_wait", "stall_short_sb", "stall_branch_resolving", "stall_lg", "stall_math", "stall_no_inst", "stall_not_selected",
                                    "stall_barrier", "stall_membar", "stall_mio", "stall_dispatch", "stall_drain", "stall_sleep", "stall_selected", "stall_misc") if k in h}
    lines = []
    for r in b["rows"]:
        if r[0] == "": continue
        try: lines.append((int(r[isamp]), int(r[iinst]), {k: int(r[i]) for k, i in cols.items()}, r[0], r[1][:100]))
        except ValueError: pass
    ts = sum(l[0] for l in lines) or 1; ti = sum(l[1] for l in lines) or 1
    print(f"== {b['name'][:60]}: samples {ts}, warp-instructions {ti}")
    tot = {k: sum(l[2][k] for l in lines) for k in cols}
    print("   stall totals:", {k.replace('stall_', ''): v for k, v in tot.items()})
    for l in sorted(lines, key=lambda x: -x[0])[:top]:
        st = ",".join(f"{k.replace('stall_','')}={v}" for k, v in l[2].items() if v > l[0] * 0.15)
        print(f"  {100*l[0]/ts:5.1f}% smp {100*l[1]/ti:5.1f}% inst  L{l[3]:>4}: {l[4]}   [{st}]")
