#!/usr/bin/env python
"""Top source lines of a kernel by executed warp-instructions (ncu source page).
usage: ncu_inst.py report.ncu-rep kernel-substring [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; kern = sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks = []; cur = None
for r in rows:
    if r and r[0] == "Function Name": cur = {"name": r[1], "rows": [], "hdr": None}; blocks.append(cur)
    elif r and r[0] == "Line No" and cur is not None: cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and len(r) >= len(cur["hdr"]): cur["rows"].append(r)
for b in blocks:
    if kern not in b["name"]: continue
    h = b["hdr"]; iinst = h.index("Instructions Executed"); ithr = h.index("Thread Instructions Executed") if "Thread Instructions Executed" in h else None
    lines = []
    for r in b["rows"]:
        if r[0] == "": continue
        try: lines.append((int(r[iinst]), int(r[ithr]) if ithr is not None else 0, r[0], r[1][:110]))
        except ValueError: pass
    ti = sum(l[0] for l in lines) or 1
    print(f"== {b['name'][:60]}: warp-instructions {ti}")
    for l in sorted(lines, key=lambda x: -x[0])[:top]:
        print(f"  {100*l[0]/ti:5.1f}% inst  (avg active {l[1]/max(1,l[0]):4.1f})  L{l[2]:>4}: {l[3]}")
    break
