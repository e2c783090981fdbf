#!/usr/bin/env python
"""Per CUDA source line: global-memory sectors requested from L2 ("L2 Theoretical Sectors Global" of the ncu source page).
usage: ncu_sectors.py report.ncu-rep kernel-substring [units-per-launch] [top]"""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
units = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; path = ""; agg = collections.defaultdict(lambda: [0, 0, "", ""])
for r in rows:
    if r and r[0] == "File Path": path = r[1].split("/")[-1]
    elif r and r[0] == "Function Name": cur = {"name": r[1], "hdr": None, "file": path}
    elif r and r[0] == "Line No" and cur is not None: cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and len(r) >= len(cur["hdr"]) and kern in cur["name"] and r[0] != "":
        h = cur["hdr"]
        try:
            sec = int(r[h.index("L2 Theoretical Sectors Global")]); tag = int(r[h.index("L1 Tag Requests Global")])
        except ValueError:
            continue
        if sec or tag:
            a = agg[(cur["file"], r[0])]
            a[0] += sec; a[1] += tag; a[2] = r[1][:90]; a[3] = r[h.index("Access Operation")] if "Access Operation" in h else ""
tot = sum(a[0] for a in agg.values())
print(f"global-memory sectors requested from L2 by {kern}: {tot} = {tot / units:.1f} per unit")
for (f, ln), a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"  {100 * a[0] / max(tot, 1):5.1f}%  {a[0] / units:7.2f}/unit  {f}:{ln}: {a[2]}")
