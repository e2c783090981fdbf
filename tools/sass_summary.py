#!/usr/bin/env python
"""Register / spill / shared-memory usage and the SASS opcode histogram of the step kernels of libdsim_b200.so.
usage: sass_summary.py [kernel ...] > profiles/<round>_sass.md   (runs cuobjdump on the built library; no GPU needed)"""
import collections, os, re, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "dispersal-simulation_b200", "libdsim_b200.so")
want = sys.argv[1:] or ["k_move", "k_patch_small", "k_children", "k_patch_mid", "k_patch", "k_lifecycle", "k_segments", "k_scatter"]
res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
usage = {}
name = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        name = m.group(1)
    elif name and "REG:" in line:
        usage[name] = line.strip()
        name = None
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs, cur = {}, None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []
    elif cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
        funcs[cur].append(line)
print("# SASS summary of the step kernels (`cuobjdump --dump-resource-usage` / `-sass` on libdsim_b200.so, sm_100a)\n")
for k in want:
    for f, lines in funcs.items():
        if re.fullmatch(r"_Z\d+%s8StepArgs.*" % k, f):
            ops = collections.Counter()
            for l in lines:
                t = re.sub(r"^\s+/\*[0-9a-f]+\*/\s+", "", l)
                t = re.sub(r"^@!?U?P\w+\s+", "", t)
                ops[t.split()[0].rstrip(";")] += 1
            fam = collections.Counter()
            for o, c in ops.items():
                fam[o.split(".")[0]] += c
            print(f"## `{k}` — {len(lines)} instructions\n")
            print(f"`{usage.get(f, '?')}`\n")
            print("Opcode families: " + ", ".join(f"{o} {c}" for o, c in fam.most_common(24)) + "\n")
            tags = {"TMA bulk copy (UBLKCP)": "UBLKCP", "mbarrier (SYNCS)": "SYNCS", "cp.async (LDGSTS)": "LDGSTS", "128-bit global loads (LDG.E.128)": "LDG.E.128",
                    "64-bit multiply (IMAD.WIDE.U32: one per Philox multiplier per round)": "IMAD.WIDE.U32", "global atomics (ATOMG/RED)": ("ATOMG", "RED"),
                    "local-memory spills (STL/LDL)": ("STL", "LDL")}
            for label, pat in tags.items():
                pats = pat if isinstance(pat, tuple) else (pat,)
                n = sum(c for o, c in ops.items() if any(o.startswith(p) for p in pats))
                print(f"* {label}: {n}")
            print()
