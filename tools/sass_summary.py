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

# ---- excerpts of k_move: the instructions that show what the kernel is built from ------------------------------------
def clean(l):
    return re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l).rstrip()
km = next((lines for f, lines in funcs.items() if re.fullmatch(r"_Z\d+k_move8StepArgs.*", f)), None)
if km:
    def at(pred, before=0, after=1, which=0):
        ks = [n for n, l in enumerate(km) if pred(l)]
        if not ks:
            return ["(not found in this build)"]
        k = ks[min(which, len(ks) - 1)]
        return [clean(l) for l in km[max(0, k - before):k + after]]
    best = max(range(max(1, len(km) - 14)), key=lambda k: sum(("-0x2daee0ad" in l or "-0x326172a9" in l) for l in km[k:k + 14]))
    print("## `k_move` excerpts (`cuobjdump -sass`)\n")
    print("Prologue: the packed family record (128-bit loads):\n\n```\n" + "\n".join(at(lambda l: "LDG.E.128" in l, 0, 1) + at(lambda l: "LDG.E.128" in l, 0, 1, 1)) + "\n```\n")
    print("TMA bulk copy of a reach-row column into shared memory (one elected lane per column), completion on the warp's mbarrier:\n\n```\n"
          + "\n".join(at(lambda l: "UBLKCP" in l, 6, 2)) + "\n```\n")
    print("Waiting for it:\n\n```\n" + "\n".join(at(lambda l: "SYNCS" in l and "TRYWAIT" in l)) + "\n```\n")
    print("Philox rounds (per round two 32x32->64 multiplies `IMAD.WIDE.U32` by the Philox multipliers 0xD2511F53 / 0xCD9E8D57 and two "
          "3-input XORs `LOP3.LUT ... 0x96` with the round key: 40 instructions per block of four words):\n\n```\n"
          + "\n".join(clean(l) for l in km[best:best + 14]) + "\n```\n")
    print("A 16-byte candidate view gathered straight into shared memory (cp.async = `LDGSTS`, with the L2 eviction-priority descriptor):\n\n```\n"
          + "\n".join(at(lambda l: "LDGSTS.E.128" in l)) + "\n```")
