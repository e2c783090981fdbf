#!/usr/bin/env python
"""Per-kernel table of an ncu launch list (`--metrics gpu__time_duration.sum --csv`).
usage: launch_table.py launches.csv "command line of the capture" > profiles/<name>.md"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rows:
    name = r[4].split("(")[0]
    tot[name] += float(r[-1]) / 1e3
    cnt[name] += 1
total = sum(tot.values())
print(f"# ncu launch list of `{sys.argv[2]}`\n")
print("`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` (cold-cache, serialised: compare SHARES). "
      f"Raw list: `{sys.argv[1].split('/')[-1]}`.\n")
print("| kernel | launches | total (us) | mean (us) | share |\n|---|---|---|---|---|")
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"| {k} | {cnt[k]} | {tot[k]:.1f} | {tot[k] / cnt[k]:.1f} | {100 * tot[k] / total:.1f}% |")
step = [k for k in tot if cnt[k] == max(cnt.values())]
s = sum(tot[k] for k in step)
print(f"\nKernels of the step loop only ({', '.join(sorted(step))}): {s:.1f} us over {max(cnt.values())} steps; "
      f"k_move = {100 * tot.get('k_move', 0) / s:.1f}% of the step.")
