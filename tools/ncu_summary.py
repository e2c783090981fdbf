#!/usr/bin/env python
"""Summarise one kernel of an `ncu --set full` report into profiles/<name>.md and .json.
usage: ncu_summary.py report.ncu-rep kernel-substring out-prefix families-per-launch "command line of the capture" """
import csv, json, subprocess, sys
rep, kern, out, fam, cmd = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4]), sys.argv[5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
row = next(r for r in rows[2:] if kern in r[h.index("Kernel Name")])
d = dict(zip(h, row)); u = dict(zip(h, units))
def val(k):
    v = float(d[k]); un = u[k]
    mult = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1}.get(un, 1)
    return v * mult
keys = {
    "duration_s": "gpu__time_duration.sum", "dram_read_bytes": "dram__bytes_read.sum", "dram_write_bytes": "dram__bytes_write.sum",
    "dram_sectors_read": "dram__sectors_read.sum", "warp_instructions": "smsp__inst_executed.sum",
    "registers_per_thread": "launch__registers_per_thread", "achieved_occupancy_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "ipc_active": "sm__inst_executed.avg.per_cycle_active", "issue_slots_busy_pct": "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "l1_hit_pct": "l1tex__t_sector_hit_rate.pct", "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "eligible_warps_per_scheduler": "smsp__warps_eligible.avg.per_cycle_active", "active_warps_per_scheduler": "smsp__warps_active.avg.per_cycle_active",
    "avg_active_threads_per_warp": "smsp__thread_inst_executed_per_inst_executed.ratio",
}
s = {}
for name, k in keys.items():
    if k in d and d[k] not in ("", "n/a"):
        try: s[name] = val(k)
        except ValueError: pass
s["kernel"] = d["Kernel Name"]; s["grid"] = d.get("launch__grid_size"); s["block"] = d.get("launch__block_size")
s["families_per_launch"] = fam
s["dram_bytes_per_launch"] = s["dram_read_bytes"] + s["dram_write_bytes"]
s["dram_bytes_per_family_update"] = s["dram_bytes_per_launch"] / fam
s["warp_instructions_per_family_update"] = s["warp_instructions"] / fam
s["dram_gbs"] = s["dram_bytes_per_launch"] / s["duration_s"] / 1e9
s["dram_sectors_read_per_s"] = s.get("dram_sectors_read", 0) / s["duration_s"]
s["command"] = cmd
json.dump(s, open(out + ".json", "w"), indent=1)
with open(out + ".md", "w") as f:
    f.write(f"# ncu `--set full --clock-control none` summary of `{s['kernel']}`\n\nCapture: `{cmd}`  \n(one launch, profiler-serialised and cold-cache: compare shares and per-unit figures, not absolute times)\n\n| metric | value |\n|---|---|\n")
    for k, v in s.items():
        if k in ("command", "kernel"): continue
        f.write(f"| {k} | {v:.6g} |\n" if isinstance(v, float) else f"| {k} | {v} |\n")
print(json.dumps(s))
