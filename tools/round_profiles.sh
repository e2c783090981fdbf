#!/bin/bash
# Turns the files tools/round_capture.sh brought back in gpurun_out/ into the tracked summaries under profiles/.
set -eu
cd "$(dirname "$0")/.."
live=$(grep -o "[0-9]* live families on [0-9]* patches" gpurun_out/r02_ncu2.log | tail -1)
fam=$(echo "$live" | awk '{print $1}'); occ=$(echo "$live" | awk '{print $5}')
cmd="ncu --set full --import-source on --clock-control none -k regex:k_move --launch-skip 15 --launch-count 1 python bench.py --steps 2 --warmup 15 --no-cpu-baseline --no-e2e --profile-steps 0 --no-graphs (step 15 = the middle of the driver's 20-step window after 5 warm-up steps: $fam families on $occ patches)"
python tools/ncu_summary.py gpurun_out/r02_k_move.ncu-rep k_move profiles/r02_k_move_ncu "$fam" "$cmd" > /dev/null
python - "$fam" "$occ" <<'PY'
import json, sys
p = "profiles/r02_k_move_ncu.json"
d = json.load(open(p)); d["rho"] = round(float(sys.argv[1]) / float(sys.argv[2]), 3); d["occupied_patches"] = int(sys.argv[2])
json.dump(d, open(p, "w"), indent=1)
PY
python tools/ncu_lines.py gpurun_out/r02_k_move.ncu-rep k_move 45 | cut -c1-200 > profiles/r02_k_move_lines.txt
python tools/ncu_sectors.py gpurun_out/r02_k_move.ncu-rep k_move "$fam" 24 > profiles/r02_k_move_sectors.txt
cp gpurun_out/r02_launches.csv profiles/r02_launches.csv
python tools/launch_table.py profiles/r02_launches.csv "python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --profile-steps 0 --no-graphs" > profiles/r02_launches.md
cp gpurun_out/r02_bench_n1.json gpurun_out/r02_bench_n1_20.json gpurun_out/r02_bench_reference.json profiles/
python tools/sass_summary.py k_move k_patch_small k_patch_one k_children k_patch_mid k_patch k_patch_huge k_lifecycle k_segments k_scatter > profiles/r02_sass.md
echo "profiles refreshed: $fam families on $occ patches"
