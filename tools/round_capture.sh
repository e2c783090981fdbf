#!/bin/bash
# One GPU-box call that refreshes what profiles/ is built from: the default bench line, the ncu launch list and the
# steady-state ncu capture of k_move.  Usage (here): gpurun --timeout 900 -- tools/round_capture.sh
set -u
mkdir -p gpurun_out
timeout 170 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.log; echo "bench rc=$?"
tail -c 600 gpurun_out/r02_bench_n1.json
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --profile-steps 0 --no-graphs > gpurun_out/r02_ncu1.log 2>&1; echo "launch list rc=$?"
timeout 150 ncu --set full --import-source on --clock-control none -k regex:k_move --launch-skip 104 --launch-count 1 -f -o gpurun_out/r02_k_move \
    python bench.py --steps 2 --warmup 104 --no-cpu-baseline --no-e2e --profile-steps 0 --no-graphs > gpurun_out/r02_ncu2.log 2>&1; echo "k_move capture rc=$?"
grep -i "live families" gpurun_out/r02_ncu2.log | tail -2
