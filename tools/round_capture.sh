#!/bin/bash
# One GPU-box call that refreshes what profiles/ is built from: the bench lines (driver regime and default), the ncu launch
# list and the ncu --set full capture of k_move in the regime the driver's bench times (20 steps after 5: step 15).
# Usage (here): gpurun --timeout 900 -- tools/round_capture.sh ; then tools/round_profiles.sh
set -u
mkdir -p gpurun_out
timeout 170 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_20.json 2> gpurun_out/r02_bench_n1_20.log; echo "bench 20 rc=$?"
timeout 170 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.log; echo "bench rc=$?"
tail -c 400 gpurun_out/r02_bench_n1.json
timeout 120 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.log; echo "reference rc=$?"
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --profile-steps 0 --no-graphs > gpurun_out/r02_ncu1.log 2>&1; echo "launch list rc=$?"
timeout 150 ncu --set full --import-source on --clock-control none -k regex:k_move --launch-skip 15 --launch-count 1 -f -o gpurun_out/r02_k_move \
    python bench.py --steps 2 --warmup 15 --no-cpu-baseline --no-e2e --profile-steps 0 --no-graphs > gpurun_out/r02_ncu2.log 2>&1; echo "k_move capture rc=$?"
grep -i "live families" gpurun_out/r02_ncu2.log | tail -2
