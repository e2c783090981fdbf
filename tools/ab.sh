#!/bin/bash
# usage: tools/ab.sh <config> <steps> lib1.so lib2.so ...   (A/B of library builds on the GPU box; "default" = the product build)
cfg=$1; steps=$2; shift 2
for lib in "$@"; do
  if [ "$lib" = default ]; then unset DSIM_B200_LIB; else export DSIM_B200_LIB=$PWD/$lib; fi
  python bench.py --config $cfg --steps $steps --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/ab.json 2> gpurun_out/ab.log || { echo "$lib FAILED"; tail -3 gpurun_out/ab.log; continue; }
  python - "$lib" <<PY
import json,sys
d=json.load(open("gpurun_out/ab.json"))
print(sys.argv[1], cfg:="$cfg", "%.3e"%d["value"], "ms/step %.4f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], {k:round(v["ms_per_launch"]*1000,1) for k,v in d["kernels"].items()})
PY
done
