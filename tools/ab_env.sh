#!/bin/bash
# usage: tools/ab_env.sh <config> <steps> "<ENV=val ...>" ...   (A/B of environment knobs on the GPU box, same library)
cfg=$1; steps=$2; shift 2
for envs in "$@"; do
  env $envs python bench.py --config $cfg --steps $steps --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/ab.json 2> gpurun_out/ab.log || { echo "$envs FAILED"; tail -3 gpurun_out/ab.log; continue; }
  python - "$envs" <<PY
import json,sys
d=json.load(open("gpurun_out/ab.json"))
print(sys.argv[1], "$cfg", "%.3e"%d["value"], "ms/step %.4f"%d["ms_per_step"], {k:round(v["ms_per_launch"]*1000,1) for k,v in d["kernels"].items()})
PY
done
