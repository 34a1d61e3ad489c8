#!/bin/bash
# The non-default bench configurations on one GPU (BASELINE configs[2..4]) + the default line with its CPU baseline.
# usage: tools/gpu_configs.sh tag
tag=${1:-r02}
mkdir -p gpurun_out
run() {  # name, args...
  name=$1; shift
  timeout 900 python bench.py "$@" > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  echo "$name rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_${name}.json").read().strip().splitlines()[-1])
    print("  %-4s value %.4g %s  ms/step %.3f  e2e %.4g  clocks %s" % ("$name", d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d.get("clocks")))
except Exception as e:
    print("  $name parse failed", e)
PY
  tail -2 gpurun_out/${tag}_${name}.err
}
run c2full --steps 20 --warmup 5
run c3 --config c3 --steps 10 --warmup 3
run c4 --config c4 --steps 2 --warmup 3
run c5 --config c5 --steps 3 --warmup 3
run ref --impl reference --steps 2 --warmup 1
