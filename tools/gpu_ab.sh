#!/bin/bash
# A/B of library variants under a tight GPU budget: per variant the render parity tests and one short bench line.
# usage: tools/gpu_ab.sh tag variant ...     (variant = neuralbody_b200/libnb_<variant>.so; "default" = the shipped library)
tag=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  lib=""; [ "$v" != default ] && lib="$PWD/neuralbody_b200/libnb_${v}.so"
  NB_LIB_PATH=$lib timeout 300 python -m pytest tests/test_render_gpu.py -m gpu -q -x --timeout 120 > gpurun_out/${tag}_${v}_pytest.log 2>&1
  echo "$v pytest rc=$? $(tail -1 gpurun_out/${tag}_${v}_pytest.log)"
  NB_LIB_PATH=$lib timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_${v}.json 2> gpurun_out/${tag}_${v}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_${v}.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("%-10s rays/s %.3e  ms/step %.3f  kernel_ms %.3f  frac %.3f  e2e %.3e" % ("$v", d["value"], d["ms_per_step"], r["kernel_ms"], r["frac"], d["e2e"]["value"]))
except Exception as e:
    print("$v", "bench parse failed", e)
PY
done
