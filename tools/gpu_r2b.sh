#!/bin/bash
# Round-2b GPU iteration (tight GPU budget): parity tests on the shipped library, then one short bench line per library
# variant (NB_LIB_PATH), then the in-kernel timeline of the shipped library.  Everything lands in gpurun_out/.
# usage: tools/gpu_r2b.sh tag [variant ...]     (variant = neuralbody_b200/libnb_<variant>.so; "default" = the shipped one)
tag=$1; shift
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 200 > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
for v in default "$@"; do
  lib=""; [ "$v" != default ] && lib="$PWD/neuralbody_b200/libnb_${v}.so"
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
timeout 200 python tools/trace_timeline.py tc_fp16x3 40,100 1 > gpurun_out/${tag}_trace.txt 2> gpurun_out/${tag}_trace.err
echo "trace rc=$?"; tail -6 gpurun_out/${tag}_trace.txt
