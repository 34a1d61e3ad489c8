#!/bin/bash
# One GPU iteration: parity tests, a short bench line and the in-kernel timeline; everything lands in gpurun_out/.
# usage: tools/gpu_iter.sh [tag] [pytest -k expression]
tag=${1:-iter}
kexpr=${2:-}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
if [ -n "$kexpr" ]; then
  timeout 900 python -m pytest tests -m gpu -q --timeout 300 -k "$kexpr" > gpurun_out/${tag}_pytest.log 2>&1
else
  timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/${tag}_pytest.log 2>&1
fi
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -15 gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("rays/s %.3e  ms/step %.3f  kernel_ms %.3f  frac %.3f  e2e %.3e" % (d["value"], d["ms_per_step"], r["kernel_ms"], r["frac"], d["e2e"]["value"]))
except Exception as e:
    print("bench parse failed", e)
PY
timeout 300 python tools/trace_timeline.py tc_fp16x3 40,100,200,280 1 > gpurun_out/${tag}_trace.txt 2> gpurun_out/${tag}_trace.err
echo "trace rc=$?"
tail -12 gpurun_out/${tag}_trace.txt
if [ -n "$NB_NCU" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_tc_list -s 2 -c 1 -f -o gpurun_out/${tag}_prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1
  echo "ncu rc=$?"; ls -la gpurun_out/${tag}_prof.ncu-rep
fi
