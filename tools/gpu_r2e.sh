#!/bin/bash
# last GPU call of the round (about a minute of budget): render parity tests on the shipped library and one bench line
tag=${1:-r2e}
mkdir -p gpurun_out
timeout 40 python -m pytest tests/test_render_gpu.py -m gpu -q -x --timeout 30 > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/${tag}_pytest.log)"
timeout 25 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_default.json 2> gpurun_out/${tag}_default.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_default.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("rays/s %.3e  ms/step %.3f  kernel_ms %.3f  frac %.3f  e2e %.3e" % (d["value"], d["ms_per_step"], r["kernel_ms"], r["frac"], d["e2e"]["value"]))
except Exception as e:
    print("bench parse failed", e)
PY
