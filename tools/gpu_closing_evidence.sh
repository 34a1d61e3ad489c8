#!/bin/bash
# Closing evidence on one GPU (about two minutes of box time): full GPU test suite, the default bench line, c5, optional
# library variants (neuralbody_b200/libnb_<variant>.so), the ncu launch list, one ncu --set full capture of the decoder
# (raw page exported on the box), the in-kernel timeline.      usage: tools/gpu_closing_evidence.sh tag [variant ...]
tag=${1:-closing}; shift
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --timeout 200 > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/${tag}_pytest.log)"
timeout 300 python bench.py > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench_c2.err
echo "bench c2 rc=$?"; tail -c 600 gpurun_out/${tag}_bench_c2.json | head -c 300; echo
for v in "$@"; do
  NB_LIB_PATH=$PWD/neuralbody_b200/libnb_${v}.so timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_${v}.json 2> gpurun_out/${tag}_${v}.err
done
timeout 300 python bench.py --config c5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_c5.json 2> gpurun_out/${tag}_bench_c5.err
echo "bench c5 rc=$?"
python - <<PY
import json
for v in ["bench_c2", "bench_c5"] + "$*".split():
    try:
        d = json.loads(open("gpurun_out/${tag}_%s.json" % v).read().strip().splitlines()[-1])
        r = d["roofline"]
        print("%-10s value %.3e  ms/step %.3f  kernel_ms %.3f  frac %.3f  e2e %.3e" % (v, d["value"], d["ms_per_step"], r["kernel_ms"], r["frac"], d["e2e"]["value"]))
    except Exception as e:
        print(v, "bench parse failed", e)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_launches_c2.log 2>&1
echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:render_tc_list -s 2 -c 1 -f -o gpurun_out/${tag}_decoder \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu full rc=$?"
ncu -i gpurun_out/${tag}_decoder.ncu-rep --page raw --csv > gpurun_out/${tag}_decoder_raw.csv 2> /dev/null
ls -la gpurun_out/${tag}_decoder.ncu-rep
sz=$(stat -c %s gpurun_out/${tag}_decoder.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 45000000 ]; then rm -f gpurun_out/${tag}_decoder.ncu-rep; echo "ncu-rep too large for the return channel, raw page kept"; fi
timeout 200 python tools/trace_timeline.py tc_fp16x3 40,100 1 > gpurun_out/${tag}_trace.txt 2> gpurun_out/${tag}_trace.err
echo "trace rc=$?"; tail -4 gpurun_out/${tag}_trace.txt
