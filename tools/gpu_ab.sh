#!/bin/bash
# A/B of library variants built with neuralbody_b200._build.build(defines=..., out=...): one short bench line each.
# usage: tools/gpu_ab.sh tag variant1 variant2 ...   (variant "default" = the shipped library; "prec:tc_fp16" = 1-pass mode)
tag=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  extra=""
  lib=""
  case "$v" in
    default) ;;
    prec:*) extra="--precision ${v#prec:}" ;;
    *) lib="$PWD/neuralbody_b200/libnb_${v}.so" ;;
  esac
  NB_LIB_PATH=$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $extra > gpurun_out/${tag}_${v//:/_}.json 2> gpurun_out/${tag}_${v//:/_}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_${v//:/_}.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("%-14s rays/s %.3e  ms/step %.3f  kernel_ms %.3f  frac %.3f  parity?" % ("$v", d["value"], d["ms_per_step"], r["kernel_ms"], r["frac"]))
except Exception as e:
    print("$v", "bench parse failed", e)
PY
done
