#!/bin/bash
# Multi-GPU checks (gpurun --gpus N): the default scaling line and the c4 / c5 configurations at N ranks.
# usage: tools/gpu_multi.sh tag N [quick]
tag=${1:-multi}; N=${2:-2}; quick=$3
mkdir -p gpurun_out
run() {  # name, args...
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N "$@" > gpurun_out/${tag}_${name}_n${N}.json 2> gpurun_out/${tag}_${name}_n${N}.err
  echo "$name rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_${name}_n${N}.json").read().strip().splitlines()[-1])
    print("  %-4s N=%d value %.3e %s  ms/step %.3f  e2e %.3e  bit_identical %s  share %s" % ("$name", d["n_gpus"], d["value"], d["unit"], d["ms_per_step"],
          d["e2e"]["value"], d.get("multi_gpu_bit_identical"), d["roofline"].get("kernel_share_of_step")))
except Exception as e:
    print("  $name parse failed", e)
PY
  tail -2 gpurun_out/${tag}_${name}_n${N}.err
}
if [ -n "$quick" ]; then
  run c2 --steps 10 --warmup 3
  run c4 --config c4 --views 24 --steps 2
  run c5 --config c5 --poses $N --c5-size 512 --steps 2
else
  run c2 --steps 20 --warmup 5
  run c4 --config c4 --steps 2
  run c5 --config c5 --steps 3
fi
