#!/bin/bash
# c3 with more steps (per-step distribution) + the launch list of one training step
mkdir -p gpurun_out
timeout 300 python bench.py --config c3 --steps 40 --warmup 5 > gpurun_out/tr2_c3.json 2> gpurun_out/tr2_c3.err; echo "c3 rc=$?"; tail -3 gpurun_out/tr2_c3.err
python -c "
import json
d=json.loads(open('gpurun_out/tr2_c3.json').read().strip().splitlines()[-1])
print('ms/step mean %.3f median %.3f  e2e %.3f  mallocs %s' % (d['ms_per_step'], d['median_step_ms'], d['e2e']['ms_per_step'], d['cuda_mallocs_in_timed_steps']))
print([round(x,2) for x in d['step_ms']])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/tr2_launches_c3.csv python bench.py --config c3 --steps 1 --warmup 3 > gpurun_out/tr2_launches_c3.log 2>&1; echo "launch list rc=$?"
