#!/bin/bash
# c3 per-step distribution + one full ncu capture of a fine-pass GEMM (forward fc_0: the 5th gemm launch of a step)
mkdir -p gpurun_out
timeout 300 python bench.py --config c3 --steps 40 --warmup 5 > gpurun_out/tr3_c3.json 2> gpurun_out/tr3_c3.err; echo "c3 rc=$?"; tail -3 gpurun_out/tr3_c3.err
python -c "
import json
d=json.loads(open('gpurun_out/tr3_c3.json').read().strip().splitlines()[-1])
print('ms/step mean %.3f median %.3f  e2e %.3f  mallocs %s clocks %s' % (d['ms_per_step'], d['median_step_ms'], d['e2e']['ms_per_step'], d['cuda_mallocs_in_timed_steps'], d['clocks']))
print([round(x,2) for x in d['step_ms']])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s ${NB_GEMM_SKIP:-52} -c 1 -f -o gpurun_out/tr3_gemm python bench.py --config c3 --steps 1 --warmup 3 > gpurun_out/tr3_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/tr3_ncu.log
