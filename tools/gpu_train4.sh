#!/bin/bash
# full GPU test suite + c3 with operand-layout variants of the training GEMM + launch list of a step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/tr4_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/tr4_pytest.log
for v in default al128 sbo128; do
  lib=$PWD/neuralbody_b200/libnb_$v.so; [ $v = default ] && lib=$PWD/neuralbody_b200/libneuralbody_b200.so
  NB_LIB_PATH=$lib timeout 300 python bench.py --config c3 --steps 20 --warmup 5 > gpurun_out/tr4_c3_$v.json 2> gpurun_out/tr4_c3_$v.err
  python -c "
import json
d=json.loads(open('gpurun_out/tr4_c3_$v.json').read().strip().splitlines()[-1])
print('$v: ms/step mean %.3f median %.3f e2e %.3f' % (d['ms_per_step'], d['median_step_ms'], d['e2e']['ms_per_step']))"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/tr4_launches_c3.csv python bench.py --config c3 --steps 1 --warmup 3 > gpurun_out/tr4_launches_c3.log 2>&1; echo "launch list rc=$?"
