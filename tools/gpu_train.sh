#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_train_gemm_gpu.py -m gpu -q -s --timeout 120 > gpurun_out/tr_gemm.log 2>&1; echo "gemm rc=$?"; tail -25 gpurun_out/tr_gemm.log
if ! grep -q " passed" gpurun_out/tr_gemm.log || grep -q failed gpurun_out/tr_gemm.log; then
  NB_LIB_PATH=$PWD/neuralbody_b200/libnb_sbo128.so timeout 300 python -m pytest tests/test_train_gemm_gpu.py -m gpu -q -s --timeout 120 > gpurun_out/tr_gemm_sbo128.log 2>&1; echo "gemm sbo128 rc=$?"; tail -25 gpurun_out/tr_gemm_sbo128.log
fi
timeout 600 python -m pytest tests/test_backward.py -m gpu -q -s --timeout 300 > gpurun_out/tr_bwd.log 2>&1; echo "bwd rc=$?"; tail -30 gpurun_out/tr_bwd.log
timeout 300 python bench.py --config c3 --steps 10 --warmup 3 > gpurun_out/tr_c3.json 2> gpurun_out/tr_c3.err; echo "c3 rc=$?"; tail -3 gpurun_out/tr_c3.err; cut -c1-1500 gpurun_out/tr_c3.json
