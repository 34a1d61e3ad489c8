#!/bin/bash
# Evidence runs (one GPU): launch list of the default bench step, ncu sections for every kernel, compute-sanitizer.
# usage: tools/gpu_evidence.sh tag
tag=${1:-r02}
mkdir -p gpurun_out
# 1. every launch of two bench steps with its device time (shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_launches_c2.log 2>&1
echo "launch list rc=$?"
# 2. throughput sections of every kernel of the library (second instance of each = warm)
timeout 900 ncu --clock-control none --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy \
    --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed \
    --csv --log-file gpurun_out/${tag}_ncu_all_kernels.csv python tools/profile_all.py > gpurun_out/${tag}_ncu_all_kernels.log 2>&1
echo "all-kernel sections rc=$?"
# 3. compute-sanitizer on the smoke render (tiny: 576 rays x 64 samples through both precisions)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/${tag}_sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -3 gpurun_out/${tag}_sanitizer_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_sanitizer_synccheck.log 2>&1
echo "synccheck rc=$?"; tail -3 gpurun_out/${tag}_sanitizer_synccheck.log
