#!/bin/bash
# ncu launch list + full capture of the conv kernel on the full-LF inference workload (B200_PROFILING.md recipe)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --workload infer --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_infer.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_infer.csv $CMD > gpurun_out/ncu_launch.log 2>&1
$CMD > gpurun_out/plain_infer2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv2x2_tc -s 143 -c 2 -f -o gpurun_out/prof_conv $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_launch.log gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep
