#!/bin/bash
# One gpurun call worth of round evidence: -m gpu suites, benches (train / infer / reference arm), ncu launch lists
# and one full capture of the conv kernel (B200_PROFILING.md recipe).  Everything lands in gpurun_out/.
#   usage: tests/run_gpu_round.sh [tag]      (tag is appended to the output file names)
cd "$(dirname "$0")/.."
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/gpu.txt 2>&1
P="python -m pytest -q -m gpu -p no:cacheprovider"
rm -f $O/parity_report.jsonl
timeout 1200 $P tests/test_gpu_kernels.py > $O/test_kernels_$TAG.log 2>&1; echo "kernels: $?"; tail -n 2 $O/test_kernels_$TAG.log
timeout 1200 $P tests/test_gpu_model.py > $O/test_model_$TAG.log 2>&1; echo "model: $?"; tail -n 2 $O/test_model_$TAG.log
timeout 600 python __graft_entry__.py smoke > $O/smoke_$TAG.log 2>&1; echo "smoke: $?"; tail -n 1 $O/smoke_$TAG.log

timeout 900 python bench.py > $O/bench_train_$TAG.json 2> $O/bench_train_$TAG.err; echo "bench train: $?"
timeout 600 python bench.py --workload infer > $O/bench_infer_$TAG.json 2> $O/bench_infer_$TAG.err; echo "bench infer: $?"
timeout 600 python bench.py --bs 64 --no-cpu-baseline > $O/bench_train_bs64_$TAG.json 2> $O/bench_train_bs64_$TAG.err; echo "bench bs64: $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_$TAG.json 2> $O/bench_reference_$TAG.err; echo "bench reference: $?"

if [ "$2" != "noncu" ]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
  $CMD > $O/plain_train.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 1000 --csv --log-file $O/launches_train_$TAG.csv $CMD > $O/ncu_launch_train.log 2>&1
  CMD="python bench.py --workload infer --steps 2 --warmup 3 --no-cpu-baseline"
  $CMD > $O/plain_infer.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_infer_$TAG.csv $CMD > $O/ncu_launch_infer.log 2>&1
  $CMD > $O/plain_infer2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:conv2x2_tc -s 143 -c 2 -f -o $O/prof_conv_$TAG $CMD > $O/ncu_full.log 2>&1
  tail -n 2 $O/ncu_launch_train.log $O/ncu_launch_infer.log $O/ncu_full.log
fi
ls -la $O | tail -n 30
