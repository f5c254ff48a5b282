#!/bin/bash
# One gpurun call worth of round evidence: -m gpu suites, benches (train / infer / ese / reference arm), the per-kernel
# roofline table, ncu launch lists and full captures of the dominant kernels (B200_PROFILING.md recipe).
# Everything lands in gpurun_out/.   usage: tests/run_gpu_round.sh [tag] [noncu]
cd "$(dirname "$0")/.."
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/gpu.txt 2>&1
P="python -m pytest -q -m gpu -p no:cacheprovider"
rm -f $O/parity_report.jsonl
timeout 1200 $P tests/test_gpu_kernels.py > $O/test_kernels_$TAG.log 2>&1; echo "kernels: $?"; tail -n 2 $O/test_kernels_$TAG.log
timeout 1200 $P tests/test_gpu_model.py > $O/test_model_$TAG.log 2>&1; echo "model: $?"; tail -n 2 $O/test_model_$TAG.log
timeout 1200 $P tests/test_gpu_fullsize.py tests/test_gpu_cli.py > $O/test_fullsize_cli_$TAG.log 2>&1; echo "fullsize+cli: $?"; tail -n 2 $O/test_fullsize_cli_$TAG.log
timeout 600 python __graft_entry__.py smoke > $O/smoke_$TAG.log 2>&1; echo "smoke: $?"; tail -n 1 $O/smoke_$TAG.log

timeout 900 python bench.py > $O/bench_train_$TAG.json 2> $O/bench_train_$TAG.err; echo "bench train: $?"
timeout 600 python bench.py --workload infer > $O/bench_infer_$TAG.json 2> $O/bench_infer_$TAG.err; echo "bench infer: $?"
timeout 600 python bench.py --workload ese --steps 3 > $O/bench_ese_$TAG.json 2> $O/bench_ese_$TAG.err; echo "bench ese: $?"
timeout 600 python bench.py --workload infer --precision split --no-cpu-baseline > $O/bench_infer_split_$TAG.json 2> $O/bench_infer_split_$TAG.err; echo "bench infer split: $?"
timeout 600 python bench.py --workload bands --no-cpu-baseline > $O/bench_bands_$TAG.json 2> $O/bench_bands_$TAG.err; echo "bench bands: $?"
timeout 600 python bench.py --variant upr --steps 4 --no-cpu-baseline > $O/bench_train_upr_$TAG.json 2> $O/bench_train_upr_$TAG.err; echo "bench upr: $?"
timeout 600 python bench.py --variant dpp --steps 4 --no-cpu-baseline > $O/bench_train_dpp_$TAG.json 2> $O/bench_train_dpp_$TAG.err; echo "bench dpp: $?"
timeout 600 python bench.py --bs 64 --no-cpu-baseline > $O/bench_train_bs64_$TAG.json 2> $O/bench_train_bs64_$TAG.err; echo "bench bs64: $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_$TAG.json 2> $O/bench_reference_$TAG.err; echo "bench reference: $?"
timeout 900 python tools/kernel_bench.py --out $O/kernel_roofline_$TAG.jsonl > $O/kernel_bench_$TAG.log 2>&1; echo "kernel bench: $?"

if [ "$2" != "noncu" ]; then
  NCU="ncu --clock-control none"
  CMD="python bench.py --bs 64 --steps 2 --warmup 3 --no-cpu-baseline"
  $CMD > $O/plain_train.log 2>&1 &&
  $NCU --metrics gpu__time_duration.sum -s 1400 -c 1000 --csv --log-file $O/launches_train_bs64_$TAG.csv $CMD > $O/ncu_launch_train.log 2>&1
  CMD="python bench.py --workload infer --steps 2 --warmup 3 --no-cpu-baseline"
  $CMD > $O/plain_infer.log 2>&1 &&
  $NCU --metrics gpu__time_duration.sum -c 300 --csv --log-file $O/launches_infer_$TAG.csv $CMD > $O/ncu_launch_infer.log 2>&1
  # full captures, one kernel each, from the per-kernel bench (3 warm-up launches skipped)
  cap() {   # name, kernel regex, --only filter
    python tools/kernel_bench.py --only "$3" --reps 1 > $O/plain_$1.log 2>&1 &&
    $NCU --set full --import-source on -k "regex:$2" -s 3 -c 1 -f -o $O/prof_$1_$TAG python tools/kernel_bench.py --only "$3" --reps 1 > $O/ncu_$1.log 2>&1
    tail -n 1 $O/ncu_$1.log
  }
  cap conv280 conv2x2_tc2 "conv2x2 280->280 pad0 train"
  cap conv70 conv2x2_tc2 "conv2x2 70->70 pad0 train"
  cap wgrad280 conv2x2_wgrad2 "wgrad 280->280 pad0 train"
  cap bn_apply slot_map "bn_apply_relu"
  cap bn_bwd_reduce col_reduce "bn_bwd_reduce"
  cap bn_bwd_apply slot_map "bn_bwd_apply"
  cap lf_shift lf_shift "lf_shift_kernel full"
  cap shift_pack pack_views "shift_pack_kernel full"
  cap loss_ce loss_ce "loss_ce_kernel 64x108x96x96 on-the-fly"
  cap dpp_head dpp_head "dpp_head"
  cap augment augment_views "augment"
fi
ls -la $O | tail -n 40
