#!/bin/bash
# compute-sanitizer memcheck + racecheck over the small-case kernel parity tests.  The caching
# allocator is disabled so that every tensor is its own cudaMalloc and out-of-bounds accesses cannot hide inside a pool.
# usage: tests/run_sanitizer.sh [tag] [memcheck timeout s] [racecheck timeout s];  logs land in gpurun_out/
cd "$(dirname "$0")/.."
TAG=${1:-r02}
TM=${2:-700}
TR=${3:-500}
O=gpurun_out
mkdir -p $O
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
SEL_MEM='test_conv_tc_vs_oracle or test_conv_tc_epilogues or test_conv_wgrad or test_conv_dgrad_and_gate or test_conv_fused_epilogue_outputs or test_conv_split_precision or test_conv_fused_bn_backward_statistics or test_bn_train_roundtrip or test_lf_extract_bit_exact or test_lf_shift_bit_exact or test_pack_views_and_shift_pack or test_heads_against_golden or test_targets_against_golden or test_losses_against_golden or test_ese_reduce_against_golden or test_adam_against_golden or test_weight_pack_variants or test_texture_mask or test_augmentation_chain'
SEL_RACE='test_conv_tc_vs_oracle or test_conv_wgrad or test_conv_dgrad_and_gate or test_bn_train_roundtrip or test_lf_shift_bit_exact or test_heads_against_golden or test_losses_against_golden'
P="python -m pytest -q -m gpu -x -p no:cacheprovider tests/test_gpu_kernels.py"
timeout $TM compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 30 --log-file $O/sanitizer_memcheck_$TAG.log \
  $P -k "$SEL_MEM" > $O/sanitizer_memcheck_pytest_$TAG.log 2>&1
echo "memcheck rc: $?"; tail -n 3 $O/sanitizer_memcheck_pytest_$TAG.log; tail -n 4 $O/sanitizer_memcheck_$TAG.log
timeout $TR compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 30 --log-file $O/sanitizer_racecheck_$TAG.log \
  $P -k "$SEL_RACE" > $O/sanitizer_racecheck_pytest_$TAG.log 2>&1
echo "racecheck rc: $?"; tail -n 3 $O/sanitizer_racecheck_pytest_$TAG.log; tail -n 4 $O/sanitizer_racecheck_$TAG.log
