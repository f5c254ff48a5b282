#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
P="python -m pytest -q -m gpu -p no:cacheprovider -x"
for bo in 1 0; do
  echo "== MMLF_CONV_BO=$bo"
  MMLF_CONV_BO=$bo timeout 600 $P tests/test_gpu_kernels.py -k "conv" > $O/ashare_bo$bo.log 2>&1; echo "conv tests: $?"; tail -n 5 $O/ashare_bo$bo.log
done
