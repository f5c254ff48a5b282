#!/bin/bash
# Conv-kernel iteration loop: conv parity tests, role cycle counters for three layer shapes, one ncu capture.
cd "$(dirname "$0")/.."
TAG=${1:-it}
O=gpurun_out; mkdir -p $O
P="python -m pytest -q -m gpu -p no:cacheprovider -x"
timeout 600 $P tests/test_gpu_kernels.py -k "conv" > $O/q_kernels.log 2>&1; echo "conv tests: $?"; tail -n 4 $O/q_kernels.log
for a in "1 512 512 280 280 0" "64 96 96 280 280 1" "1 512 512 70 70 1" "1 512 512 27 70 0"; do timeout 300 python tests/gpu_conv_stats.py $a; done > $O/conv_stats_$TAG.txt 2>&1
cat $O/conv_stats_$TAG.txt
if [ "$2" != "noncu" ]; then bash tests/run_ncu_conv.sh $TAG; fi
