#!/bin/bash
# ncu full capture (with SASS source counters) of one 280->280 conv launch.  usage: run_ncu_conv.sh tag [B H W cin cout type]
cd "$(dirname "$0")/.."
TAG=${1:-x}; shift
ARGS=${@:-1 512 512 280 280 0}
O=gpurun_out; mkdir -p $O
CMD="python tests/gpu_conv_stats.py $ARGS"
$CMD > $O/ncu_conv_plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv2x2_tc2 -s 3 -c 1 -f -o $O/prof_conv_$TAG $CMD > $O/ncu_conv_$TAG.log 2>&1
tail -n 3 $O/ncu_conv_plain_$TAG.log $O/ncu_conv_$TAG.log
