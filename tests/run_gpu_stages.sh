#!/bin/bash
# Runs the -m gpu suites as separate processes (a trapped kernel poisons only its own stage); logs to gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
P="python -m pytest -q -m gpu -p no:cacheprovider"
timeout 600 $P tests/test_gpu_kernels.py -k "lf_ or pack_views or simt or weight_pack or bn_ or heads or targets or ese or losses or ce_ or adam" > gpurun_out/stage_a.log 2>&1; echo "stage a: $?"
timeout 300 $P tests/test_gpu_kernels.py -k "tc_vs_oracle" > gpurun_out/stage_b.log 2>&1; echo "stage b: $?"
timeout 300 $P tests/test_gpu_kernels.py -k "tc_epilogues or many_tiles or dgrad" > gpurun_out/stage_c.log 2>&1; echo "stage c: $?"
timeout 300 $P tests/test_gpu_kernels.py -k "wgrad" > gpurun_out/stage_d.log 2>&1; echo "stage d: $?"
timeout 900 $P tests/test_gpu_model.py > gpurun_out/stage_e.log 2>&1; echo "stage e: $?"
tail -n 5 gpurun_out/stage_*.log
