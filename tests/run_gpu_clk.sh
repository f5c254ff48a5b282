cd /root/repo
for reps in 3 300 3000; do echo "REPS=$reps"; CONV_STATS_REPS=$reps python tests/gpu_conv_stats.py 64 96 96 280 280 1 | grep -v "wait\|total"; done
CONV_STATS_REPS=3000 python tests/gpu_conv_stats.py 1 512 512 280 280 0
