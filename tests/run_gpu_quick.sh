#!/bin/bash
# Quick GPU check: kernel + model parity suites as separate processes, conv cycle stats, short benches.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
P="python -m pytest -q -m gpu -p no:cacheprovider -x"
timeout 900 $P tests/test_gpu_kernels.py > $O/q_kernels.log 2>&1; echo "kernels: $?"; tail -n 15 $O/q_kernels.log
timeout 900 $P tests/test_gpu_model.py > $O/q_model.log 2>&1; echo "model: $?"; tail -n 15 $O/q_model.log
timeout 300 python tests/gpu_conv_stats.py 1 512 512 280 280 0 > $O/conv_stats.txt 2>&1
timeout 300 python tests/gpu_conv_stats.py 64 96 96 280 280 1 >> $O/conv_stats.txt 2>&1
timeout 300 python tests/gpu_conv_stats.py 1 512 512 70 70 1 >> $O/conv_stats.txt 2>&1
cat $O/conv_stats.txt
timeout 600 python bench.py --steps 5 --no-cpu-baseline > $O/q_bench_train.json 2> $O/q_bench_train.err; echo "bench train: $?"; tail -c 400 $O/q_bench_train.err
timeout 600 python bench.py --workload infer --no-cpu-baseline > $O/q_bench_infer.json 2> $O/q_bench_infer.err; echo "bench infer: $?"; tail -c 400 $O/q_bench_infer.err
timeout 600 python bench.py --bs 64 --steps 5 --no-cpu-baseline > $O/q_bench_bs64.json 2> $O/q_bench_bs64.err; echo "bench bs64: $?"
python - <<'PY'
import json
for f in ['q_bench_train','q_bench_infer','q_bench_bs64']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['value'],1), d['unit'], round(d['ms_per_step'],2),'ms', 'conv TF', round(d['roofline']['achieved'],1), 'e2e', round(d['e2e']['value'],1), d['clocks'])
        print('   ', d['kernel_ms_per_step'])
    except Exception as e:
        print(f, 'ERR', e)
PY
