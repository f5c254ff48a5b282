#!/bin/bash
# Conv + wgrad parity tests, role cycle counters, model tests, short benches (no ncu).
cd "$(dirname "$0")/.."
TAG=${1:-it}
O=gpurun_out; mkdir -p $O
P="python -m pytest -q -m gpu -p no:cacheprovider -x"
timeout 600 $P tests/test_gpu_kernels.py -k "conv" > $O/q_kernels.log 2>&1; echo "conv tests: $?"; tail -n 3 $O/q_kernels.log
for a in "1 512 512 280 280 0" "64 96 96 280 280 1" "1 512 512 70 70 1"; do timeout 300 python tests/gpu_conv_stats.py $a; done > $O/conv_stats_$TAG.txt 2>&1
grep "280\|70->\|mma\|TFLOP" $O/conv_stats_$TAG.txt
timeout 900 $P tests/test_gpu_model.py > $O/q_model.log 2>&1; echo "model: $?"; tail -n 3 $O/q_model.log
timeout 600 python bench.py --steps 5 --no-cpu-baseline > $O/q_bench_train.json 2> $O/q_bench_train.err; echo "bench train: $?"; tail -c 300 $O/q_bench_train.err
timeout 600 python bench.py --workload infer --no-cpu-baseline > $O/q_bench_infer.json 2> $O/q_bench_infer.err; echo "bench infer: $?"
timeout 600 python bench.py --bs 64 --steps 5 --no-cpu-baseline > $O/q_bench_bs64.json 2> $O/q_bench_bs64.err; echo "bench bs64: $?"
python - <<'PY'
import json
for f in ['q_bench_train','q_bench_infer','q_bench_bs64']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        r=d['roofline']
        print(f, round(d['value'],1), d['unit'], round(d['ms_per_step'],2),'ms', 'conv TF', round(r['achieved'],1), 'wgrad TF', round(r.get('wgrad',{}).get('achieved',0),1), 'e2e', round(d['e2e']['value'],1), d['clocks'])
        print('   ', d['kernel_ms_per_step'])
    except Exception as e:
        print(f, 'ERR', e)
PY
