#!/bin/bash
# model tests + short benches (no ncu).
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
P="python -m pytest -q -m gpu -p no:cacheprovider -x"
timeout 900 $P tests/test_gpu_model.py > $O/q_model.log 2>&1; echo "model: $?"; tail -n 3 $O/q_model.log
timeout 600 python bench.py --steps 5 --no-cpu-baseline > $O/q_bench_train.json 2> $O/q_bench_train.err; echo "bench train: $?"; tail -c 300 $O/q_bench_train.err
timeout 600 python bench.py --workload infer --no-cpu-baseline > $O/q_bench_infer.json 2> $O/q_bench_infer.err; echo "bench infer: $?"; tail -c 300 $O/q_bench_infer.err
timeout 600 python bench.py --workload ese --steps 3 --no-cpu-baseline > $O/q_bench_ese.json 2> $O/q_bench_ese.err; echo "bench ese: $?"; tail -c 300 $O/q_bench_ese.err
timeout 600 python bench.py --bs 64 --steps 5 --no-cpu-baseline > $O/q_bench_bs64.json 2> $O/q_bench_bs64.err; echo "bench bs64: $?"
python - <<'PY'
import json
for f in ['q_bench_train','q_bench_infer','q_bench_ese','q_bench_bs64']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        r=d['roofline']
        print(f, round(d['value'],2), d['unit'], round(d['ms_per_step'],2),'ms host', d.get('host_enqueue_ms_per_step'), 'conv TF', round(r['achieved'],1), 'wgrad TF', round(r.get('wgrad',{}).get('achieved',0),1), 'e2e', round(d['e2e']['value'],2), 'launches', d['gpu_launches'], d['clocks'])
        print('   ', d['kernel_ms_per_step'])
    except Exception as e:
        print(f, 'ERR', e)
PY
