#!/bin/bash
# N-GPU checks (gpurun --gpus N): torchrun bench lines for train / ese / infer, as the driver launches them.
#   usage: tests/run_gpu_multi.sh N [tag]
cd "$(dirname "$0")/.."
N=${1:-2}
TAG=${2:-r01}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv,noheader > $O/gpus_n$N.txt
nvidia-smi topo -m >> $O/gpus_n$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR tests/gpu_multi_check.py > $O/multi_check_n${N}_$TAG.txt 2>&1; echo "multi check n=$N: $?"; grep -E "OK|MISMATCH" $O/multi_check_n${N}_$TAG.txt
for wl in train ese bands infer; do
  timeout 900 $TR bench.py --gpus $N --workload $wl --steps 5 --warmup 3 > $O/bench_${wl}_n${N}_$TAG.json 2> $O/bench_${wl}_n${N}_$TAG.err
  echo "bench $wl n=$N: $?"; tail -c 400 $O/bench_${wl}_n${N}_$TAG.err
done
timeout 600 $TR bench.py --gpus $N --impl reference --steps 1 --warmup 1 > $O/bench_reference_n${N}_$TAG.json 2> $O/bench_reference_n${N}_$TAG.err; echo "reference n=$N: $?"
python - <<PY
import json
for wl in ['train','ese','infer','reference']:
    try:
        d=json.loads(open('gpurun_out/bench_%s_n${N}_${TAG}.json' % wl).read().strip().splitlines()[-1])
        print(wl, round(d['value'],3), d['unit'], round(d['ms_per_step'],2),'ms e2e', round(d['e2e']['value'],3), d.get('clocks'))
        if 'kernel_ms_per_step' in d: print('   ', d['kernel_ms_per_step'])
    except Exception as e:
        print(wl, 'ERR', e)
PY
