#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider -x tests/test_gpu_model.py 2>&1 | tail -n 3
for ov in 1 0; do
  for bs in 512 64; do
    MMLF_BN_FUSE=$ov timeout 600 python bench.py --bs $bs --steps 6 --no-cpu-baseline > $O/ov_${ov}_$bs.json 2> $O/ov_${ov}_$bs.err
    python - <<PY
import json
d=json.loads(open('gpurun_out/ov_${ov}_$bs.json').read().strip().splitlines()[-1])
print('BN_FUSE=$ov bs=$bs', round(d['value'],1), 'patches/s', round(d['ms_per_step'],2), 'ms host', d['host_enqueue_ms_per_step'], 'conv', round(d['roofline']['achieved'],1), 'wgrad', round(d['roofline']['wgrad']['achieved'],1), 'e2e', round(d['e2e']['value'],1), d['clocks'])
PY
  done
done
