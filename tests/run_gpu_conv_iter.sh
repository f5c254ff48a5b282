#!/bin/bash
# conv / wgrad parity tests + the per-kernel roofline lines of the tensor-core kernels
cd "$(dirname "$0")/.."
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider -x tests/test_gpu_kernels.py -k "conv" 2>&1 | tail -n 4
python tools/kernel_bench.py --only "${1:-conv2x2_tc2}" --reps 30 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['case'], round(d['ms'],3), round(d['achieved'],1), round(d['frac'],2))
"
