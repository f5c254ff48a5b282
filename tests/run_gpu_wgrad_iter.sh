#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider -x tests/test_gpu_kernels.py -k "wgrad" 2>&1 | tail -n 3
for g in 1 2 0; do echo "GROUP=$g (0 = default rule)"; MMLF_WGRAD_GROUP=$g python tools/kernel_bench.py --only "wgrad" --reps 30 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['case'], round(d['ms'],3), round(d['achieved'],1), round(d['frac'],2))
"; done
