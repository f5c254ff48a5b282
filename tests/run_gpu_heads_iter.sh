#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider -x tests/test_gpu_kernels.py -k "heads or targets or losses or ce_ or ese" 2>&1 | tail -n 3
python tools/kernel_bench.py --only "loss_,dpp_head,upr_post,reg_to" --reps 30 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['kernel'], '|', d['case'], round(d['ms'],3), round(d['achieved'],1), round(d['frac'],2))
"
