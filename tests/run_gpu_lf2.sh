#!/bin/bash
cd "$(dirname "$0")/.."
for r in 1 2 4 8; do echo "ROWS=$r"; MMLF_SHIFT_ROWS=$r python tools/kernel_bench.py --only "lf_shift" --reps 50 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['kernel'], '|', d['case'], round(d['ms'],3), round(d['achieved'],1), round(d['frac'],2))
"; done
