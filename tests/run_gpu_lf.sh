#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider -x tests/test_gpu_kernels.py -k "lf_ or pack or shift" 2>&1 | tail -n 8
python tools/kernel_bench.py --only "lf_shift,pack_views,shift_pack,lf_extract" 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['kernel'], '|', d['case'], round(d['ms'],3), round(d['achieved'],1), round(d['frac'],2))
"
