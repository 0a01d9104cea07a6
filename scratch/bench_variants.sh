#!/bin/bash
for v in "$@"; do
  cp scratch/lib_$v.so ppcseq_b200/libppcseq_b200.so
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$v', 'ms_per_step', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4))
    else: print(l.rstrip())
"
done
