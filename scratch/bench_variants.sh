#!/bin/bash
# usage: bench_variants.sh "<extra bench args>" v1 v2 ...   (restores the last variant as the live library)
extra="$1"; shift
for v in "$@"; do
  cp scratch/lib_$v.so ppcseq_b200/libppcseq_b200.so
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras $extra 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$v', '$extra', 'ms_per_step', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']))
    else: print(l.rstrip())
"
done
