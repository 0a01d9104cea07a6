#!/bin/bash
# A/B timing of kernel variants on the GPU box: builds of the library under gpurun_ab/ (made by build_variant.sh here),
# each timed cold (L2 flushed) and warm, interleaved so that box-to-box and time drift cancel.
# usage: ab.sh [bench args...]
cd "$(dirname "$0")/../.."
for rep in 1 2; do
for so in gpurun_ab/*.so; do
  for f in "" "--no-flush"; do
    PPCSEQ_B200_LIB=$PWD/$so python bench.py --steps 40 --warmup 5 --no-extras --no-cpu-baseline --batch 1 $f "$@" | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$so', '$f' or 'cold', round(d['ms_per_step']*1e3,2), 'us  median', round(d['step_ms']['median']*1e3,2))"
  done
done
done
