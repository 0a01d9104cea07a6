#!/bin/bash
# usage: mgpu_check.sh N  -- dist tests (N=2 only) + bench at N GPUs, flushed and warm L2
N=$1
if [ "$N" = "2" ]; then python -m pytest tests/test_dist_gpu.py -m gpu -x -q 2>&1 | tail -2; fi
for f in "" "--no-flush"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 30 --warmup 5 $f 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N=$N', '$f', 'ms_per_step', round(d['ms_per_step'],4), 'evals/s', round(d['value']), d['step_ms'], 'e2e', round(d['e2e']['value']))"
done
