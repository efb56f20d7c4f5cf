#!/bin/bash
# depth sweep of the pipelined strong-scaling figure at N GPUs
N=${1:-2}
for dp in ${2:-3 4 5 6 8}; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$dp bench.py --gpus $N --steps 24 --warmup 3 --no-cpu-baseline --no-e2e --no-sweep --no-weak --pipeline-depth $dp 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], 'depth', $dp, 'pipelined ms', round(d['ms_per_step'],4), 'Gbp/s', round(d['value'],1), 'serial ms', round(d['serial']['ms_per_step'],4), {k: v['ms'] for k, v in d['stages'].items()})"
done
