#!/bin/bash
TAG=${1:-r2x}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $O/${TAG}_pytest.log
echo "== count (40 regs, 6 CTAs/SM)"; python scripts/dev_count_time.py 2>&1 | tail -1
echo "== count (48 regs, 5 CTAs/SM)"; AM_COUNT_MINB=5 python scripts/dev_count_time.py 2>&1 | tail -1
for dp in 2 3 4 6 8; do python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-e2e --no-sweep --pipeline-depth $dp 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('depth', $dp, 'pipelined ms', round(d['ms_per_step'],4), 'Gbp/s', round(d['value'],1), 'serial ms', round(d['serial']['ms_per_step'],4), {k: v['ms'] for k, v in d['stages'].items()})"; done
du -sh $O
