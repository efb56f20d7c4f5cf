#!/bin/bash
TAG=${1:-r2x}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/${TAG}_pytest.log
echo "== tridiag fused (production timing)"; AM_TRIDIAG_PROBE= python scripts/dev_eig_once.py 2>&1 | head -1
echo "== tridiag deferred update (production timing)"; AM_TRIDIAG_DEFER=1 AM_TRIDIAG_PROBE= python scripts/dev_eig_once.py 2>&1 | head -1
echo "== tridiag fused (probe)"; python scripts/dev_eig_once.py 2>&1 | tail -9
echo "== tridiag deferred (probe)"; AM_TRIDIAG_DEFER=1 python scripts/dev_eig_once.py 2>&1 | tail -9
python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open("$O/${TAG}_bench.json"))
    print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "stages", {k: v["ms"] for k, v in d["stages"].items()})
    print("serial", d.get("serial"))
    print("e2e", {k: d["e2e"][k] for k in ("value", "ms_per_step", "serial_latency_ms", "h2d_bytes_per_step")}, "ascii", {k: d["e2e"]["ascii"][k] for k in ("value", "ms_per_step")})
except Exception as e:
    print("bench parse failed", e)
PY
for dp in 3 4; do python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e --no-sweep --pipeline-depth $dp 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('depth', d['pipelined']['depth'], 'pipelined ms', round(d['pipelined']['ms_per_step'],4), 'serial ms', round(d['ms_per_step'],4))"; done
du -sh $O
