#!/bin/bash
# N-GPU session: NCCL parity of the sharded path, then bench at N (strong C3 + weak) ; logs kept small
N=${1:-2}
TAG=${2:-r2x}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 scripts/check_dist_parity.py > $O/${TAG}_dist_parity_${N}gpu.log 2>&1; echo "parity rc=$?"
grep -E "OK|MISMATCH|PASS|FAIL|Error|error" $O/${TAG}_dist_parity_${N}gpu.log | tail -20
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 10 --warmup 3 > $O/${TAG}_bench_${N}gpu.json 2> $O/${TAG}_bench_${N}gpu.err; echo "bench rc=$?"
tail -3 $O/${TAG}_bench_${N}gpu.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("$O/${TAG}_bench_${N}gpu.json") if l.startswith("{")][-1])
    print("N", d["n_gpus"], d["scaling"], "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "serial", round(d["serial"]["value"], 1), round(d["serial"]["ms_per_step"], 4))
    print("stages", {k: v["ms"] for k, v in d["stages"].items()})
    print("weak", d.get("weak"))
    print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "ascii", d["e2e"]["ascii"]["value"])
except Exception as e:
    print("bench parse failed", e)
PY
if [ "$3" == "c5" ]; then
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --config c5 --steps 10 --warmup 3 --no-e2e > $O/${TAG}_bench_c5_${N}gpu.json 2> $O/${TAG}_bench_c5_${N}gpu.err; echo "c5 rc=$?"
tail -3 $O/${TAG}_bench_c5_${N}gpu.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("$O/${TAG}_bench_c5_${N}gpu.json") if l.startswith("{")][-1])
    print("C5 N", d["n_gpus"], d["scaling"], "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "serial", round(d["serial"]["value"], 1), round(d["serial"]["ms_per_step"], 4))
    print("stages", {k: v["ms"] for k, v in d["stages"].items()})
except Exception as e:
    print("c5 parse failed", e)
PY
fi
du -sh $O
