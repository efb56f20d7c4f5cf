#!/bin/bash
# One GPU-box session: tests, bench, experiments, ncu summaries.  Everything that must come back is SMALL
# (.ncu-rep files are reduced to csv on the box and deleted: gpurun_out/ is capped at 64 MiB).
TAG=${1:-r2x}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q -x > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${TAG}_pytest.log
python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open("$O/${TAG}_bench.json"))
    print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "stages", {k: v["ms"] for k, v in d["stages"].items()})
    print("serial", d.get("serial"))
    print("e2e", {k: d["e2e"][k] for k in ("value", "ms_per_step", "serial_latency_ms", "h2d_bytes_per_step")}, "ascii", {k: d["e2e"]["ascii"][k] for k in ("value", "ms_per_step")})
    print("flush", d.get("l2_flush_check"), "sweep", d["sweep"]["value"], d["sweep"].get("api"))
    print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
except Exception as e:
    print("bench parse failed", e)
PY
if [ "$2" != "short" ]; then
python scripts/dev_plane_sweep.py > $O/${TAG}_plane_sweep.log 2>&1; echo "planes rc=$?"; cat $O/${TAG}_plane_sweep.log | tail -20
python scripts/time_api.py > $O/${TAG}_api.log 2>&1; echo "api rc=$?"; cat $O/${TAG}_api.log | tail -4
fi
python scripts/time_sweep_batch.py > $O/${TAG}_sweep_batch.json 2> $O/${TAG}_sweep_batch.err; echo "sweep batch rc=$?"; cat $O/${TAG}_sweep_batch.json
python scripts/dev_sweep_step_time.py > $O/${TAG}_sweep_step_time.json 2>> $O/${TAG}_sweep_batch.err; cat $O/${TAG}_sweep_step_time.json
# ncu: the neighbour kernel (eps_union_warp_kernel) at eps 0.3 / 2.0 / 5.0 (launches 0, 17, 47 of the sweep)
python scripts/dev_sweep_once.py > $O/${TAG}_sweep_plain.log 2>&1 && for s in 0 17 47; do
  ncu --set full --clock-control none -k regex:eps_union_warp_kernel -s $s -c 1 -o $O/eps_$s python scripts/dev_sweep_once.py > $O/${TAG}_ncu_eps_$s.log 2>&1
  ncu -i $O/eps_$s.ncu-rep --page raw --csv > $O/${TAG}_eps_union_launch${s}_raw.csv 2>/dev/null; rm -f $O/eps_$s.ncu-rep
done; echo "ncu eps rc=$?"
# ncu: launch list of a short bench, and full sets of the count kernel and the tridiagonalisation
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-sweep --pipeline-depth 1 > $O/${TAG}_bench_short.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-sweep --pipeline-depth 1 > $O/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-sweep --pipeline-depth 1 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'count_kernel|tridiag_reg|gram_i8x2|normalize512_planes|project_i8' -s 10 -c 6 -o $O/full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-sweep --pipeline-depth 1 > $O/${TAG}_ncu_full.log 2>&1
ncu -i $O/full.ncu-rep --page raw --csv > $O/${TAG}_full_raw.csv 2>/dev/null
ncu -i $O/full.ncu-rep --page source --csv -k regex:count_kernel > $O/${TAG}_count_source.csv 2>/dev/null
ls -la $O/full.ncu-rep; rm -f $O/full.ncu-rep; echo "ncu full rc=$?"
du -sh $O
