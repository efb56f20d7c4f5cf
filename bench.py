#!/usr/bin/env python
"""bench.py — featurise throughput (5-mer count + am_clr + PCA(50)) on synthetic assemblies.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W     # CPU port of the reference path

A step = one pass of the hot path (count -> am_clr -> PCA fit + scores) over the rank's
synthetic contig shard.  N=1 workload: BASELINE.json configs[1] (C2: 200k contigs / 1 Gbp).
N>1: one process per GPU (torchrun); by default the SAME 1 Gbp assembly is sharded by bp across
the ranks (BASELINE.json configs[2], C3: strong scaling), and the weak-scaling figure (every rank
its own 200k-contig / 1 Gbp shard) is measured afterwards and reported under `weak`.  The only
cross-GPU traffic is ONE all-reduce per step: [512x512 Gram | column sums | row count | column
presence | all-NA row count] (NCCL).  `--config c5` = BASELINE.json configs[4]: 1M contigs /
5 Gbp, ilr, strong scaling over the ranks.

`value`: whole-job Gbp/s with the 2-bit-packed contigs already in HBM, CUDA events, max over
ranks.  `e2e`: the same metric from pinned HOST ASCII bytes through H2D, the device packer,
the pipeline and the D2H copy of the scores.  Inputs/outputs per step (0.25 GB codes, 0.41 GB
counts, 0.82 GB normalised) exceed the 126 MB L2, so no explicit L2 flush is needed.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "featurise throughput (5-mer count + am_clr + PCA50)"
UNIT = "Gbp/s"
DTYPE = "int32 counts / f64 features (Gram + projection: exact int8 digit planes on tcgen05, fp64 result)"
N_CONTIGS = 200_000
TOTAL_BP = 1_000_000_000
SEED = 20261019 + 2
K = 5
NCOMP = 50
# the same string in both arms (the driver compares `config.workload` of the two lines)
WORKLOADS = {
    "c1": "C1: 5-mer count + am_clr + PCA(50), 10k contigs / 50 Mbp seeded synthetic assembly",
    "c2": "C2: 5-mer count + am_clr + PCA(50), 200k contigs / 1 Gbp seeded synthetic assembly",
    "c5": "C5: 5-mer count + ilr + PCA(50), 1M contigs / 5 Gbp seeded synthetic assembly (large-data mode)",
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops", 1590.0)), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None, t_load=None):
        """Median SM clock over the samples taken inside [t0, t1] (the timed region); when the
        region is shorter than the sampling period, over the samples since `t_load` (first warm-up)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        window = "timed region"
        rows = [ln for (ts, ln) in self.lines if t0 is None or (t0 - 0.02 <= ts <= t1 + 0.12)]
        if len(rows) < 3 and t_load is not None:
            rows = [ln for (ts, ln) in self.lines if ts >= t_load]
            window = "warm-up + timed region (timed region shorter than 3 samples)"
        for ln in rows:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(mx)) if mx else None,
            "samples": len(sm),
            "window": window,
            "reasons": sorted(reasons),
        }


def cpu_port_step(asm, threads):
    """One pass of the CPU port (oracle) over `asm`: C count + C am_clr + sklearn PCA(50)."""
    from sklearn.decomposition import PCA

    from oracle import oracle_c

    counts = oracle_c.count(asm.seq, asm.offsets, K, threads)
    keep = counts.sum(axis=1) > 0
    X = oracle_c.am_clr(counts if keep.all() else counts[keep], threads)
    p = PCA(n_components=NCOMP, random_state=np.random.RandomState(42))
    scores = p.fit_transform(X)
    return scores, p.explained_variance_


def cpu_baseline(asm_full, threads, sample_bp=1_000_000_000, passes=3):
    """Oracle port timed on a bounded sample (first contigs totalling ~sample_bp)."""
    n = int(np.searchsorted(asm_full.offsets, asm_full.offsets[0] + sample_bp))
    n = max(10 * 512 + 1, min(n, asm_full.n_contigs))
    sub = asm_full.slice(0, n)
    cpu_port_step(sub.slice(0, min(n, 6000)), threads)  # warm the BLAS/thread pools
    t0 = time.perf_counter()
    for _ in range(passes):
        cpu_port_step(sub, threads)
    dt = (time.perf_counter() - t0) / passes
    return {
        "value": sub.total_bp / dt / 1e9,
        "unit": UNIT,
        "cores": threads,
        "kind": "port",
        "sample": f"first {n} contigs / {sub.total_bp / 1e6:.1f} Mbp of the workload: C port of the count + am_clr "
                  f"loops on {threads} threads + sklearn PCA(50) (covariance_eigh), mean of {passes} passes, {dt:.2f} s each",
    }


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; the Python reference cannot
    travel to the GPU box) on the host cores, same metric/config, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from autometa_b200 import synth

    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    sample_contigs, sample_bp = 200_000, 1_000_000_000  # the whole C2 workload per step (about 1 s of host time)
    asm = synth.make_assembly(sample_contigs, sample_bp, seed=SEED)
    # torchrun exports OMP_NUM_THREADS=1 to its workers: lift the BLAS/OpenMP pools back to the host's
    # cores so that the CPU arm is the same whether it is launched directly or under torchrun
    from threadpoolctl import threadpool_limits

    with threadpool_limits(limits=threads):
        for _ in range(max(args.warmup, 1)):
            cpu_port_step(asm.slice(0, 6000), threads)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_port_step(asm, threads)
        dt = (time.perf_counter() - t0) / args.steps
    val = asm.total_bp / dt / 1e9
    sample = (f"{sample_contigs} contigs / {sample_bp / 1e6:.0f} Mbp per step (the whole 200k-contig / 1 Gbp "
              f"workload): C port of the reference count + am_clr loops on {threads} threads + sklearn PCA(50)")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": {"workload": WORKLOADS["c2"], "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def sweep_bench(dev, torch, cpu=True, reps=3):
    """Second headline metric (BASELINE.json configs[3]): seconds for the reference's
    48-value eps grid (0.3 .. 5.0, recursive_dbscan.py:172-173,215) on 200k points of
    [x_1, x_2, coverage], producing the sklearn-identical int64 label vector at EVERY eps.
    Features are resident in HBM; CUDA events; median of `reps` sweeps after one warm-up sweep."""
    from autometa_b200 import device as D
    from autometa_b200 import synth
    from oracle import oracle_np as O

    table, _ = synth.make_sweep_input()
    X = np.ascontiguousarray(table[["x_1", "x_2", "coverage"]].to_numpy())
    Xd = torch.from_numpy(X).to(dev)
    grid = O.eps_schedule(48)
    sw = D.EpsSweep(Xd)
    L0 = None
    times = []
    ncl = []
    from autometa_b200 import _lib as L

    for rep in range(reps + 1):
        sw.reset()
        torch.cuda.synchronize()
        l0 = L.launch_count()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for eps in grid:
            sw.step(eps)
        e1.record()
        torch.cuda.synchronize()
        if rep:
            times.append(e0.elapsed_time(e1) * 1e-3)
        L0 = L.launch_count() - l0
        ncl = [int(sw.n_clusters.item())]
    t = float(np.median(times))
    n = X.shape[0]
    out = {
        "metric": "DBSCAN eps-sweep (48 eps, 200k points x [x_1,x_2,coverage], labels at every eps)",
        "value": t, "unit": "s", "higher_is_better": False, "ms_per_eps": t / len(grid) * 1e3,
        "n_points": n, "n_eps": len(grid), "clusters_at_last_eps": ncl[0], "gpu_launches": int(L0),
        "algorithmic_bytes_per_eps": 8 * 3 * n + 8 * n,
        "achieved_GBps_algorithmic": (8 * 3 * n + 8 * n) * len(grid) / t / 1e9,
        "note": "latency/instruction-bound (6.4 MB per eps is HBM-trivial); incremental union-find forest across "
                "eps; neighbour pass = one warp per point over a hashed cell list (eps_union_warp_kernel)",
    }
    # the same sweep through the reference-facing API (recursive_dbscan.py:114-234 semantics: eps
    # schedule until one cluster remains, per-eps metrics + cutoff filter + best-eps rule), host
    # DataFrames in, (clustered_df, unclustered_df) out, metrics on the device
    try:
        from autometa_b200.binning import recursive_dbscan as R

        _, markers = synth.make_sweep_input()
        tb = table.copy()
        R.recursive_dbscan(tb, markers, 20.0, 95.0, 25.0, 5.0)  # warm-up
        tb = table.copy()
        t0 = time.perf_counter()
        cl, un = R.recursive_dbscan(tb, markers, 20.0, 95.0, 25.0, 5.0)
        out["api"] = {"value": time.perf_counter() - t0, "unit": "s",
                      "call": "recursive_dbscan(table, markers_df, 20, 95, 25, 5) incl. DataFrame in/out, all eps "
                              "until one cluster remains; eps step, add_metrics / filter / median AND the loop's "
                              "decisions (best eps, step x 10, termination) on the device: one captured CUDA graph "
                              "per step, replayed in batches of 16, one word read back per batch (csrc/am_sweep.cu)",
                      "clustered": int(cl.shape[0]), "unclustered": int(un.shape[0])}
    except Exception as exc:  # keep the bench line even if the API leg fails
        out["api"] = {"error": repr(exc)}
    # the callers' shape of the same work (large-data mode): many small get_clusters problems — the
    # taxa of one rank — one at a time against in flight together (get_clusters_many)
    try:
        import pandas as pd

        from autometa_b200.binning import recursive_dbscan as R

        parts, per = 32, 2000
        tables, mk = [], []
        for i in range(parts):
            t, m = synth.make_sweep_input(n_points=per, n_clusters=6, n_markers=40, seed=100 + i)
            t.index = pd.Index([f"T{i}_{c}" for c in t.index], name="contig")
            m.index = pd.Index([f"T{i}_{c}" for c in m.index], name="contig")
            tables.append(t)
            mk.append(m)
        mk = pd.concat(mk)
        cut = (20.0, 90.0, 25.0, 5.0)
        R.get_clusters_many([t.copy() for t in tables[:8]], mk, *cut, method="dbscan")  # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        one = [R.get_clusters(t.copy(), mk, *cut, method="dbscan") for t in tables]
        torch.cuda.synchronize()
        t_one = time.perf_counter() - t0
        t0 = time.perf_counter()
        many = R.get_clusters_many([t.copy() for t in tables], mk, *cut, method="dbscan")
        torch.cuda.synchronize()
        t_many = time.perf_counter() - t0
        same = all(a["cluster"].fillna("").tolist() == b["cluster"].fillna("").tolist() for a, b in zip(one, many))
        out["taxa_batch"] = {
            "what": f"get_clusters on {parts} tables of {per} contigs (DataFrames in and out, all rounds): one call "
                    "after the other / get_clusters_many (sweeps in flight together, each a chain of CUDA-graph "
                    "replays on its own stream)",
            "one_at_a_time_s": t_one, "in_flight_together_s": t_many,
            "ms_per_table": {"one_at_a_time": t_one / parts * 1e3, "in_flight_together": t_many / parts * 1e3},
            "bins": int(sum(x["cluster"].nunique() for x in many)), "same_bins": bool(same),
        }
    except Exception as exc:
        out["taxa_batch"] = {"error": repr(exc)}
    if cpu:
        # reference arm of the same metric: what recursive_dbscan.py:109 executes, on the host cores,
        # at 4 of the 48 eps values (bounded sample), scaled to the grid
        from sklearn.cluster import DBSCAN

        idx = [0, 15, 31, 47]
        t0 = time.perf_counter()
        for i in idx:
            DBSCAN(eps=grid[i], min_samples=1, n_jobs=-1).fit(X)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {
            "value": dt / len(idx) * len(grid), "unit": "s", "cores": os.cpu_count(), "kind": "reference",
            "sample": f"sklearn DBSCAN(min_samples=1, n_jobs=-1) at eps indices {idx} of the 48-value grid "
                      f"({dt:.2f} s), scaled x{len(grid) / len(idx):.0f}",
        }
    return out


STAGE_BYTES = None


def stage_accounting(n_contigs, total_bp, D=512, P=NCOMP, Dout=512):
    """Algorithmic bytes / flops per launch, for this rank's shard (Dout = 511 for ilr)."""
    return {
        "count": {"bytes": total_bp / 4 + 8 * (n_contigs + 1) + 4 * D * n_contigs, "bound": "hbm"},
        "normalise": {"bytes": 4 * D * n_contigs + 8 * Dout * n_contigs, "bound": "hbm"},
        "gram": {"bytes": 8 * Dout * n_contigs, "flops": 2.0 * n_contigs * Dout * Dout, "bound": "tensor"},
        "eigh": {"bytes": 2 * 8 * Dout * Dout, "bound": "latency"},
        "project": {"bytes": 8 * Dout * n_contigs + 8 * P * n_contigs, "flops": 2.0 * n_contigs * Dout * P,
                    "bound": "hbm"},
    }


def make_shard(args, synth, DD, world, rank, scaling):
    """This rank's contigs.  strong: the ONE seeded assembly, sharded by bp (identical data at every N);
    weak: a rank-own assembly of the full size; c5: rank-own 1/N shares of the 1M-contig / 5 Gbp job
    (the 5 GB ASCII assembly is not built 8 times on one host)."""
    if args.config == "c5":
        n = args.contigs // world + (1 if rank < args.contigs % world else 0)
        bp = args.bp // world
        return synth.make_assembly(n, bp, seed=20261019 + 5 + 1000 * rank)
    if scaling == "weak" or world == 1:
        return synth.make_assembly(args.contigs, args.bp, seed=SEED + 1000 * rank)
    full = synth.make_assembly(args.contigs, args.bp, seed=SEED)
    lo, hi = DD.shard_ranges(full.lengths, world)[rank]
    return full.slice(lo, hi)


STAGES = ["count", "normalise", "gram", "eigh", "project"]


def time_steps(torch, dist, pipeline, L, packed, norm, steps, warmup, distributed, dev, sampler=None, rank=0,
               flush=None):
    """W untimed steps, then EXACTLY `steps` timed ones between barrier + synchronize on both sides; CUDA
    events on the launching stream; max over ranks.  `flush`: a buffer larger than L2 rewritten before
    every step, with per-step events so that the flush itself is not timed."""

    def run_step(marks=None):
        return pipeline.featurise(packed, K, norm, NCOMP, distributed=distributed, keep=False, marks=marks)

    t_load = time.time()
    for _ in range(warmup):
        run_step()
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    torch.cuda.synchronize()
    step_marks = []
    launches0 = L.launch_count()
    t_timed0 = time.time()
    if flush is None:
        ev_start = torch.cuda.Event(enable_timing=True)
        ev_end = torch.cuda.Event(enable_timing=True)
        ev_start.record()
        for _ in range(steps):
            marks = []
            run_step(lambda name: marks.append((name, _rec(torch))))
            step_marks.append(marks)
        ev_end.record()
        torch.cuda.synchronize()
        ms_total = ev_start.elapsed_time(ev_end)
    else:
        pairs = []
        for _ in range(steps):
            flush.add_(1)
            a, b = _rec(torch), None
            marks = []
            run_step(lambda name: marks.append((name, _rec(torch))))
            b = _rec(torch)
            pairs.append((a, b))
            step_marks.append(marks)
        torch.cuda.synchronize()
        ms_total = sum(a.elapsed_time(b) for a, b in pairs)
    if distributed:
        dist.barrier()
    torch.cuda.synchronize()
    launches = L.launch_count() - launches0
    clocks = sampler.stop(t_timed0, time.time(), t_load) if (sampler is not None and rank == 0) else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / steps
    stage_ms = {s: 0.0 for s in STAGES}
    for marks in step_marks:
        d = dict(marks)
        order = ["start"] + STAGES
        for a, b in zip(order[:-1], order[1:]):
            stage_ms[b] += d[a].elapsed_time(d[b])
    stage_ms = {s: v / steps for s, v in stage_ms.items()}
    return ms_per_step, stage_ms, launches, clocks


def time_pipelined(torch, dist, pipeline, L, packed, norm, steps, warmup, distributed, dev, depth):
    """Throughput mode (pipeline.FeatureStream): `depth` steps in flight, stage 1 (count -> normalise -> Gram) of
    step i+1 on one stream while stage 2 (eigensolve + projection) of step i runs on another; with several GPUs the
    eigensolve of step i is done by rank i mod N and broadcast.  Every step's work is complete inside the timed
    region (drain() before the closing event); barrier + synchronize on both sides; max over ranks."""
    consumed = []  # the consumer keeps only the last pass's outputs (a service hands them on and lets go)

    def consume(res):
        consumed[:] = [res["scores"], res["explained_variance"]]

    fs = pipeline.FeatureStream(dev, K, norm, NCOMP, distributed=distributed, depth=depth, on_result=consume)
    for _ in range(max(warmup, depth + 1)):  # every buffer of every step in flight exists before the timed region
        fs.submit(packed)
    fs.drain()
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = L.launch_count()
    t0 = time.time()
    fs.host_queue_s = fs.host_wait_s = 0.0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        fs.submit(packed)
    fs.drain()
    ev1.record()
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    torch.cuda.synchronize()
    t1 = time.time()
    launches = L.launch_count() - l0
    tp = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
    host = {"queue_ms_per_step": round(fs.host_queue_s / steps * 1e3, 4),
            "wait_ms_per_step": round(fs.host_wait_s / steps * 1e3, 4),
            "what": "rank 0's host time per step spent queueing work (Python + launches + NCCL enqueues) and waiting "
                    "for the oldest step in flight; queue_ms ~ ms_per_step means the stream is launch-bound"}
    del fs
    return float(tp.item()) / steps, launches, t0, t1, host


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c5"],
                    help="c2: BASELINE.json configs[1]/[2] (200k contigs / 1 Gbp, am_clr); c1: configs[0] (10k contigs / "
                         "50 Mbp, the reference's CPU-runnable case: L2-resident, a parity config rather than a roofline "
                         "one); c5: configs[4] (1M contigs / 5 Gbp, ilr, sharded over the ranks)")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"],
                    help="N>1 default: strong (the same assembly sharded by bp, BASELINE.json configs[2]); the weak "
                         "figure is then measured too and reported under `weak`")
    ap.add_argument("--contigs", type=int, default=None)
    ap.add_argument("--bp", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--no-weak", action="store_true", help="N>1: skip the extra weak-scaling measurement")
    ap.add_argument("--norm", default=None, choices=["am_clr", "clr", "ilr"])
    ap.add_argument("--pipeline-depth", type=int, default=0,
                    help="steps in flight in the timed region (0 = max(3, N + 2); 1 = strictly one step at a time)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    c5 = args.config == "c5"
    if args.contigs is None:
        args.contigs = 1_000_000 if c5 else (10_000 if args.config == "c1" else N_CONTIGS)
    if args.bp is None:
        args.bp = 5_000_000_000 if c5 else (50_000_000 if args.config == "c1" else TOTAL_BP)
    if args.norm is None:
        args.norm = "ilr" if c5 else "am_clr"

    import torch
    import torch.distributed as dist

    from autometa_b200 import _lib as L
    from autometa_b200 import assembly, dist as DD, pipeline, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    args.warmup = max(args.warmup, 3)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    scaling = args.scaling or ("strong" if (distributed or c5) else "weak")
    numa_bound = False
    if distributed:  # one process per GPU: keep each rank's pinned staging buffers on the GPU's own socket
        numa_bound = DD.bind_to_gpu_numa_node(local_rank)
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # ---- synthetic shard for this rank
    asm = make_shard(args, synth, DD, world, rank, scaling)
    packed = assembly.pack_to_device(asm, device=dev, on_device=True)
    shard_bp, shard_n = asm.total_bp, asm.n_contigs
    job = torch.tensor([float(shard_bp), float(shard_n)], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(job)
    job_bp, job_n = float(job[0].item()), float(job[1].item())

    stages = STAGES
    # clocks are sampled from the first warm-up step to the end of the timed region (nvidia-smi needs
    # ~0.3 s to start and samples every 100 ms, so a short timed region alone could fall between samples)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.5)
    t_load = time.time()
    ms_serial, stage_ms, launches_serial, _ = time_steps(torch, dist, pipeline, L, packed, args.norm, args.steps,
                                                         args.warmup, distributed, dev)
    depth = args.pipeline_depth if args.pipeline_depth > 0 else max(3, world + 2)
    if depth > 1:
        ms_per_step, launches, tt0, tt1, host_acct = time_pipelined(torch, dist, pipeline, L, packed, args.norm,
                                                                    args.steps, args.warmup, distributed, dev, depth)
    else:
        ms_per_step, launches, tt0, tt1, host_acct = ms_serial, launches_serial, t_load, time.time(), None
    clocks = sampler.stop(tt0, tt1, t_load) if rank == 0 else None
    value = job_bp / (ms_per_step * 1e-3) / 1e9
    serial = {"value": job_bp / (ms_serial * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms_serial,
              "gpu_launches": int(launches_serial),
              "what": "one step at a time on one stream (the latency of a single featurise call); the per-stage "
                      "times under `stages` come from this run"}

    # the same steps with a 512 MB buffer rewritten before each one (L2 = 126 MB): states once what the
    # missing flush is worth; the headline keeps back-to-back steps (inputs and outputs exceed L2)
    flush_check = None
    if not c5:
        fl = torch.zeros(512 << 20, dtype=torch.uint8, device=dev)
        ms_f, _, _, _ = time_steps(torch, dist, pipeline, L, packed, args.norm, max(3, min(args.steps, 5)), 1,
                                   distributed, dev, flush=fl)
        del fl
        flush_check = {"serial_ms_per_step_with_l2_flush": round(ms_f, 4), "serial_ms_per_step": round(ms_serial, 4)}

    # ---- N>1: the weak-scaling figure next to the strong one (every rank its own full-size assembly)
    weak = None
    if distributed and scaling == "strong" and not c5 and not args.no_weak:
        asm_w = make_shard(args, synth, DD, world, rank, "weak")
        packed_w = assembly.pack_to_device(asm_w, device=dev, on_device=True)
        ws = max(3, min(args.steps, 10))
        ms_ws, stage_w, _, _ = time_steps(torch, dist, pipeline, L, packed_w, args.norm, ws, 3, distributed, dev)
        ms_w = ms_ws
        if depth > 1:
            ms_w, _, _, _, _ = time_pipelined(torch, dist, pipeline, L, packed_w, args.norm, ws, 3, distributed, dev, depth)
        weak = {"value": world * asm_w.total_bp / (ms_w * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms_w,
                "serial_ms_per_step": ms_ws,
                "per_gpu": f"{asm_w.n_contigs} contigs / {asm_w.total_bp / 1e9:.3f} Gbp",
                "stages_ms": {k2: round(v, 4) for k2, v in stage_w.items()}}
        del packed_w, asm_w

    # ---- end to end from host buffers: pinned 2-bit-packed host codes (headline) and pinned ASCII
    e2e = None
    if not args.no_e2e:
        def e2e_leg(fmt):
            hf = pipeline.HostFeaturiser(asm, device=dev, k=K, method=args.norm, n_components=NCOMP, host_format=fmt)
            for _ in range(2):
                hf.run(distributed=distributed)
            torch.cuda.synchronize()
            if distributed:
                dist.barrier()
            e_steps = max(3, min(args.steps, 10)) * (3 if fmt == "packed" else 1)  # 3 steps in flight: amortise fill/drain
            # (a) serial latency of one call; (b) throughput of back-to-back calls with the next step's H2D
            # copy overlapping the current step's compute (double buffering; every step still copies its
            # input from pinned host memory and reads its scores back inside the timed region)
            t0 = time.perf_counter()
            for _ in range(3):
                hf.run(distributed=distributed)
            torch.cuda.synchronize()
            lat_ms = (time.perf_counter() - t0) / 3 * 1e3
            many = hf.run_stream if fmt == "packed" else hf.run_many
            many(5, distributed=distributed)
            torch.cuda.synchronize()
            if distributed:
                dist.barrier()
            t0 = time.perf_counter()
            many(e_steps, distributed=distributed)
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if distributed:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            e_ms = float(dt.item()) / e_steps * 1e3
            leg = {"value": job_bp / (e_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": e_ms,
                   "serial_latency_ms": lat_ms, "steps": e_steps,
                   "h2d_bytes_per_step": int(hf.h2d_bytes), "d2h_bytes_per_step": int(hf.d2h_bytes),
                   "numa_bound": numa_bound, "host_format": fmt, "path": hf.describe()}
            if hf.host_pack_ms is not None:
                leg["host_pack_ms_outside_timed_region"] = round(hf.host_pack_ms, 2)
            del hf
            return leg

        e2e = e2e_leg("packed")
        e2e["ascii"] = e2e_leg("ascii")

    if rank == 0:
        hbm, bf16, how = measured_peaks()
        acct = stage_accounting(shard_n, shard_bp, Dout=511 if args.norm == "ilr" else 512)
        stage_report = {}
        for s in stages:
            gbs = acct[s]["bytes"] / (stage_ms[s] * 1e-3) / 1e9 if stage_ms[s] > 0 else None
            stage_report[s] = {"ms": round(stage_ms[s], 4), "algorithmic_GB": round(acct[s]["bytes"] / 1e9, 4),
                               "achieved_GBps": None if gbs is None else round(gbs, 1),
                               "frac_of_hbm_peak": None if gbs is None else round(gbs / hbm, 4)}
            if "flops" in acct[s] and stage_ms[s] > 0:
                stage_report[s]["achieved_TFLOPs"] = round(acct[s]["flops"] / (stage_ms[s] * 1e-3) / 1e12, 2)
        # roofline of the dominant DATA-PROPORTIONAL kernel (the stages whose work scales with the
        # assembly); the 512x512 eigensolve is N-independent and latency-bound (510 dependent
        # Householder steps) — it is reported in `stages` and as `n_independent_ms`
        scaling_stages = ["count", "normalise", "gram", "project"]
        dom = max(scaling_stages, key=lambda s: stage_ms[s])
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(dom)
            except Exception:
                traffic = None
        # the tensor-core stage: credited FLOPs = 2 N D^2 over the stage time; the kernel actually
        # executes 24 exact int8 digit-pair products on the 20 upper-triangular 128x64 tiles (CTA-pair kernel:
        # per 32-row K step and tile, 768 accumulator columns x 256 rows x K=32)
        t_g = stage_ms["gram"] * 1e-3
        executed_ops = 2.0 * shard_n * 20 * 768 * 256
        gram_roof = {
            "kernel": "gram (gram_i8x2_kernel: tcgen05.mma.cta_group::2 kind::i8 + fixed-order reduce)", "bound": "tensor",
            "achieved": round(acct["gram"]["flops"] / t_g / 1e12, 2), "peak": bf16, "unit": "TFLOP/s",
            "frac": round(acct["gram"]["flops"] / t_g / 1e12 / bf16, 4),
            "peak_source": f"{how} (MEASURED_PEAKS.json bf16_tflops, burst)",
            "executed_int8_TOPs": round(executed_ops / t_g / 1e12, 1),
            "executed_frac_of_2x_bf16_peak": round(executed_ops / t_g / 1e12 / (2 * bf16), 4),
            "note": "achieved credits 2*N*D^2 fp64-equivalent FLOPs; fp64 accuracy costs 24 int8 digit-pair "
                    "products on the 20 upper-triangular 128x64 tiles (nominal int8 dense peak = 2x bf16), "
                    "executed_* count those",
        }
        if dom == "gram":
            roofline = dict(gram_roof, traffic=traffic)
        else:
            roofline = {
                "kernel": dom, "bound": "hbm",
                "achieved": stage_report[dom]["achieved_GBps"], "peak": hbm, "unit": "GB/s",
                "frac": stage_report[dom]["frac_of_hbm_peak"], "traffic": traffic,
                "peak_source": f"{how} (MEASURED_PEAKS.json hbm_gbs)" if how == "measured" else "fallback 6.65 TB/s",
                "note": "dominant DATA-PROPORTIONAL stage; its measured bound is shared-memory atomic wavefronts "
                        "(count) rather than HBM (DESIGN.md 4).  The longest kernel chain of the step is the "
                        "N-independent eigensolve (see `n_independent_ms`); the tensor-core stage is reported "
                        "under `tensor_kernel`",
                "tensor_kernel": gram_roof,
            }
        comp_bytes = sum(a["bytes"] for a in acct.values())
        roofline["composite"] = {
            "algorithmic_GB": round(comp_bytes / 1e9, 3),
            "achieved_GBps": round(comp_bytes / (ms_per_step * 1e-3) / 1e9, 1),
            "frac": round(comp_bytes / (ms_per_step * 1e-3) / 1e9 / hbm, 4),
            "serial_frac": round(comp_bytes / (ms_serial * 1e-3) / 1e9 / hbm, 4),
        }
        cpu = None
        if not args.no_cpu_baseline and world == 1 and args.config == "c2":  # reported once, on rank 0 at N=1
            cpu = cpu_baseline(asm, os.cpu_count() or 1)
        sweep = None
        if not args.no_sweep and world == 1 and args.config == "c2":  # the 200k-point sweep fits one GPU: replicas only
            sweep = sweep_bench(dev, torch, cpu=not args.no_cpu_baseline)
        line = {
            "metric": METRIC if not c5 else METRIC.replace("am_clr", args.norm), "value": value, "unit": UNIT,
            "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": DTYPE,
            "data": "synthetic",
            "config": {
                "workload": WORKLOADS[args.config],
                "shard": f"{shard_n} contigs / {shard_bp / 1e9:.3f} Gbp on rank 0 of {world}; job {int(job_n)} contigs / "
                         f"{job_bp / 1e9:.3f} Gbp",
                "k": K, "norm": args.norm, "pca_dimensions": NCOMP,
                "parallelism": f"contigs sharded by bp over {world} GPU(s); ONE NCCL all-reduce per step "
                               "([Gram | column sums | row count | column presence | all-NA rows], 2.0 MiB)"
                               + ("; the eigensolve of step i runs on rank i mod N and is broadcast (0.2 MB)"
                                  if (world > 1 and depth > 1) else ""),
                "pipelining": (f"{depth} steps in flight (pipeline.FeatureStream): count/normalise/Gram of step i+1 on one "
                               "stream while the eigensolve of step i and its projection run on two others; all K "
                               "steps complete inside the timed region; `serial` = one step at a time") if depth > 1 else "none",
                "l2": "inputs larger than L2 (0.25 GB codes, 0.41 GB counts, 0.82 GB features per step at N=1); no flush "
                      "in the headline, see `l2_flush_check`",
            },
            "roofline": roofline, "stages": stage_report,
            "n_independent_ms": round(stage_ms["eigh"], 4),
            "n_independent_note": "top-50 eigenpairs of the 512x512 covariance (tridiagonalisation + bisection + inverse "
                                  "iteration + back-transformation): the same time whatever the assembly size, replicated "
                                  "on every rank; the longest stage of the step, not a bandwidth-bound one",
            "l2_flush_check": flush_check, "serial": serial, "host": host_acct, "weak": weak,
            "cpu_baseline": cpu, "e2e": e2e, "sweep": sweep,
            "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if distributed:
        dist.destroy_process_group()


def _rec(torch):
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


if __name__ == "__main__":
    main()
