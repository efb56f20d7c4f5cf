"""The eps sweeps of large-data mode: many small (<= 10 000 contig) get_clusters problems — the taxa of one rank.
(a) one sweep at a time with the schedule on the host (one read-back per eps: AUTOMETA_B200_SWEEP=host),
(b) one at a time with the loop on the device (controller kernel + CUDA-graph replays, csrc/am_sweep.cu),
(c) all of them in flight together (get_clusters_many).  Also the C4-size sweep API (200k points) both ways.
Synthetic: `parts` tables of `per` contigs with 6 clusters each, 40 markers."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd, torch
from autometa_b200 import _lib as L, synth
from autometa_b200.binning import recursive_dbscan as R

parts, per = int(os.environ.get("PARTS", 64)), int(os.environ.get("PER", 2000))
tables, markers = [], None
for i in range(parts):
    t, m = synth.make_sweep_input(n_points=per, n_clusters=6, n_markers=40, seed=100 + i)
    t.index = pd.Index([f"T{i}_{c}" for c in t.index], name="contig")
    m.index = pd.Index([f"T{i}_{c}" for c in m.index], name="contig")
    tables.append(t)
    markers = m if markers is None else pd.concat([markers, m])
cut = (20.0, 90.0, 25.0, 5.0)
out = {"partitions": parts, "contigs_per_partition": per, "markers": 40}


def run(mode, many):
    os.environ["AUTOMETA_B200_SWEEP"] = mode
    torch.cuda.synchronize()
    l0, t0 = L.launch_count(), time.perf_counter()
    if many:
        res = R.get_clusters_many([t.copy() for t in tables], markers, *cut, method="dbscan")
    else:
        res = [R.get_clusters(t.copy(), markers, *cut, method="dbscan") for t in tables]
    torch.cuda.synchronize()
    return res, time.perf_counter() - t0, L.launch_count() - l0


run("device", True)  # warm-up (allocator, streams)
a, ta, la = run("host", False)
b, tb, lb = run("device", False)
c, tc, lc = run("device", True)
for x, y in zip(a, c):
    assert x["cluster"].fillna("").tolist() == y["cluster"].fillna("").tolist()
out.update(host_loop_one_at_a_time_s=round(ta, 4), device_loop_one_at_a_time_s=round(tb, 4),
           device_loop_in_flight_together_s=round(tc, 4), kernels_host_loop=la, kernels_device_loop=lc,
           bins=int(sum(x["cluster"].nunique() for x in c)))
# host share: the frames (pandas) of the same calls with the sweeps stubbed out is not separable here; report
# the per-partition time instead
out["ms_per_partition"] = {"host_loop": round(ta / parts * 1e3, 3), "device_loop": round(tb / parts * 1e3, 3),
                           "in_flight_together": round(tc / parts * 1e3, 3)}
if os.environ.get("C4", "1") == "1":
    table, mk = synth.make_sweep_input()
    for mode in ("host", "device"):
        os.environ["AUTOMETA_B200_SWEEP"] = mode
        R.recursive_dbscan(table.copy(), mk, 20.0, 95.0, 25.0, 5.0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cl, un = R.recursive_dbscan(table.copy(), mk, 20.0, 95.0, 25.0, 5.0)
        torch.cuda.synchronize()
        out[f"c4_api_{mode}_loop_s"] = round(time.perf_counter() - t0, 4)
        out[f"c4_clustered_{mode}"] = int(cl.shape[0])
print(json.dumps(out))
