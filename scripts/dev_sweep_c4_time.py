"""The 48-eps C4 sweep (200k points) timed with CUDA events: eps step + labels per eps (what bench.py's `sweep`
reports), for the current AM_EPS_WARP_MAX (points up to which the warp-per-point neighbour kernel is used)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from autometa_b200 import device as D, synth
from oracle import oracle_np as O
table, _ = synth.make_sweep_input()
X = np.ascontiguousarray(table[["x_1", "x_2", "coverage"]].to_numpy())
Xd = torch.from_numpy(X).cuda()
grid = O.eps_schedule(48)
res = {}
for rep in range(3):
    sw = D.EpsSweep(Xd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for eps in grid:
        sw.step(eps)
    e1.record()
    e1.synchronize()
    res[f"rep{rep}_ms"] = round(e0.elapsed_time(e1), 3)
res["clusters_at_last_eps"] = int(sw.n_clusters.item())
res["AM_EPS_WARP_MAX"] = os.environ.get("AM_EPS_WARP_MAX", "default")
print(json.dumps(res))
