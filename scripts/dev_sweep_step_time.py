"""GPU time per eps step of the device-resident sweep (CUDA events around batches of steps), graph replays
against plain launches, for a few table sizes; `NCU=1` runs one short plain-launch sweep only (for an ncu
launch list of the step's kernels)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from autometa_b200 import device as D, synth
from autometa_b200.binning import recursive_dbscan as R
from autometa_b200.binning.utilities import markers_coo

dev = torch.device("cuda", 0)
out = {}
sizes = (int(os.environ.get("N", 2000)),) if os.environ.get("NCU") else (500, 2000, 10000, 200000)
for n in sizes:
    table, markers = synth.make_sweep_input(n_points=n, n_clusters=max(3, n // 400), n_markers=40, seed=7)
    X = R._feature_matrix(table)
    cov, gc = table["coverage"].to_numpy(), table["gc_content"].to_numpy()
    rows, cols, vals = markers_coo(markers, table.index)
    for graph in ((False,) if os.environ.get("NCU") else (True, False)):
        run = D.DeviceSweep(X, cov, gc, rows, cols, vals, markers.shape[1], (20.0, 90.0, 25.0, 5.0), dev, graph=graph)
        run.launch(8)
        run.wait()
        if os.environ.get("NCU"):
            break
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(run.stream)
        run.launch(32)
        e1.record(run.stream)
        t_host = time.perf_counter() - t0
        e1.synchronize()
        out[f"n{n}_{'graph' if graph else 'plain'}"] = {"gpu_us_per_step": round(e0.elapsed_time(e1) / 32 * 1e3, 1),
                                                        "host_us_per_step": round(t_host / 32 * 1e6, 1)}
        run.close()
print(json.dumps(out))
