import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from autometa_b200 import device as D, synth
from autometa_b200.binning.utilities import markers_coo
table, markers = synth.make_sweep_input()
X = np.ascontiguousarray(table[["x_1", "x_2", "coverage"]].to_numpy())
cov = table["coverage"].to_numpy(np.float64); gc = table["gc_content"].to_numpy(np.float64)
rows, cols, vals = markers_coo(markers, table.index)
sw = D.EpsSweep(torch.from_numpy(X).cuda())
met = D.ClusterMetrics(len(table), rows, cols, vals, markers.shape[1], cov, gc, (20.0, 95.0, 25.0, 5.0), sw.dev)
for rep in range(2):
    sw.reset()
    eps, step = 0.3, 0.1; rounds = {0: 0}; ncl = 1 << 30; it = 0; log = []
    torch.cuda.synchronize(); t0 = time.perf_counter()
    while ncl > 1:
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
        e0.record(); lab = sw.step(eps); e1.record(); s = met.compute(lab, sw.n_clusters); e2.record()
        s = s.cpu().numpy(); ncl = int(s[0]); med = float(s[1])
        log.append((eps, ncl, e0.elapsed_time(e1), e1.elapsed_time(e2)))
        rounds[ncl] = rounds.get(ncl, 0) + 1
        if rounds[ncl] > 10: step *= 10
        if med == 0: rounds[0] += 1
        if rounds[0] >= 10: break
        eps += step; it += 1
    dt = time.perf_counter() - t0
print("iterations", len(log), "wall %.3f s" % dt, "sum dbscan ms %.1f  sum metrics ms %.1f" % (sum(l[2] for l in log), sum(l[3] for l in log)))
for i in list(range(0, len(log), 12)) + [len(log) - 1]:
    print(i, "eps=%.4g clusters=%d dbscan=%.2f ms metrics=%.2f ms" % log[i])
