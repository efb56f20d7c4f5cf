"""Per-kernel durations of eigh_topk in-stream (warm caches) via torch.profiler / CUPTI."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from autometa_b200 import device as D
rng = np.random.default_rng(0)
n, P = 512, 50
A_ = rng.normal(size=(n, 5000)); S = A_ @ A_.T / 5000
A = torch.from_numpy(S).cuda()
for _ in range(3): D.eigh_topk(A, P)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(10): D.eigh_topk(A, P)
    torch.cuda.synchronize()
rows = [(e.key[:60], e.device_time_total / e.count, e.count) for e in prof.key_averages() if e.device_time_total > 0]
for k, t, c in sorted(rows, key=lambda r: -r[1]):
    print("%-62s %8.1f us x %d" % (k, t, c))
