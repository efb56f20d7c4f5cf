import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from autometa_b200 import device as D
rng = np.random.default_rng(0)
n, d = 200000, 512
X = torch.from_numpy(rng.normal(size=(n, d))).cuda()
cs, cm = D.colstats(X); mu = cs / n
e, sc = D.plane_scales(mu, cm)
planes = D.quantize_planes(X, mu, sc)
fn = lambda: D.gram_planes(planes, n, d, e)
for _ in range(2): fn()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): fn()
e1.record(); torch.cuda.synchronize()
print("mode", os.environ.get("AM_GRAM_DBG", "0"), "gram_planes: %.3f ms" % (e0.elapsed_time(e1) / 5))
