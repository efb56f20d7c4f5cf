"""dev check: tensor-core projection vs fp64 projection."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from autometa_b200 import device as D

rng = np.random.default_rng(1)
for n, d, P in [(300, 512, 50), (5000, 512, 50), (6007, 136, 20), (200000, 512, 50)]:
    X = rng.normal(size=(n, d)) * rng.uniform(0.2, 3.0, size=d) + rng.normal(size=d)
    Xd = torch.from_numpy(X).cuda()
    cs, cm = D.colstats(Xd)
    mu = cs / n
    center = mu + 0.01 * torch.randn_like(mu)
    e, sc = D.plane_scales(center, cm)
    planes = D.quantize_planes(Xd, center, sc)
    Q, _ = np.linalg.qr(rng.normal(size=(d, P)))
    comps = torch.from_numpy(np.ascontiguousarray(Q.T)).cuda()
    s8 = D.project_planes(planes, n, d, e, center, mu, comps)
    s64 = D.project(Xd, mu, comps)
    torch.cuda.synchronize()
    a = s8.cpu().numpy(); b = s64.cpu().numpy()
    print(f"n={n} d={d} P={P}: max|s8-s64|/max|s| = {np.max(np.abs(a-b))/np.max(np.abs(b)):.3e}")
    if n == 200000:
        for fn, name in ((lambda: D.project_planes(planes, n, d, e, center, mu, comps), "i8"), (lambda: D.project(Xd, mu, comps), "f64")):
            for _ in range(2): fn()
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): fn()
            e1.record(); torch.cuda.synchronize()
            print(f"   {name}: {e0.elapsed_time(e1)/5:.3f} ms")
