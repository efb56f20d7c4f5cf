"""dev check: am_gram_i8 (tcgen05 int8 slices) vs fp64 Gram + timing."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from autometa_b200 import device as D

rng = np.random.default_rng(0)
def run(n, d, S=6, timing=False):
    X = rng.normal(size=(n, d)) * rng.uniform(0.2, 3.0, size=d) + rng.normal(size=d)
    Xd = torch.from_numpy(X).cuda()
    cs, cm = D.colstats(Xd)
    mu = cs / n
    G8 = D.gram_centered_i8(Xd, mu, cm, slices=S)
    G64 = D.gram_centered(Xd, mu)
    torch.cuda.synchronize()
    Xc = (X - mu.cpu().numpy()).astype(np.longdouble)
    ref = (Xc.T @ Xc).astype(np.float64) if n * d * d < 3e9 else None
    a = G8.cpu().numpy(); b = G64.cpu().numpy()
    sc = np.sqrt(np.outer(np.diag(b), np.diag(b)))
    print(f"n={n} d={d} S={S}: max|G8-G64|/sqrt(GiiGjj) = {np.max(np.abs(a-b)/sc):.3e}", end="")
    if ref is not None:
        print(f"  |G8-ref| = {np.max(np.abs(a-ref)/sc):.3e}  |G64-ref| = {np.max(np.abs(b-ref)/sc):.3e}", end="")
    print("  sym:", float(np.max(np.abs(a - a.T))))
    if timing:
        for fn, name in ((lambda: D.gram_centered_i8(Xd, mu, cm, slices=S), "i8"), (lambda: D.gram_centered(Xd, mu), "f64")):
            for _ in range(2): fn()
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): fn()
            e1.record(); torch.cuda.synchronize()
            print(f"   {name}: {e0.elapsed_time(e1)/5:.3f} ms")

run(600, 64, 6)
run(5000, 512, 6)
run(5000, 512, 4)
run(5000, 500, 7)
run(6007, 136, 6)
run(200000, 512, 6, timing=True)
