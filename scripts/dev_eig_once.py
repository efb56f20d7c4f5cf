"""One eigh_topk on a 512x512 covariance + the phase cycle counters of the tridiagonalisation (CTA 0)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
os.environ.setdefault("AM_TRIDIAG_PROBE", "1")  # cycle counters on (set to "" for production timing)
if not os.environ["AM_TRIDIAG_PROBE"]: del os.environ["AM_TRIDIAG_PROBE"]
from autometa_b200 import device as D
rng = np.random.default_rng(0)
n, P = 512, 50
A_ = rng.normal(size=(n, 5000)); S = A_ @ A_.T / 5000
A = torch.from_numpy(S).cuda()
for _ in range(3): D.eigh_topk(A, P)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): D.eigh_topk(A, P)
e1.record(); torch.cuda.synchronize()
print("eigh_topk %.3f ms" % (e0.elapsed_time(e1) / 10))
al = lambda b: (b + 255) & ~255
npad = (n + 31) & ~31
off = al(n * npad * 8) + 3 * al(npad * 8) + al(128 * 8) + al(P * npad * 8)
w = [v for k, v in D._work_cache.items() if k[0] == "eigtop"][0]
dbg = w[off:off + 256].view(torch.int64).cpu().numpy()
tot = dbg[:4].sum() + dbg[4] + dbg[5]
names = ["S scalars", "B update+matvec+send", "wait inbox", "C w/next column"]
for nm, v in zip(names, dbg[:4]):
    print("%-24s %9d clk  %5.1f %%  %6.0f clk/step" % (nm, v, 100.0 * v / tot, v / (n - 2)))
print("  of B: compute %d clk/step, reduce %d clk/step, send = rest" % (dbg[4] / (n - 2), dbg[5] / (n - 2)))
print("B compute at k = 0, 64, ..., 448:", dbg[8:16].tolist())
print("S scalars  at k = 0, 64, ...:", dbg[24:32].tolist())
print("inbox wait at k = 0, 64, ...:", dbg[16:24].tolist())
print("total %d clk = %.3f ms at 1.965 GHz" % (tot, tot / 1.965e6))
