"""dev check: am_eigh_topk vs numpy.linalg.eigh + timing (GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from autometa_b200 import device as D

rng = np.random.default_rng(0)
def check(S, P, name):
    A = torch.from_numpy(S).cuda()
    w, V, tr = D.eigh_topk(A, P)
    torch.cuda.synchronize()
    w = w.cpu().numpy(); V = V.cpu().numpy(); tr = float(tr.cpu())
    wr, Vr = np.linalg.eigh(S)
    wr = wr[::-1]; Vr = Vr[:, ::-1].T
    n = S.shape[0]
    ev_err = np.max(np.abs(w - wr[:P])) / abs(wr[0])
    orth = np.max(np.abs(V @ V.T - np.eye(P)))
    resid = np.max(np.abs(S @ V.T - V.T * w)) / abs(wr[0])
    cos = np.abs(np.sum(V * Vr[:P], axis=1))
    print(f"{name:28s} n={n:4d} P={P:3d} ev_err={ev_err:.2e} orth={orth:.2e} resid={resid:.2e} min|cos|={cos.min():.12f} trace_err={abs(tr-np.trace(S))/abs(np.trace(S)):.1e}")

for n, P in [(3, 2), (5, 5), (17, 4), (64, 50), (136, 50), (511, 50), (512, 50), (512, 128), (576, 50)]:
    A_ = rng.normal(size=(n, n + 3)); S = A_ @ A_.T / n
    check(S, P, "wishart")
# decaying spectrum like k-mer covariance
n = 512
Q, _ = np.linalg.qr(rng.normal(size=(n, n)))
lam = 1.0 / (1 + np.arange(n)) ** 1.5; lam[-1] = 0.0
S = (Q * lam) @ Q.T; S = 0.5 * (S + S.T)
check(S, 50, "power-law spectrum")
lam2 = lam.copy(); lam2[3] = lam2[4] = lam2[5]  # triple eigenvalue
S = (Q * lam2) @ Q.T; S = 0.5 * (S + S.T)
A = torch.from_numpy(S).cuda(); w, V, tr = D.eigh_topk(A, 50); w = w.cpu().numpy(); V = V.cpu().numpy()
print("degenerate: ev_err", np.max(np.abs(w - np.sort(lam2)[::-1][:50])), "orth", np.max(np.abs(V @ V.T - np.eye(50))), "resid", np.max(np.abs(S @ V.T - V.T * w)))
S = np.diag(np.arange(1, 513, dtype=float)); check(S, 50, "diagonal")
# timing
A_ = rng.normal(size=(512, 5000)); S = A_ @ A_.T / 5000
A = torch.from_numpy(S).cuda()
for _ in range(3): D.eigh_topk(A, 50)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): D.eigh_topk(A, 50)
e1.record(); torch.cuda.synchronize()
print("eigh_topk 512/50: %.3f ms" % (e0.elapsed_time(e1) / 20))
