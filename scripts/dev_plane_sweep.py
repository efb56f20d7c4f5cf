"""How many int8 digit planes does the Gram need?  

C2 workload (200k contigs / 1 Gbp, am_clr).  For S = 3..7 digit planes the centred Gram is formed on the
tensor cores (am_gram_i8: the generic 1-CTA kernel takes any S), the covariance goes through the production
eigensolver, and the scores are computed in fp64 (am_project) so that only the Gram's precision varies.
Compared with sklearn PCA(50) on the same fp64 matrix (the parity oracle): the three gates are
|dEV|/EV <= 1e-6, 1 - |cos| <= 1e-9, |dscore| <= 1e-8 max|score|.
Two fixed-point ranges: per-column max|x| (a second pass over the data) and the data-independent hard bound
log1p(longest contig) the fused production pass uses (no extra pass)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sklearn.decomposition import PCA
from autometa_b200 import assembly, device as D, synth, pipeline

n, bp = (200_000, 1_000_000_000) if len(sys.argv) < 2 else (int(sys.argv[1]), int(sys.argv[2]))
asm = synth.make_assembly(n, bp, seed=20261019 + 2)
packed = assembly.pack_to_device(asm, device="cuda:0", on_device=True)
res = pipeline.featurise(packed, 5, "am_clr", 50)
X = res["normalized"]
Xh = X.cpu().numpy()
p = PCA(n_components=50, random_state=np.random.RandomState(42))
sc_ref = p.fit_transform(Xh)
ev_ref, V_ref = p.explained_variance_, p.components_
smax = np.abs(sc_ref).max()


def errs(ev, V, sc):
    return dict(ev=float(np.max(np.abs(ev - ev_ref) / ev_ref)),
                one_minus_cos=float(np.max(1.0 - np.abs(np.sum(V * V_ref, axis=1)))),
                score=float(np.max(np.abs(sc - sc_ref)) / smax))


out = {"production_S6_fused": errs(res["explained_variance"].cpu().numpy(), res["components"].cpu().numpy(),
                                   res["scores"].cpu().numpy())}
nrows = X.shape[0]
cs, cm = D.colstats(X)
mu = cs / nrows
bound = float(np.log1p(float(packed.lengths.max())))
for rng_name, cmax in (("per_column_maxabs", cm), ("hard_bound_log1p_maxlen", torch.full_like(cm, bound))):
    for S in (3, 4, 5, 6, 7):
        G = D.gram_centered_i8(X, mu, cmax, slices=S)
        C = G / (nrows - 1.0)
        ev, V, tr = D.eigh_topk(C, 50)
        sc = D.project(X, mu, V)
        out[f"{rng_name}_S{S}"] = errs(ev.cpu().numpy(), V.cpu().numpy(), sc.cpu().numpy())
# fp64 CUDA-core Gram as the floor
G = D.gram_centered(X, mu)
ev, V, tr = D.eigh_topk(G / (nrows - 1.0), 50)
out["fp64_cuda_core_gram"] = errs(ev.cpu().numpy(), V.cpu().numpy(), D.project(X, mu, V).cpu().numpy())
gaps = np.abs(np.diff(ev_ref)) / ev_ref[:-1]
out["spectrum"] = dict(ev1=float(ev_ref[0]), ev50=float(ev_ref[-1]), min_rel_gap=float(gaps.min()),
                       median_rel_gap=float(np.median(gaps)))
for k, v in out.items():
    print(k, json.dumps(v))
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out",
                                 "plane_sweep.json"), "w"), indent=1)
