import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from autometa_b200 import assembly, device as D, synth, pipeline
asm = synth.make_assembly(200000, 1000000000, seed=20261019 + 2)
packed = assembly.pack_to_device(asm, device="cuda:0", on_device=True)
dev = packed.device
for _ in range(3): pipeline.featurise(packed, 5, "am_clr", 50, keep=False)
torch.cuda.synchronize()
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
for rep in range(2):
    marks = []
    marks.append(("start", ev()))
    counts, present, row_sum = D.count_packed(packed, 5); marks.append(("count", ev()))
    flags = torch.stack([present.min(), row_sum.min()]); marks.append(("flags", ev()))
    n_out = packed.n_contigs; Dm = 512
    cs = torch.zeros(Dm, dtype=torch.float64, device=dev)
    sample = pipeline._sample_rows(n_out, dev, packed._plans)
    cs_s = torch.zeros(Dm + 1, dtype=torch.float64, device=dev); cs_s[Dm] = float(sample.numel()); marks.append(("zeros", ev()))
    D.normalize_counts(counts, "am_clr", sample, None, colsum=cs_s[:Dm]); marks.append(("sample_norm", ev()))
    center = cs_s[:Dm] / cs_s[Dm]
    e, sc = D.plane_scales(center, None, float(np.log1p(float(packed.lengths.max())))); marks.append(("scales", ev()))
    X, planes = D.normalize_planes(counts, "am_clr", None, center, sc, cs); marks.append(("normalize_planes", ev()))
    n_t = torch.tensor([float(X.shape[0])], dtype=torch.float64, device=dev); mu = cs / n_t; marks.append(("mu", ev()))
    G = D.gram_planes(planes, X.shape[0], Dm, e); marks.append(("gram_planes", ev()))
    dlt = mu - center; G = G - n_t * torch.outer(dlt, dlt); C = G / (n_t - 1.0); marks.append(("finishC", ev()))
    evals, comps, total = D.eigh_topk(C, 50); marks.append(("eigh", ev()))
    evals = torch.clamp(evals, min=0.0)
    scores = D.project_planes(planes, X.shape[0], Dm, e, center, mu, comps); marks.append(("project", ev()))
    torch.cuda.synchronize()
    if rep:
        for (a, ea), (b, eb) in zip(marks[:-1], marks[1:]):
            print(f"{b:18s} {ea.elapsed_time(eb)*1e3:8.1f} us")
        print("total %.1f us" % (marks[0][1].elapsed_time(marks[-1][1]) * 1e3))
    del X, planes, G, C, scores
