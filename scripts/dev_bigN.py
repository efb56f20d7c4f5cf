"""C5-shaped sanity run on one GPU: 1M contigs (short ones, ~1.1 Gbp), am_clr fused path and ilr path."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from autometa_b200 import assembly, pipeline, synth
t0 = time.time()
asm = synth.make_assembly(1_000_000, 3_300_000_000, seed=7)
print("synth %.1fs, %d contigs, %.2f Gbp" % (time.time() - t0, asm.n_contigs, asm.total_bp / 1e9))
packed = assembly.pack_to_device(asm, device="cuda:0", on_device=True)
for method in ("am_clr", "ilr"):
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        res = pipeline.featurise(packed, 5, method, 50, keep=True)
        torch.cuda.synchronize(); dt = time.time() - t0
    X = res["normalized"]; sc = res["scores"]; ev = res["explained_variance"]
    # cheap self-consistency: scores variance == explained variance; columns of scores uncorrelated
    var = sc.var(dim=0, unbiased=True)
    print(method, "%.1f ms" % (dt * 1e3), tuple(X.shape), tuple(sc.shape),
          "max|var(scores)/ev - 1| = %.2e" % float(((var / ev) - 1).abs().max()),
          "rowsum(counts)==L-k:", bool((res["row_sum"].cpu().numpy() <= np.maximum(asm.lengths - 5, 0)).all()))
    del res, X, sc
    torch.cuda.empty_cache()
