"""Probe: does the count kernel (shared-memory-atomic bound, 16 % of HBM) share the chip with the normalise pass
(HBM bound) when both are in flight on two streams?  Times count alone, normalise alone, one after the other and
side by side (count of assembly B with the normalise of assembly A's counts), for the current AM_COUNT_OCC."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from autometa_b200 import assembly, device as D, pipeline, synth
dev = torch.device("cuda", 0)
asm = synth.make_assembly(200_000, 1_000_000_000, seed=20261019 + 2)
packed = assembly.pack_to_device(asm, device="cuda:0", on_device=True)
counts, present, row_sum = D.count_packed(packed, 5)
n = counts.shape[0]
sample = pipeline._sample_rows(n, dev, packed._plans)
cs = torch.zeros(512, dtype=torch.float64, device=dev)
D.normalize_counts(counts, "am_clr", sample, None, colsum=cs)
center = cs / sample.numel()
e, sc = D.plane_scales(center, None, pipeline._bound_for(packed, "am_clr", 512))
X = torch.empty((n, 512), dtype=torch.float64, device=dev)
planes = D.alloc_planes(n, 512, dev)
colsum = torch.zeros(512, dtype=torch.float64, device=dev)
counts2 = torch.empty_like(counts); present2 = torch.zeros_like(present); row_sum2 = torch.empty(n, dtype=torch.int32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def count(st):
    with torch.cuda.stream(st):
        D.count_packed(packed, 5, counts=counts2, col_present=present2, row_sum=row_sum2)
def norm(st):
    with torch.cuda.stream(st):
        D.normalize_planes(counts, "am_clr", None, center, sc, colsum, out=X, planes=planes)
def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
import inspect
out = {"AM_COUNT_OCC": os.environ.get("AM_COUNT_OCC", "default")}
out["count_ms"] = round(timed(lambda: count(s1)), 4)
out["normalise_ms"] = round(timed(lambda: norm(s1)), 4)
out["one_after_the_other_ms"] = round(timed(lambda: (count(s1), norm(s1))), 4)
out["side_by_side_ms"] = round(timed(lambda: (count(s1), norm(s2))), 4)
out["side_by_side_norm_first_ms"] = round(timed(lambda: (norm(s2), count(s1))), 4)
print(json.dumps(out))
