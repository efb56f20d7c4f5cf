"""Host-side cost of queueing one pass (no GPU wait): FeatureStream.submit on a small assembly, where the GPU work
is shorter than the launch sequence, and cProfile of where the time goes."""
import os, sys, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autometa_b200 import assembly, pipeline, synth
asm = synth.make_assembly(25_000, 125_000_000, seed=5)
packed = assembly.pack_to_device(asm, device="cuda:0", on_device=True)
fs = pipeline.FeatureStream(torch.device("cuda", 0), 5, "am_clr", 50, depth=8)
for _ in range(10): fs.submit(packed)
fs.drain(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(40): fs.submit(packed)
t1 = time.perf_counter()
fs.drain(); torch.cuda.synchronize()
t2 = time.perf_counter()
print("host us per submit (queue only): %.1f   incl. drain: %.1f us per pass" % ((t1 - t0) / 40 * 1e6, (t2 - t0) / 40 * 1e6))
pr = cProfile.Profile(); pr.enable()
for _ in range(40): fs.submit(packed)
pr.disable(); fs.drain()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
