"""Count-kernel time for 1 Gbp cut into different numbers of contigs (per-item overheads vs per-window work)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from autometa_b200 import assembly, device as D, synth
for n in (20000, 200000, 320000):
    asm = synth.make_assembly(n, 1000000000, seed=3)
    packed = assembly.pack_to_device(asm, device="cuda:0", on_device=True)
    for _ in range(3): D.count_packed(packed, 5)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): c = D.count_packed(packed, 5)
    e1.record(); torch.cuda.synchronize()
    print("%8d contigs (mean %6.0f bp): count_packed %.3f ms" % (n, 1e9 / n, e0.elapsed_time(e1) / 10), flush=True)
    del packed, c
