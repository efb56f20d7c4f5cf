"""Time am_gram_planes alone at C2 (random int8 digit planes), AM_GRAM_STAGES / AM_GRAM_PAIR from the env."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from autometa_b200 import device as D, _lib as L
n, d, S = 200000, 512, 6
dev = torch.device("cuda:0")
nbytes = L.lib.am_planes_bytes(n, d, S)
planes = torch.randint(-128, 128, (nbytes,), dtype=torch.int8, device=dev)
e = torch.zeros(d, dtype=torch.int32, device=dev)
G = torch.zeros((d, d), dtype=torch.float64, device=dev)
for _ in range(3): D.gram_planes(planes, n, d, e, out=G)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): D.gram_planes(planes, n, d, e, out=G)
e1.record(); torch.cuda.synchronize()
G.zero_(); D.gram_planes(planes, n, d, e, out=G); torch.cuda.synchronize()
print("stages=%s pair=%s: %.3f ms  checksum %.6e" % (os.environ.get("AM_GRAM_STAGES", "8"), os.environ.get("AM_GRAM_PAIR", "1"),
      e0.elapsed_time(e1) / 20, float(G.sum().item())))
