import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from autometa_b200 import device as D, synth
from oracle import oracle_np as O
table, _ = synth.make_sweep_input()
X = np.ascontiguousarray(table[["x_1", "x_2", "coverage"]].to_numpy())
sw = D.EpsSweep(torch.from_numpy(X).cuda())
for eps in O.eps_schedule(48):
    sw.step(eps)
torch.cuda.synchronize()
print("ok", int(sw.n_clusters.item()))
