import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from autometa_b200 import device as D
rng = np.random.default_rng(0)
A_ = rng.normal(size=(512, 5000)); S = A_ @ A_.T / 5000
A = torch.from_numpy(S).cuda()
for _ in range(3): D.eigh_topk(A, 50)
torch.cuda.synchronize()
print("ok")
# phase counters live at the end of the work buffer
import ctypes
from autometa_b200 import _lib as L
nb = L.lib.am_eigh_topk_work_bytes(512, 50)
w = D._work_cache[("eigtop", 0, 0)]
dbg = w[nb - 256: nb - 256 + 32].view(torch.int64).cpu().numpy()
print("phase cycles/step A,B,S,C:", dbg / 510.0)
