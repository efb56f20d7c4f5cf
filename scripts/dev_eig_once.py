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
