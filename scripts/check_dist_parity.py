"""N-rank parity of the sharded featurise path (run under torchrun, one rank per GPU):
every rank featurises its bp-balanced contig shard with distributed=True; rank 0 also featurises the
WHOLE assembly alone and compares explained variance, components and the gathered scores."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from autometa_b200 import assembly as A, dist as D, pipeline, synth

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
full = synth.make_assembly(20000, 60_000_000, seed=77)
lo, hi = D.shard_ranges(full.lengths, world)[rank]
shard = full.slice(lo, hi)
ok = True
for method in ("am_clr", "clr", "ilr"):
    packed = A.pack_to_device(shard, device="cuda", on_device=True)
    res = pipeline.featurise(packed, 5, method, 50, distributed=True)
    scores = D.allgather_rows(res["scores"])
    if rank == 0:
        ref = pipeline.featurise(A.pack_to_device(full, device="cuda", on_device=True), 5, method, 50)
        ev, ev0 = res["explained_variance"].cpu().numpy(), ref["explained_variance"].cpu().numpy()
        V, V0 = res["components"].cpu().numpy(), ref["components"].cpu().numpy()
        s, s0 = scores.cpu().numpy(), ref["scores"].cpu().numpy()
        e1 = np.max(np.abs(ev - ev0) / ev0)
        e2 = np.max(1.0 - np.abs(np.sum(V * V0, axis=1)))
        e3 = np.max(np.abs(s - s0)) / np.max(np.abs(s0))
        good = s.shape == s0.shape and e1 < 1e-9 and e2 < 1e-9 and e3 < 1e-8
        ok &= bool(good)
        print(f"{method}: ranks={world} |dEV|/EV={e1:.2e} 1-|cos|={e2:.2e} |dscore|/max={e3:.2e} {'OK' if good else 'MISMATCH'}", flush=True)
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("DIST PARITY", "PASS" if ok else "FAIL", flush=True)
    sys.exit(0 if ok else 1)
