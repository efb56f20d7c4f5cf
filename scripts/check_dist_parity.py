"""N-rank parity of the sharded featurise path (run under torchrun, one rank per GPU).

Every rank featurises its bp-balanced contig shard with distributed=True; rank 0 also featurises the WHOLE
assembly alone and compares explained variance, components and the gathered scores.  Cases:
  * am_clr / clr / ilr on a plain assembly (one all-reduce per pass);
  * `tiny`: the last rank's shard ends with contigs of length 4, 5, 6 (all-NA rows, kmers.py:187-192) -> every
    rank takes the re-run with row gathers (am_clr) or raises skbio's closure error (clr);
  * `empty`: rank 0 holds every contig, the other ranks none;
  * `stream`: pipeline.FeatureStream (several passes in flight, the eigensolve of pass i on rank i mod N +
    broadcast) against the serial single-GPU pass, for four different assemblies.
Prints one line per case and DIST PARITY PASS / FAIL."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from autometa_b200 import assembly as A, dist as D, pipeline, synth

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
ok = True


def compare(tag, res, scores, ref):
    global ok
    ev, ev0 = res["explained_variance"].cpu().numpy(), ref["explained_variance"].cpu().numpy()
    V, V0 = res["components"].cpu().numpy(), ref["components"].cpu().numpy()
    s, s0 = scores.cpu().numpy(), ref["scores"].cpu().numpy()
    e1 = np.max(np.abs(ev - ev0) / ev0)
    e2 = np.max(1.0 - np.abs(np.sum(V * V0, axis=1)))
    e3 = np.max(np.abs(s - s0)) / np.max(np.abs(s0)) if s.shape == s0.shape else np.inf
    good = s.shape == s0.shape and e1 < 1e-9 and e2 < 1e-9 and e3 < 1e-8
    ok &= bool(good)
    print(f"{tag}: ranks={world} rows={s.shape[0]} |dEV|/EV={e1:.2e} 1-|cos|={e2:.2e} |dscore|/max={e3:.2e} "
          f"{'OK' if good else 'MISMATCH'}", flush=True)


def shard_of(full, case):
    if case == "empty":
        return full.slice(0, full.n_contigs) if rank == 0 else full.slice(full.n_contigs, full.n_contigs)
    lo, hi = D.shard_ranges(full.lengths, world)[rank]
    return full.slice(lo, hi)


full = synth.make_assembly(20000, 60_000_000, seed=77)
full_tiny = synth.make_assembly(12000, 36_000_000, seed=78, tiny=True)
for case, asm, method in (("plain", full, "am_clr"), ("plain", full, "clr"), ("plain", full, "ilr"),
                          ("tiny", full_tiny, "am_clr"), ("tiny", full_tiny, "clr"), ("empty", full, "am_clr")):
    packed = A.pack_to_device(shard_of(asm, case), device="cuda", on_device=True)
    try:
        res = pipeline.featurise(packed, 5, method, 50, distributed=True)
        err = None
    except ValueError as exc:
        res, err = None, str(exc)
    errs = [None] * world
    dist.all_gather_object(errs, err)
    if case == "tiny" and method == "clr":
        good = all(e and "rows with all zeros" in e for e in errs)
        ok &= good
        if rank == 0:
            print(f"{case}/{method}: every rank raised the closure error: {'OK' if good else 'MISMATCH ' + repr(errs)}", flush=True)
        continue
    assert err is None, err
    scores = D.allgather_rows(res["scores"])
    if rank == 0:
        ref = pipeline.featurise(A.pack_to_device(asm, device="cuda", on_device=True), 5, method, 50)
        compare(f"{case}/{method}", res, scores, ref)
    dist.barrier()

# throughput mode with the eigensolves spread over the ranks
asms = [synth.make_assembly(9000 + 500 * i, (9000 + 500 * i) * 3300, seed=90 + i) for i in range(4)]
fs = pipeline.FeatureStream(torch.device("cuda", torch.cuda.current_device()), 5, "am_clr", 50, distributed=True,
                            depth=world + 1)
for a in asms:
    fs.submit(A.pack_to_device(shard_of(a, "plain"), device="cuda", on_device=True))
outs = fs.drain()
for i, (a, res) in enumerate(zip(asms, outs)):
    scores = D.allgather_rows(res["scores"])
    if rank == 0:
        ref = pipeline.featurise(A.pack_to_device(a, device="cuda", on_device=True), 5, "am_clr", 50)
        compare(f"stream[{i}] (eigensolve on rank {i % world})", res, scores, ref)
    dist.barrier()
dist.barrier()
if rank == 0:
    print("DIST PARITY", "PASS" if ok else "FAIL", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
