"""Cost of the GC% by-product inside the device packer at C2 (200k contigs / 1 Gbp)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from autometa_b200 import assembly as A, synth
asm = synth.make_assembly(200000, 1000000000, seed=20261019 + 2)
dev = torch.device("cuda:0")
offsets = np.ascontiguousarray(asm.offsets, dtype=np.int64)
starts, lengths, total = A.pack_plan(offsets)
n = asm.n_contigs
d_seq = torch.empty(asm.total_bp + 64, dtype=torch.uint8, device=dev)
d_seq[: asm.total_bp].copy_(torch.from_numpy(asm.seq[: asm.total_bp]))
d_off = torch.from_numpy(offsets).to(dev)
d_starts = torch.from_numpy(starts).to(dev)
for gc in (False, True):
    for _ in range(3):
        p = A.pack_device_arrays(d_seq, d_off, d_starts, n, total, lengths, starts, asm.total_bp, gc=gc, exc_capacity=1 << 20)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        p = A.pack_device_arrays(d_seq, d_off, d_starts, n, total, lengths, starts, asm.total_bp, gc=gc, exc_capacity=1 << 20)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("pack gc=%s: %.3f ms (%.0f GB/s of ASCII)" % (gc, ms, asm.total_bp / ms / 1e6), flush=True)
gc = p.gc_content.cpu().numpy()
# size-independent check: recompute GC% of 2000 random contigs on the host
rng = np.random.default_rng(0)
for i in rng.integers(0, n, 2000):
    s = asm.seq[offsets[i]:offsets[i + 1]]
    u = s & 0xDF
    g = int(np.isin(u, [71, 67, 83]).sum()); a = int(np.isin(u, [65, 84, 87, 85]).sum())
    want = (g / (g + a) * 100) if g + a else 0.0
    assert gc[i] == want, (i, gc[i], want)
print("gc spot check ok; mean gc %.3f" % gc.mean())
