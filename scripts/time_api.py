"""Wall time of the drop-in Python API, files in / files out, as `autometa-kmers` runs it
(reference kmers.py:709-851): count(fasta, out=tsv) -> normalize(df, out=tsv) -> pca(norm) on C1
(10k contigs / 50 Mbp) and C2 (200k contigs / 1 Gbp).  Run on the GPU box; prints one JSON line per config."""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from autometa_b200 import synth
from autometa_b200.common import kmers

for name, n, bp, seed in (("C1", 10_000, 50_000_000, 20261019 + 1), ("C2", 200_000, 1_000_000_000, 20261019 + 2)):
    if len(sys.argv) > 1 and name not in sys.argv[1:]:
        continue
    asm = synth.make_assembly(n, bp, seed=seed)
    tmp = tempfile.mkdtemp()
    fa = os.path.join(tmp, "asm.fna")
    synth.write_fasta(asm, fa)
    del asm
    kmers.count(os.path.join(ROOT, "tests", "golden", "small.fna"), size=5)  # CUDA context, library load
    out = {}
    t0 = time.perf_counter()
    df = kmers.count(fa, size=5, out=os.path.join(tmp, "kmers.tsv"))
    out["count_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    norm = kmers.normalize(df, method="am_clr", out=os.path.join(tmp, "kmers.am_clr.tsv"))
    out["normalize_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    res = kmers.pca(norm.to_numpy(), n_components=50)
    out["pca_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    kmers.load(os.path.join(tmp, "kmers.am_clr.tsv"))
    out["load_normalized_s"] = time.perf_counter() - t0
    tot = out["count_s"] + out["normalize_s"] + out["pca_s"]
    print(json.dumps({"config": f"{name}: {n} contigs / {bp / 1e6:.0f} Mbp, count(fasta,out) -> normalize(am_clr,out) -> pca(50)",
                      **{k: round(v, 3) for k, v in out.items()}, "total_s": round(tot, 3),
                      "Mbp_per_s": round(bp / 1e6 / tot, 1),
                      "fasta_MB": round(os.path.getsize(fa) / 1e6, 1),
                      "tsv_MB": [round(os.path.getsize(os.path.join(tmp, f)) / 1e6, 1) for f in ("kmers.tsv", "kmers.am_clr.tsv")]}),
          flush=True)
    for f in os.listdir(tmp):
        os.remove(os.path.join(tmp, f))
