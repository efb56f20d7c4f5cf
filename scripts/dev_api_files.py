"""File-level pipeline through the public API (count -> kmers.tsv -> normalize -> tsv -> PCA), timing the
fast TSV I/O against pandas on the same frames."""
import sys, os, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd, torch
from autometa_b200 import synth, tsv
from autometa_b200.common import kmers
N, BP = int(sys.argv[1]) if len(sys.argv) > 1 else 50000, int(sys.argv[2]) if len(sys.argv) > 2 else 250_000_000
asm = synth.make_assembly(N, BP, seed=5)
tmp = tempfile.mkdtemp()
fa = os.path.join(tmp, "a.fna")
with open(fa, "wb") as fh:
    for i in range(asm.n_contigs):
        fh.write(b">" + asm.ids[i].encode() + b"\n" + asm.seq[asm.offsets[i]:asm.offsets[i + 1]].tobytes() + b"\n")
torch.zeros(1, device="cuda")
def T(label, f):
    t = time.time(); r = f(); print("%-44s %7.2f s" % (label, time.time() - t), flush=True); return r
counts = T("count(fasta, out=kmers.tsv)", lambda: kmers.count(fa, size=5, out=os.path.join(tmp, "k.tsv"), force=True))
T("  of which: frame -> pandas to_csv", lambda: counts.iloc[: N // 10].to_csv(os.path.join(tmp, "k_pd.tsv"), sep="\t"))
loaded = T("load(kmers.tsv)", lambda: kmers.load(os.path.join(tmp, "k.tsv")))
T("  pandas read_csv (same file)", lambda: pd.read_csv(os.path.join(tmp, "k.tsv"), sep="\t", index_col="contig"))
norm = T("normalize(df, am_clr, out=norm.tsv)", lambda: kmers.normalize(loaded, method="am_clr", out=os.path.join(tmp, "n.tsv"), force=True))
T("  pandas to_csv of 1/10 of the rows", lambda: norm.iloc[: N // 10].to_csv(os.path.join(tmp, "n_pd.tsv"), sep="\t"))
nl = T("load(norm.tsv)", lambda: kmers.load(os.path.join(tmp, "n.tsv")))
T("  pandas read_csv (same file)", lambda: pd.read_csv(os.path.join(tmp, "n.tsv"), sep="\t", index_col="contig"))
res = T("pca(norm, 50)", lambda: kmers.pca(nl.to_numpy(), 50))
print("sizes: k.tsv %.1f MB, n.tsv %.1f MB" % (os.path.getsize(os.path.join(tmp, "k.tsv")) / 1e6, os.path.getsize(os.path.join(tmp, "n.tsv")) / 1e6))
