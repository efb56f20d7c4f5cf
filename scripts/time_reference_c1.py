"""Time the UNMODIFIED Python reference (loaded through oracle/ref_shim.py, Biopython replaced by the
shim's stub) on C1 = BASELINE.json configs[0]: 10k contigs / 50 Mbp, 5-mer count + am_clr (+ PCA(50) as
kmers.py:605 executes it), with cpus = the host's cores.  BUILD CONTAINER ONLY (/root/reference does not
travel).  Also times this repo's C port of the same loops on the same input for scale.  Prints one JSON line."""
import json, os, sys, tempfile, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")
import numpy as np
from oracle import ref_shim, oracle_c
from autometa_b200 import synth

kmers, rdb, butil = ref_shim.load()
cpus = os.cpu_count()
asm = synth.make_assembly(10_000, 50_000_000, seed=20261019 + 1)
tmp = tempfile.mkdtemp()
fa = os.path.join(tmp, "c1.fna")
synth.write_fasta(asm, fa)
t0 = time.perf_counter()
df = kmers.count(fa, size=5, out=os.path.join(tmp, "kmers.tsv"), cpus=cpus)
t_count = time.perf_counter() - t0
t0 = time.perf_counter()
norm = kmers.normalize(kmers.load(os.path.join(tmp, "kmers.tsv")), method="am_clr")
t_norm = time.perf_counter() - t0
from sklearn.decomposition import PCA
t0 = time.perf_counter()
PCA(n_components=50, random_state=np.random.RandomState(42)).fit_transform(norm.to_numpy())
t_pca = time.perf_counter() - t0
t0 = time.perf_counter()
c = oracle_c.count(asm.seq, asm.offsets, 5, cpus)
x = oracle_c.am_clr(c, cpus)
t_port = time.perf_counter() - t0
import pandas as pd
assert np.array_equal(df.fillna(0).to_numpy(dtype=np.int64), c.astype(np.int64))
tot = t_count + t_norm + t_pca
print(json.dumps({
    "config": "C1: 10k contigs / 50 Mbp, 5-mer count + am_clr + PCA(50)", "cores": cpus,
    "reference_python": {"count_s": round(t_count, 2), "count_includes": "Pool(cpus).map(record_counter) + TSV write",
                         "am_clr_s": round(t_norm, 2), "pca_s": round(t_pca, 2), "total_s": round(tot, 2),
                         "Mbp_per_s": round(50.0 / tot, 3), "count_only_Mbp_per_s": round(50.0 / t_count, 3)},
    "c_port_count_plus_am_clr_s": round(t_port, 3),
    "note": "Biopython is replaced by the shim's stub Seq/SeqIO (oracle/ref_shim.py); counts of the two arms are identical",
}))
