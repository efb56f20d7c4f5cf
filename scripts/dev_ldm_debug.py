import os, sys, tempfile, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, pandas as pd, torch
from autometa_b200 import synth
from autometa_b200.binning import large_data_mode as M
from autometa_b200.common import kmers
from oracle import oracle_c, oracle_np as O

stub = types.ModuleType("tsne"); stub.bh_sne = lambda data, d, **kw: np.ascontiguousarray(data[:, :d]); sys.modules["tsne"] = stub

def np_pca(X, P):
    mu = X.mean(0); C = np.cov(X, rowvar=False)
    w, V = np.linalg.eigh(C); w, V = w[::-1][:P], V[:, ::-1][:, :P].T.copy()
    s = np.sign(V[np.arange(P), np.argmax(np.abs(V), axis=1)]); V *= s[:, None]
    return (X - mu) @ V.T, w

asm = synth.make_assembly(2600, 2600 * 3300, seed=41)
cnt = oracle_c.count(asm.seq, asm.offsets, 5, 1).astype(np.float64); cnt[cnt == 0] = np.nan
counts = pd.DataFrame(cnt, index=pd.Index(asm.ids, name="contig"), columns=O.kmer_columns(5)[0])
for method in ("am_clr", "clr", "ilr"):
    for n in (40, 60, 100, 130, 300, 700):
        idx = counts.index[500:500 + n]
        got = M.PartitionPCA(counts, method, 20).run({"p": idx})["p"].to_numpy()
        norm = kmers.normalize(counts.loc[idx], method=method).to_numpy()
        seq = kmers.pca(norm, n_components=20)
        ref, w = np_pca(norm, 20)
        sc = np.max(np.abs(ref))
        print(f"{method} n={n}: fused-vs-numpy {np.max(np.abs(got - ref)) / sc:.2e}  sequential-vs-numpy {np.max(np.abs(seq['scores'] - ref)) / sc:.2e}"
              f"  ev seq {np.max(np.abs(seq['explained_variance'] - w) / w):.2e}", flush=True)

g = np.load(os.path.join(ROOT, "tests", "golden", "ldm_golden.npz"))
import _ldm_input
main, counts4, markers = _ldm_input.make_ldm_input()
kw = dict(norm_method="am_clr", pca_dimensions=10, embed_dimensions=2, embed_method="bhsne", completeness=20.0,
          purity=90.0, coverage_stddev=25.0, gc_content_stddev=5.0, starting_rank="superkingdom", method="dbscan")
cache = tempfile.mkdtemp()
df2 = M.cluster_by_taxon_partitioning(main.copy(), counts4.copy(), markers, max_partition_size=400, cache=cache, **kw)
files = sorted(os.path.relpath(os.path.join(d, f), cache) for d, _, fs in os.walk(cache) for f in fs)
print("cache files equal:", files == g["ldm400_cache_files"].tolist())
if files != g["ldm400_cache_files"].tolist():
    print(" ours-only", sorted(set(files) - set(g["ldm400_cache_files"].tolist())), " ref-only", sorted(set(g["ldm400_cache_files"].tolist()) - set(files)))
for f in g["ldm400_cache_files"].tolist():
    if "binning_checkpoints" in f or f not in files: continue
    emb = pd.read_csv(os.path.join(cache, f), sep="\t", index_col="contig")
    ri, rv = g[f"emb::{f}::index"].tolist(), g[f"emb::{f}::values"]
    same_idx = emb.index.tolist() == ri
    d = np.max(np.abs(emb.to_numpy() - rv)) / np.max(np.abs(rv)) if same_idx and emb.shape == rv.shape else float("nan")
    print(f"{f}: rows {emb.shape[0]} index equal {same_idx} max rel diff {d:.2e}")
got = np.array(["" if pd.isna(c) else str(c) for c in df2["cluster"]])
print("index equal", df2.index.tolist() == g["ldm400_index"].tolist(), "labels equal", np.array_equal(got, g["ldm400_cluster"]),
      "n binned ours/ref", (got != "").sum(), (g["ldm400_cluster"] != "").sum())
# partition-level agreement up to renaming
if df2.index.tolist() == g["ldm400_index"].tolist():
    pairs = set(zip(got.tolist(), g["ldm400_cluster"].tolist()))
    print("distinct (ours, ref) label pairs:", len(pairs), "ours bins", len(set(got.tolist())), "ref bins", len(set(g["ldm400_cluster"].tolist())))
    print(sorted(pairs)[:30])
