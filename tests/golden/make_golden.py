"""Generate the committed golden vectors by running the UNMODIFIED reference
(/root/reference, loaded through oracle/ref_shim.py) on small seeded inputs.

Run in the build container only:  python tests/golden/make_golden.py
Outputs (committed): tests/golden/*.npz, tests/golden/small.fna
The reference ships no fixtures for this path (tests/data is not in the tree),
so these outputs of the reference's own code are what pins the oracle.
"""
import os
import sys
import tempfile
import warnings

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import ref_shim  # noqa: E402

kmers, rdb, butil = ref_shim.load()


def make_small_fasta(path):
    rng = np.random.default_rng(20261019)
    recs = []
    lens = [3, 4, 5, 6, 7, 40, 63, 64, 65, 66, 67, 68, 69, 127, 128, 129, 133, 300, 2000, 2047, 2048, 2049,
            2052, 2053, 2054, 4100, 5000, 33000, 70001]
    lens += [int(x) for x in rng.integers(70, 4000, size=40)]
    for i, L in enumerate(lens):
        s = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=L)].copy()
        if i % 3 == 0 and L > 20:  # N runs
            for _ in range(1 + L // 1500):
                a = int(rng.integers(0, L - 1))
                s[a : a + int(rng.integers(1, 50))] = ord("N")
        if i % 4 == 1 and L > 30:  # soft-masked run
            a = int(rng.integers(0, L - 12))
            s[a : a + int(rng.integers(5, 200))] |= 0x20
        if i % 5 == 2 and L > 10:  # IUPAC letters
            for a in rng.integers(0, L, size=3):
                s[a] = ord("RYKMSWBDHVU"[int(rng.integers(0, 11))])
        recs.append((f"contig_{i}_len{L}", s.tobytes()))
    # boundary cases: invalid base at group / tile edges
    s = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=4200)].copy()
    for a in (59, 60, 63, 64, 67, 68, 127, 2043, 2047, 2048, 2051, 4095, 4096, 4194, 4195, 4199):
        s[a] = ord("N")
    recs.append(("edges", s.tobytes()))
    recs.append(("allN", b"N" * 100))
    recs.append(("homopolymer", b"A" * 500))
    recs.append(("dup", b"ACGTACGTAGCTAGCTAGCATCGATCGATCGACTAGCTAGC"))
    recs.append(("dup", b"TTTTTTTTTTGGGGGGGGGGCCCCCCCCCCAAAAAAAAAA"))
    with open(path, "w") as fh:
        for j, (cid, s) in enumerate(recs):
            desc = " some description" if j % 2 else ""
            fh.write(f">{cid}{desc}\n")
            s = s.decode()
            w = 60 if j % 7 else 71
            for a in range(0, len(s), w):
                fh.write(s[a : a + w] + "\n")
    return recs


def _frame_to_npz(prefix, df, res):
    res[f"{prefix}_index"] = np.array(df.index.tolist())
    res[f"{prefix}_columns"] = np.array(df.columns.tolist())
    res[f"{prefix}_cluster"] = np.array(["" if pd.isna(c) else str(c) for c in df["cluster"]])
    for m in ["completeness", "purity", "coverage_stddev", "gc_content_stddev"]:
        res[f"{prefix}_{m}"] = pd.to_numeric(df[m], errors="coerce").to_numpy(dtype=np.float64, na_value=np.nan)


def make_binning_golden():
    """The sweep's callers through the reference: get_clusters and taxon_guided_binning."""
    from autometa_b200 import synth

    res = {}
    table, markers = synth.make_sweep_input(n_points=1500, n_clusters=9, n_markers=40, seed=21)
    df = rdb.get_clusters(table.copy(), markers, 20.0, 90.0, 25.0, 5.0, method="dbscan", n_jobs=1)
    _frame_to_npz("get_clusters", df, res)
    # a table where nothing passes: every contig stays unclustered with NA metrics
    df0 = rdb.get_clusters(table.iloc[:40].copy(), markers, 99.0, 99.0, 0.001, 0.001, method="dbscan", n_jobs=1)
    _frame_to_npz("get_clusters_none", df0, res)
    tax = synth.add_taxonomy(table, seed=3)
    dft = rdb.taxon_guided_binning(tax.copy(), markers, 20.0, 90.0, 25.0, 5.0, method="dbscan", n_jobs=1)
    _frame_to_npz("taxon", dft, res)
    dfr = rdb.taxon_guided_binning(tax.copy(), markers, 20.0, 90.0, 25.0, 5.0, starting_rank="phylum",
                                   method="dbscan", reverse_ranks=True, n_jobs=1)
    _frame_to_npz("taxon_rev", dfr, res)
    np.savez_compressed(os.path.join(HERE, "binning_golden.npz"), **res)
    print("binning golden:", {k: v.shape for k, v in res.items() if k.endswith("_cluster")},
          "bins:", sorted(set(res["get_clusters_cluster"]))[:6], sorted(set(res["taxon_cluster"]))[:6])


def main():
    if "binning" in sys.argv[1:]:
        return make_binning_golden()
    fna = os.path.join(HERE, "small.fna")
    make_small_fasta(fna)
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        # ---- count (k = 3, 4, 5, 6), serial and multiprocessing paths
        for k in (3, 4, 5, 6):
            tsv = os.path.join(tmp, f"k{k}.tsv")
            df = kmers.count(fna, size=k, out=tsv, force=True, verbose=False, cpus=1)
            dfl = kmers.load(tsv)
            out[f"count_k{k}"] = np.nan_to_num(dfl.to_numpy(dtype=np.float64), nan=0.0).astype(np.int32)
            if k == 5:
                out["count_ids"] = np.array(dfl.index.tolist())
                out["count_columns"] = np.array(dfl.columns.tolist())
                df_mp = kmers.count(fna, size=5, force=True, verbose=False, cpus=2)
                a = df_mp.fillna(0).to_numpy().astype(np.int64)
                b = df.fillna(0).to_numpy().astype(np.int64)
                assert np.array_equal(a, b)
                # ---- normalise through the TSV exactly like the pipeline (pandas-3 caveat)
                norm = kmers.normalize(dfl, method="am_clr")
                out["am_clr_values"] = norm.to_numpy(dtype=np.float64)
                out["am_clr_index"] = np.array(norm.index.tolist())
                out["am_clr_columns"] = np.array(norm.columns.tolist())
                ok = dfl.dropna(axis="index", how="all")
                out["clr_rows"] = np.array(ok.index.tolist())
                out["clr_values"] = kmers.normalize(ok, method="clr").to_numpy(dtype=np.float64)
                out["ilr_values"] = kmers.normalize(ok, method="ilr").to_numpy(dtype=np.float64)
        # am_clr with dropped columns: k=6 table restricted to a few contigs
        tsv6 = os.path.join(tmp, "k6.tsv")
        d6 = kmers.load(tsv6).iloc[5:12]
        n6 = kmers.normalize(d6, method="am_clr")
        out["am_clr6_rows"] = np.array(n6.index.tolist())
        out["am_clr6_columns"] = np.array(n6.columns.tolist())
        out["am_clr6_values"] = n6.to_numpy(dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "kmers_golden.npz"), **out)

    # ---- PCA through the reference's own call site (embed with the stub bh_sne returning the
    # first d PCA scores), N >= 10*D so sklearn takes covariance_eigh
    rng = np.random.default_rng(7)
    N, D, P = 700, 64, 10
    Z = rng.normal(size=(N, 12)) * np.array([9, 7, 5, 4, 3, 2.5, 2, 1.5, 1.2, 1.1, 1.05, 1.0])
    X = Z @ rng.normal(size=(12, D)) + 0.05 * rng.normal(size=(N, D)) + rng.normal(size=D) * 3
    dfX = pd.DataFrame(X, index=pd.Index([f"c{i}" for i in range(N)], name="contig"))
    emb = kmers.embed(dfX, method="bhsne", embed_dimensions=3, pca_dimensions=P, seed=42)
    from sklearn.decomposition import PCA

    p = PCA(n_components=P, random_state=np.random.RandomState(42))
    sc = p.fit_transform(X)
    assert p._fit_svd_solver == "covariance_eigh"
    assert np.allclose(sc[:, :3], emb.to_numpy())
    np.savez_compressed(
        os.path.join(HERE, "pca_golden.npz"), X=X, scores=sc, explained_variance=p.explained_variance_,
        explained_variance_ratio=p.explained_variance_ratio_, components=p.components_, mean=p.mean_,
        embed_first3=emb.to_numpy(),
    )

    # ---- DBSCAN step + sweep through the reference's run_dbscan / recursive_dbscan
    from autometa_b200 import synth

    table, markers = synth.make_sweep_input(n_points=3000, n_clusters=12, n_markers=40, seed=11)
    res = {}
    eps_list = [0.3, 0.7999999999999999, 1.2, 2.0, 3.5000000000000013, 5.0, 9.0]
    for i, eps in enumerate(eps_list):
        lab = rdb.run_dbscan(table.copy(), eps=eps, n_jobs=1)["cluster"].to_numpy()
        res[f"labels_{i}"] = lab.astype(np.int64)
    res["eps_list"] = np.array(eps_list)
    clustered, unclustered = rdb.recursive_dbscan(table.copy(), markers, 20.0, 90.0, 25.0, 5.0, n_jobs=1)
    res["clustered_index"] = np.array(clustered.index.tolist())
    res["clustered_cluster"] = clustered["cluster"].to_numpy().astype(np.int64)
    for m in ["completeness", "purity", "coverage_stddev", "gc_content_stddev"]:
        res[f"clustered_{m}"] = clustered[m].to_numpy(dtype=np.float64)
    res["clustered_columns"] = np.array(clustered.columns.tolist())
    res["unclustered_index"] = np.array(unclustered.index.tolist())
    res["unclustered_cluster"] = unclustered["cluster"].to_numpy().astype(np.int64)
    # add_metrics at one eps
    binned = rdb.run_dbscan(table.copy(), eps=1.2, n_jobs=1)
    cdf, mdf = butil.add_metrics(binned, markers)
    res["metrics_cluster_index"] = mdf.index.to_numpy().astype(np.int64)
    for m in ["completeness", "purity", "coverage_stddev", "gc_content_stddev", "present_marker_count",
              "single_copy_marker_count"]:
        res[f"metrics_{m}"] = mdf[m].to_numpy(dtype=np.float64)
        res[f"contig_{m}"] = cdf[m].to_numpy(dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "dbscan_golden.npz"), **res)
    make_binning_golden()
    print("golden vectors written:", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
