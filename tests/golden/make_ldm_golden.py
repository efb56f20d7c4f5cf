"""Golden outputs of the UNMODIFIED reference ``large_data_mode.cluster_by_taxon_partitioning`` (loaded
through oracle/ref_shim.py; ``tsne.bh_sne`` stubbed to return the first d PCA scores, as for the embed golden)
on the seeded inputs of tests/_ldm_input.py.  Build container only:  python tests/golden/make_ldm_golden.py
Output (committed): tests/golden/ldm_golden.npz — final frames of two runs (max_partition_size 10000 and 400,
the latter with a cache directory: its embedding-cache file list and the checkpoint columns are kept too)."""
import os
import sys
import tempfile
import warnings

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.filterwarnings("ignore")

from oracle import ref_shim  # noqa: E402

ref_shim.load()
from autometa.binning import large_data_mode as ldm  # noqa: E402

import _ldm_input  # noqa: E402


def frame_to_dict(df, prefix, out):
    out[f"{prefix}_index"] = np.array(df.index.tolist())
    out[f"{prefix}_columns"] = np.array(df.columns.tolist())
    out[f"{prefix}_cluster"] = np.array(["" if pd.isna(c) else str(c) for c in df["cluster"]])
    for m in ["completeness", "purity", "coverage_stddev", "gc_content_stddev"]:
        out[f"{prefix}_{m}"] = pd.to_numeric(df[m], errors="coerce").to_numpy(dtype=np.float64, na_value=np.nan)


def main():
    main_df, counts, markers = _ldm_input.make_ldm_input()
    out = {}
    kw = dict(norm_method="am_clr", pca_dimensions=10, embed_dimensions=2, embed_method="bhsne", completeness=20.0,
              purity=90.0, coverage_stddev=25.0, gc_content_stddev=5.0, starting_rank="superkingdom", method="dbscan",
              n_jobs=1)
    df = ldm.cluster_by_taxon_partitioning(main_df.copy(), counts.copy(), markers, max_partition_size=10000, **kw)
    frame_to_dict(df, "ldm", out)
    print("run 1:", df["cluster"].notna().sum(), "binned,", df["cluster"].nunique(), "bins")
    with tempfile.TemporaryDirectory() as cache:
        df2 = ldm.cluster_by_taxon_partitioning(main_df.copy(), counts.copy(), markers, max_partition_size=400,
                                                cache=cache, **kw)
        frame_to_dict(df2, "ldm400", out)
        files = sorted(os.path.relpath(os.path.join(d, f), cache) for d, _, fs in os.walk(cache) for f in fs)
        out["ldm400_cache_files"] = np.array(files)
        # the cached embeddings themselves (first two PCA scores through the stub bh_sne): small, and the
        # sharpest diagnostic when a bin differs
        for f in files:
            if "binning_checkpoints" in f:
                continue
            emb = pd.read_csv(os.path.join(cache, f), sep="\t", index_col="contig")
            out[f"emb::{f}::index"] = np.array(emb.index.tolist())
            out[f"emb::{f}::values"] = emb.to_numpy(dtype=np.float64)
        info = ldm.get_checkpoint_info(os.path.join(cache, "binning_checkpoints.tsv.gz"))
        out["ldm400_checkpoint_columns"] = np.array(info["binning_checkpoints"].columns.tolist())
        out["ldm400_checkpoint_rank"] = np.array([info["starting_rank"], info["starting_rank_name_txt"]])
        print("run 2:", df2["cluster"].notna().sum(), "binned,", df2["cluster"].nunique(), "bins;", len(files), "cache files")
    np.savez_compressed(os.path.join(HERE, "ldm_golden.npz"), **out)
    print("written ldm_golden.npz")


if __name__ == "__main__":
    main()
