"""Per-eps cluster metrics and the cutoff filter used by the eps sweep.

Host-side mirror of ``autometa.binning.utilities.add_metrics`` /
``apply_binning_metrics_filter`` (reference binning/utilities.py:125-262) with the
same inputs, outputs and NA conventions, computed with array reductions over
the label vector instead of a pandas join + groupby.  ``cluster_metrics_arrays``
is the array-level core the sweep calls once per eps without building frames.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import pandas as pd

METRICS = [
    "completeness",
    "purity",
    "coverage_stddev",
    "gc_content_stddev",
    "present_marker_count",
    "single_copy_marker_count",
]


def cluster_metrics_arrays(
    labels: np.ndarray, n_clusters: int, marker_rows: np.ndarray, marker_cols: np.ndarray,
    marker_vals: np.ndarray, n_ref_markers: int, coverage: np.ndarray, gc_content: np.ndarray,
) -> Dict[str, np.ndarray]:
    """Metrics per cluster 0..n_clusters-1 (reference utilities.py:176-192).

    ``marker_*`` is the non-zero part of the contig x marker count table in COO form
    (rows index into ``labels``).  completeness = present / M * 100; purity =
    single-copy / present * 100 (NaN when nothing is present); std with ddof=1
    (NaN for single-contig clusters).
    """
    M = int(n_ref_markers)
    if len(marker_rows):
        key = labels[marker_rows].astype(np.int64) * M + marker_cols
        uk, inv = np.unique(key, return_inverse=True)
        sums = np.bincount(inv, weights=marker_vals, minlength=len(uk))
        clu = uk // M
        present = np.bincount(clu[sums >= 1], minlength=n_clusters)
        single = np.bincount(clu[sums == 1], minlength=n_clusters)
    else:
        present = np.zeros(n_clusters, dtype=np.int64)
        single = np.zeros(n_clusters, dtype=np.int64)
    with np.errstate(invalid="ignore", divide="ignore"):
        completeness = present / M * 100
        purity = single / present * 100
    cnt = np.bincount(labels, minlength=n_clusters).astype(np.float64)

    def _std(v):
        mean = np.bincount(labels, weights=v, minlength=n_clusters) / cnt
        d = v - mean[labels]
        ss = np.bincount(labels, weights=d * d, minlength=n_clusters)
        with np.errstate(invalid="ignore", divide="ignore"):
            return np.sqrt(ss / (cnt - 1.0))

    return dict(
        completeness=completeness,
        purity=purity,
        coverage_stddev=_std(coverage),
        gc_content_stddev=_std(gc_content),
        present_marker_count=present.astype(np.int64),
        single_copy_marker_count=single.astype(np.int64),
    )


def passing_mask(m: Dict[str, np.ndarray], completeness_cutoff, purity_cutoff, coverage_stddev_cutoff,
                 gc_content_stddev_cutoff) -> np.ndarray:
    """utilities.py:257-262: >=, >=, <=, <= ; comparisons with NaN are False."""
    with np.errstate(invalid="ignore"):
        return (
            (m["completeness"] >= completeness_cutoff)
            & (m["purity"] >= purity_cutoff)
            & (m["coverage_stddev"] <= coverage_stddev_cutoff)
            & (m["gc_content_stddev"] <= gc_content_stddev_cutoff)
        )


class _MarkerTable:
    """The non-zero entries of a contig x marker count table in CSR form, built once per table: the sweeps of
    a binning run (hundreds of taxa, several rounds each) all restrict the SAME marker table to their own
    contigs, and converting the whole frame each time costs more than the sweep itself."""

    def __init__(self, markers_df: pd.DataFrame):
        vals = markers_df.to_numpy(dtype=np.float64, na_value=0.0)
        vals = np.nan_to_num(vals, nan=0.0)
        r, c = np.nonzero(vals)
        self.index = markers_df.index
        self.shape = markers_df.shape
        self.indptr = np.searchsorted(r, np.arange(markers_df.shape[0] + 1)).astype(np.int64)
        self.cols = c.astype(np.int64)
        self.vals = vals[r, c]

    def coo(self, index: pd.Index):
        mrow = self.index.get_indexer(index)  # row of every contig of the table in the marker table, or -1
        have = np.flatnonzero(mrow >= 0)
        starts = self.indptr[mrow[have]]
        lens = self.indptr[mrow[have] + 1] - starts
        total = int(lens.sum())
        if total == 0:
            z = np.zeros(0, dtype=np.int64)
            return z, z.copy(), np.zeros(0, dtype=np.float64)
        rows = np.repeat(have, lens)
        first = np.cumsum(lens) - lens  # position of each contig's first entry in the output
        offs = np.repeat(starts - first, lens) + np.arange(total, dtype=np.int64)
        return rows.astype(np.int64), self.cols[offs], self.vals[offs]


_marker_tables: Dict[int, tuple] = {}


def _marker_table(markers_df: pd.DataFrame) -> "_MarkerTable":
    import weakref

    key = id(markers_df)
    hit = _marker_tables.get(key)
    if hit is not None and hit[0]() is markers_df and hit[1].shape == markers_df.shape:
        return hit[1]
    table = _MarkerTable(markers_df)
    for k in [k for k, (ref, _) in _marker_tables.items() if ref() is None]:
        del _marker_tables[k]
    _marker_tables[key] = (weakref.ref(markers_df), table)
    return table


def markers_coo(markers_df: pd.DataFrame, index: pd.Index):
    """Non-zero entries of ``markers_df`` restricted/aligned to ``index`` (contigs
    absent from the marker table count as zero, as the outer join + sum does), as
    ``(row in index, marker column, count)``."""
    if markers_df.index.is_unique:
        return _marker_table(markers_df).coo(index)
    pos = index.get_indexer(markers_df.index)
    ok = pos >= 0
    vals = markers_df.to_numpy(dtype=np.float64, na_value=0.0)[ok]
    vals = np.nan_to_num(vals, nan=0.0)
    r, c = np.nonzero(vals)
    return pos[ok][r].astype(np.int64), c.astype(np.int64), vals[r, c]


def add_metrics(df: pd.DataFrame, markers_df: pd.DataFrame) -> Tuple[pd.DataFrame, pd.DataFrame]:
    """Add per-cluster metrics to each contig of ``df`` (reference utilities.py:125-207).

    Returns ``(contig_metrics_df, cluster_metrics_df)`` with the reference's columns.
    """
    n_ref = markers_df.shape[1]
    if "cluster" not in df.columns:
        cluster_metrics_df = pd.DataFrame(data={m: pd.NA for m in METRICS}, index=df.index)
        contig_metrics_df = df.copy().convert_dtypes().drop(columns=METRICS, errors="ignore")
        contig_metrics_df[METRICS] = pd.NA
        return contig_metrics_df, cluster_metrics_df
    df = df.drop(columns=METRICS, errors="ignore")
    codes, uniques = pd.factorize(df["cluster"], sort=True)  # NA -> -1 (groupby drops NA keys)
    valid = codes >= 0
    n_clusters = len(uniques)
    sub_index = df.index[valid]
    rows, cols, vals = markers_coo(markers_df, sub_index)
    m = cluster_metrics_arrays(
        codes[valid], n_clusters, rows, cols, vals, n_ref,
        df["coverage"].to_numpy(dtype=np.float64, na_value=np.nan)[valid],
        df["gc_content"].to_numpy(dtype=np.float64, na_value=np.nan)[valid],
    )
    cluster_metrics_df = pd.DataFrame(m, index=pd.Index(uniques, name="cluster"))[METRICS]
    contig_metrics_df = pd.merge(df, cluster_metrics_df, how="left", left_on="cluster", right_index=True)
    return contig_metrics_df, cluster_metrics_df


def apply_binning_metrics_filter(
    df: pd.DataFrame,
    completeness_cutoff: float = 20.0,
    purity_cutoff: float = 95.00,
    coverage_stddev_cutoff: float = 25.0,
    gc_content_stddev_cutoff: float = 5.0,
) -> pd.DataFrame:
    """Rows of ``df`` passing the four cutoffs (reference utilities.py:210-262)."""
    metrics = ["completeness", "purity", "coverage_stddev", "gc_content_stddev"]
    for metric in metrics:
        if metric not in df.columns:
            raise KeyError(f"{metric} not in `df` columns: {df.columns}")
    return df[
        df.completeness.ge(completeness_cutoff)
        & df.purity.ge(purity_cutoff)
        & df.coverage_stddev.le(coverage_stddev_cutoff)
        & df.gc_content_stddev.le(gc_content_stddev_cutoff)
    ]


def reindex_bin_names(df: pd.DataFrame, cluster_col: str = "cluster", initial_index: int = 0) -> pd.DataFrame:
    """Rename the clusters of ``df`` to ``bin_<initial_index+1>``, ``bin_<initial_index+2>``, ... in
    order of first appearance; rows without a cluster stay null (reference utilities.py:265-325).
    Modifies and returns ``df``."""
    assigned = df[cluster_col].dropna()
    names = {old: f"bin_{initial_index + 1 + rank}" for rank, old in enumerate(assigned.unique())}
    df[cluster_col] = assigned.map(names.__getitem__)
    return df


def zero_pad_bin_names(df: pd.DataFrame, cluster_col: str = "cluster") -> pd.DataFrame:
    """Renumber the clusters ``bin_001`` ... in order of first appearance, zero-padded to the
    number of digits of the cluster count (reference utilities.py:328-363).  Modifies and returns ``df``."""
    assigned = df[cluster_col].dropna()
    width = len(str(assigned.nunique()))
    names = {old: "bin_" + str(rank + 1).zfill(width) for rank, old in enumerate(assigned.unique())}
    df[cluster_col] = assigned.map(names.__getitem__)
    return df
