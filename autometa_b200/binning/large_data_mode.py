"""Drop-in for ``autometa.binning.large_data_mode`` (SURVEY 8f-3 second half, 8f-4 caches).

The reference walks the canonical ranks and, inside each rank, every taxon: the taxon's contigs are
normalised, reduced by PCA and embedded (``get_kmer_embedding`` -> ``kmers.normalize`` + ``kmers.embed``,
large_data_mode.py:56-114), then swept by ``get_clusters`` (:382-582).  That is hundreds of small
(<= ``max_partition_size`` = 10 000 rows) normalise + PCA problems, each one bound by launch latency and by
the N-independent 1.2 ms eigensolve when run one after the other.

Here the partitions of a rank are known before its loop starts (a contig belongs to one taxon per rank, so
clustering taxon A never changes taxon B's contig set), so their normalise -> Gram -> eigensolve -> projection
chains are queued TOGETHER, round-robin over CUDA streams, against one device copy of the count table
(``PartitionPCA``): the 16-SM eigensolves of up to nine partitions run side by side and the launch gaps of one
chain are filled by the others.  The manifold step (UMAP / t-SNE; not on the north-star path) and the sweep
then run per taxon exactly as in the reference.  Control flow, bin naming, checkpoint and cache formats
(``binning_checkpoints.tsv.gz``, ``<rank>/<taxon>.<norm>_pca<P>_<method><d>.tsv.gz``) follow the reference so
that a run can be resumed from — and leaves behind — the same files.
"""
from __future__ import annotations

import gzip
import logging
import os
import re
from datetime import datetime
from types import SimpleNamespace
from typing import Dict, List, Optional, Union

import numpy as np
import pandas as pd

from .. import _lib as L
from .. import pipeline as _pipe
from .. import tsv as _tsv
from ..common import kmers
from ..common.exceptions import BinningError
from .recursive_dbscan import get_clusters_many
from .utilities import reindex_bin_names, zero_pad_bin_names

logger = logging.getLogger(__name__)

#: autometa/taxonomy/database.py: TaxonomyDatabase.CANONICAL_RANKS
CANONICAL_RANKS = ["species", "genus", "family", "order", "class", "phylum", "superkingdom", "root"]


# ----------------------------------------------------------------------------- batched normalise + PCA
class PartitionPCA:
    """normalise + PCA(P) of many row subsets of ONE k-mer count table, queued together on the device.

    ``counts``: the frame ``kmers.count`` / ``kmers.load`` returns (NA = 0).  ``run({key: index})`` returns
    ``{key: DataFrame[n_rows x P]}`` of PCA scores — what ``kmers.embed`` computes at kmers.py:588-607 from
    ``kmers.normalize(counts.loc[index], method)`` — for every partition whose shape lets the reference run the
    PCA (more columns than ``pca_dimensions``, at least ``pca_dimensions`` rows)."""

    def __init__(self, counts: pd.DataFrame, norm_method: str, pca_dimensions: int, device=None, n_streams: int = 8):
        import torch

        self.dev = L.require_cuda(device)
        self.method = norm_method.lower()
        self.P = int(pca_dimensions)
        mat, _ = kmers._counts_matrix(counts)
        self.index = counts.index
        self.columns = counts.columns
        self.pos = pd.Series(np.arange(len(counts.index), dtype=np.int64), index=counts.index)
        with torch.cuda.device(self.dev):
            self.counts = torch.from_numpy(mat).to(self.dev)
            self.streams = [torch.cuda.Stream(device=self.dev) for _ in range(max(1, n_streams))]
        longest = int(mat.sum(axis=1).max()) + 8 if mat.size else 1
        # the fixed-point range of the digit planes needs an upper bound of the contig length: windows + k
        self._fake = SimpleNamespace(_plans={}, lengths=np.array([longest], dtype=np.int64), n_contigs=len(counts.index))

    def _rows(self, index: pd.Index) -> np.ndarray:
        rows = self.pos.reindex(index).to_numpy()
        if np.isnan(rows).any():
            raise KeyError("partition holds contigs that are not in the count table")
        return np.sort(rows.astype(np.int64))  # counts.loc[counts.index.isin(...)] keeps the table's order

    def run(self, partitions: Dict[str, pd.Index]) -> Dict[str, pd.DataFrame]:
        import torch

        D = self.counts.shape[1]
        queued = []
        with torch.cuda.device(self.dev):
            cur = torch.cuda.current_stream(self.dev)
            for i, (key, index) in enumerate(partitions.items()):
                rows = self._rows(index)
                st = self.streams[i % len(self.streams)]
                st.wait_stream(cur)
                with torch.cuda.stream(st):
                    ridx = torch.from_numpy(rows).to(self.dev)
                    present = torch.empty(D, dtype=torch.int32, device=self.dev)
                    row_sum = torch.empty(max(len(rows), 1), dtype=torch.int32, device=self.dev)
                    L.check(L.lib.am_table_stats(L.t_ptr(self.counts), L.t_ptr(ridx), len(rows), D, L.t_ptr(present),
                                                 L.t_ptr(row_sum), L.stream_ptr(self.dev)))
                    fake = SimpleNamespace(_plans={}, lengths=self._fake.lengths, n_contigs=len(rows))
                    fit = _pipe._fit(fake, self.counts, (present, row_sum), self.method, ridx, None, None, False,
                                     lambda name: None)
                    res = _pipe._project(fit, self.P, lambda name: None)
                queued.append((key, rows, ridx, present, row_sum, res, st))
            out = {}
            for key, rows, ridx, present, row_sum, res, st in queued:
                st.synchronize()
                flags = res["stats"].cpu().numpy()
                cols_ok, rows_ok = bool((flags[:D] > 0).all()), bool(flags[D] == 0)
                keep_rows = rows
                if self.method != "am_clr":
                    # clr / ilr keep every column (zeros are replaced, kmers.py:418-421); an empty row is an error
                    if not rows_ok:
                        raise ValueError("Input matrix cannot have rows with all zeros")
                elif not (cols_ok and rows_ok):
                    # am_clr drops the k-mers absent from this partition and its contigs without counts
                    # (kmers.py:369): the general path with explicit gathers, on the caller's stream
                    col_index = None if cols_ok else torch.nonzero(present > 0).flatten().to(torch.int32)
                    nz = (row_sum[: len(rows)] > 0).cpu().numpy()
                    keep_rows = rows[nz]
                    ridx2 = torch.from_numpy(keep_rows).to(self.dev)
                    fit = _pipe._fit(SimpleNamespace(_plans={}, lengths=self._fake.lengths, n_contigs=len(keep_rows)),
                                     self.counts, None, self.method, ridx2, col_index, None, False, lambda name: None)
                    res = _pipe._project(fit, self.P, lambda name: None)
                scores = res["scores"].cpu().numpy()
                out[key] = pd.DataFrame(scores, index=self.index[keep_rows])
        return out


def _pca_applies(n_rows: int, n_cols: int, norm_method: str, pca_dimensions: int) -> bool:
    d_out = n_cols - 1 if norm_method == "ilr" else n_cols
    return pca_dimensions != 0 and d_out > pca_dimensions and n_rows >= pca_dimensions


# ----------------------------------------------------------------------------- embeddings + caches
def get_kmer_embedding(
    counts: pd.DataFrame,
    cache_fpath: str,
    norm_method: str,
    pca_dimensions: int,
    embed_dimensions: int,
    embed_method: str,
    pca_scores: Optional[pd.DataFrame] = None,
) -> pd.DataFrame:
    """Embedding of ``counts`` (large_data_mode.py:56-114): from ``cache_fpath`` when it exists, else
    normalise -> PCA -> manifold step, cached as TSV (gzip when the name says so).  ``pca_scores``: the PCA
    stage when it has already been computed in a batch (``PartitionPCA``)."""

    def compute():
        if pca_scores is not None:
            # the reference hands embed() the normalised frame; here the PCA stage is already done, so the
            # frame of scores goes in with pca_dimensions = 0 ("no PCA") and only the manifold step runs
            return kmers.embed(kmers=pca_scores, embed_dimensions=embed_dimensions, pca_dimensions=0,
                               method=embed_method)
        return kmers.embed(kmers=kmers.normalize(counts, method=norm_method), embed_dimensions=embed_dimensions,
                           pca_dimensions=pca_dimensions, method=embed_method)

    if not cache_fpath:
        return compute()
    if os.path.exists(cache_fpath) and os.path.getsize(cache_fpath):
        logger.debug(f"Found cached embeddings {cache_fpath}. Reading...")
        return _tsv.read_table(cache_fpath, index_col="contig")
    embedding = compute()
    _tsv.write_frame(embedding, cache_fpath)
    logger.debug(f"Cached embeddings to {cache_fpath}")
    return embedding


def get_checkpoint_info(checkpoints_fpath: str) -> Dict[str, Union[pd.DataFrame, str]]:
    """Checkpoint table + the rank / taxon it stopped at, from the ``# rank:`` / ``# name:`` header lines
    (large_data_mode.py:117-162)."""
    df = pd.read_csv(checkpoints_fpath, sep="\t", index_col="contig", comment="#")
    found = {"rank": None, "name": None}
    patterns = {"rank": re.compile(r"#\srank:\s(\S+)"), "name": re.compile(r"#\sname:\s(\S+)")}
    opener = gzip.open if checkpoints_fpath.endswith(".gz") else open
    with opener(checkpoints_fpath, "rt") as fh:
        for line in fh:
            for key, pat in patterns.items():
                m = pat.search(line)
                if m:
                    found[key] = m.group(1)
            if found["rank"] and found["name"]:
                break
    logger.debug(f"{df.shape[1]:,} binning checkpoints found. starting at {found['rank']} with {found['name']}")
    return {"binning_checkpoints": df, "starting_rank": found["rank"], "starting_rank_name_txt": found["name"]}


def checkpoint(
    checkpoints_df: pd.DataFrame,
    clustered: pd.DataFrame,
    rank: str,
    rank_name_txt: str,
    completeness: float,
    purity: float,
    coverage_stddev: float,
    gc_content_stddev: float,
    cluster_method: str,
    norm_method: str,
    pca_dimensions: int,
    embed_dimensions: int,
    embed_method: str,
    min_contigs: int,
    max_partition_size: int,
    binning_checkpoints_fpath: str,
) -> pd.DataFrame:
    """Add the taxon's bins as a column and rewrite the checkpoint file: ``#`` parameter / runtime header, then
    the table as TSV (large_data_mode.py:165-224)."""
    checkpoints_df = pd.merge(
        checkpoints_df, clustered[["cluster"]], how="outer", left_index=True, right_index=True
    ).rename(columns={"cluster": rank_name_txt})
    now = datetime.now()
    header = [
        "#-- Parameters --#",
        f"# completeness: {completeness}",
        f"# purity: {purity}",
        f"# coverage_stddev: {coverage_stddev}",
        f"# gc_content_stddev: {gc_content_stddev}",
        f"# cluster_method: {cluster_method}",
        f"# norm_method: {norm_method}",
        f"# pca_dimensions: {pca_dimensions}",
        f"# embed_dimensions: {embed_dimensions}",
        f"# embed_method: {embed_method}",
        f"# min-partition-size: {min_contigs} (max([pca_dimensions + 1, embed_dimensions + 1])",
        f"# max-partition-size: {max_partition_size}",
        "#-- Runtime Variables --#",
        f"# rank: {rank}",
        f"# name: {rank_name_txt}",
        f"# checkpoint-shape: {checkpoints_df.shape}",
        f"# current time: {now}",
        f"# timestamp: {now.timestamp()}",
    ]
    text = "\n".join(header + [checkpoints_df.to_csv(sep="\t", index=True, header=True)])
    if binning_checkpoints_fpath.endswith(".gz"):
        with gzip.open(binning_checkpoints_fpath, "wb") as fh:
            fh.write(text.encode())
    else:
        with open(binning_checkpoints_fpath, "w") as fh:
            fh.write(text)
    logger.debug(f"Checkpoint => {rank} : {rank_name_txt} ({checkpoints_df.shape[1]:,} total checkpoints)")
    return checkpoints_df


def _embedding_cache_path(cache, rank_dir, stem, norm_method, pca_dimensions, embed_method, embed_dimensions):
    if not cache:
        return None
    return os.path.join(cache, rank_dir, f"{stem}.{norm_method}_pca{pca_dimensions}_{embed_method}{embed_dimensions}.tsv.gz")


def _needs_compute(path) -> bool:
    return not (path and os.path.exists(path) and os.path.getsize(path))


# ----------------------------------------------------------------------------- the driver
def cluster_by_taxon_partitioning(
    main: pd.DataFrame,
    counts: pd.DataFrame,
    markers: pd.DataFrame,
    norm_method: str = "am_clr",
    pca_dimensions: int = 50,
    embed_dimensions: int = 2,
    embed_method: str = "umap",
    max_partition_size: int = 10000,
    completeness: float = 20.0,
    purity: float = 95.0,
    coverage_stddev: float = 25.0,
    gc_content_stddev: float = 5.0,
    starting_rank: str = "superkingdom",
    method: str = "dbscan",
    reverse_ranks: bool = False,
    cache: str = None,
    binning_checkpoints_fpath: str = None,
    n_jobs: int = -1,
    verbose: bool = False,
    device=None,
) -> pd.DataFrame:
    """Large-data-mode binning (large_data_mode.py:227-594): for every canonical rank from ``starting_rank`` on,
    and every taxon of that rank, embed the taxon's still-unbinned contigs (own embedding when the taxon has
    between ``max(pca_dimensions, embed_dimensions) + 1`` and ``max_partition_size`` contigs; the rank's
    embedding when it is smaller; nothing but a cached embedding when it is larger) and keep the bins
    ``get_clusters`` finds.  Returns ``main`` + ``cluster`` / metric columns, bins named ``bin_NNNN``.

    Raises ``FileNotFoundError`` for a missing checkpoint file given explicitly, ``NotADirectoryError`` when
    ``cache`` is a file, ``BinningError`` when no cluster is recovered at all."""
    if binning_checkpoints_fpath and not os.path.exists(binning_checkpoints_fpath):
        raise FileNotFoundError(binning_checkpoints_fpath)
    ranks = [r for r in CANONICAL_RANKS if r != "root"]
    if not reverse_ranks:
        ranks.reverse()  # superkingdom ... species
    clustered_contigs = set()
    num_clusters = 0
    clusters: List[pd.DataFrame] = []
    resume_name = None
    if cache:
        if not os.path.exists(cache):
            os.makedirs(os.path.realpath(cache), exist_ok=True)
            logger.debug(f"Created cache dir: {cache}")
        if not os.path.isdir(cache):
            raise NotADirectoryError(cache)
        if not binning_checkpoints_fpath:
            binning_checkpoints_fpath = os.path.join(cache, "binning_checkpoints.tsv.gz")
    if not _needs_compute(binning_checkpoints_fpath):
        info = get_checkpoint_info(binning_checkpoints_fpath)
        binning_checkpoints = info["binning_checkpoints"]
        starting_rank, resume_name = info["starting_rank"], info["starting_rank_name_txt"]
        # the most recent bin of every contig = last non-null entry along the checkpoint columns
        latest = binning_checkpoints.ffill(axis=1).iloc[:, -1].dropna()
        clustered_contigs = set(latest.index.unique().tolist())
        latest_df = latest.to_frame().rename(columns={resume_name: "cluster"})
        num_clusters = latest_df.cluster.nunique()
        clusters.append(latest_df)
    else:
        logger.debug(
            f"Binning checkpoints not found. Writing checkpoints to {binning_checkpoints_fpath}"
            if binning_checkpoints_fpath else "Binning checkpoints not found. Initializing..."
        )
        binning_checkpoints = pd.DataFrame()
    ranks = ranks[ranks.index(starting_rank):]
    logger.debug(f"Using canonical ranks: {', '.join(ranks)}")
    logger.debug(f"Max partition size set to: {max_partition_size}")
    min_contigs = max([pca_dimensions + 1, embed_dimensions + 1])
    batch = PartitionPCA(counts, norm_method, pca_dimensions, device=device)
    n_cols = counts.shape[1]
    resumed = False
    cache_args = (norm_method, pca_dimensions, embed_method, embed_dimensions)
    embed_args = dict(norm_method=norm_method, pca_dimensions=pca_dimensions, embed_dimensions=embed_dimensions,
                      embed_method=embed_method)
    for pos, rank in enumerate(ranks):
        classified = main[rank] != "unclassified"
        logger.info(
            f"Examining {rank}: {main.loc[classified, rank].nunique():,} unique taxa "
            f"({main.loc[classified, rank].shape[0]:,} contigs)"
        )
        # ---- plan: which embeddings does this rank need?  (all known before the per-taxon loop)
        rank_counts_index = counts.index[counts.index.isin(main.loc[classified].index)]
        plan: Dict[str, pd.Index] = {}
        rank_source = rank  # the rank whose embedding small taxa fall back on
        if len(rank_counts_index) >= min_contigs:
            rank_cache = _embedding_cache_path(cache, rank, rank, *cache_args)
            if cache:
                os.makedirs(os.path.join(cache, rank), exist_ok=True)
        else:
            # too few classified contigs at this rank: the previous rank's embedding (for the first rank the
            # reference's index -1 wraps to the last one, large_data_mode.py:436-440; mirrored)
            rank_source = ranks[pos - 1]
            prev_classified = main[rank_source].ne("unclassified")
            rank_counts_index = counts.index[counts.index.isin(main.loc[prev_classified].index)]
            rank_cache = _embedding_cache_path(cache, rank_source, rank_source, *cache_args)
        if _needs_compute(rank_cache) and _pca_applies(len(rank_counts_index), n_cols, norm_method, pca_dimensions):
            plan["__rank__"] = rank_counts_index
        taxa = []
        skipping = bool(resume_name) and not resumed
        for name, dff in main.loc[classified].groupby(rank):
            if name == resume_name:
                skipping = False
            if skipping or dff.empty:
                taxa.append((name, None, None))
                continue
            rank_df = dff.loc[dff["cluster"].isna()].copy() if "cluster" in dff.columns else dff.copy()
            if clustered_contigs:
                rank_df = rank_df.loc[~rank_df.index.isin(clustered_contigs)]
            if rank_df.empty:
                taxa.append((name, None, None))
                continue
            taxon_index = counts.index[counts.index.isin(rank_df.index)]
            taxon_cache = _embedding_cache_path(cache, rank, name, *cache_args)
            n = rank_df.shape[0]
            if n >= min_contigs and _needs_compute(taxon_cache) and _pca_applies(len(taxon_index), n_cols, norm_method,
                                                                                pca_dimensions):
                plan[name] = taxon_index
            taxa.append((name, rank_df, taxon_cache))
        # ---- one batch on the device: normalise + PCA of every partition of this rank
        scores = batch.run(plan) if plan else {}
        rank_embedding = get_kmer_embedding(counts.loc[rank_counts_index], cache_fpath=rank_cache,
                                            pca_scores=scores.get("__rank__"), **embed_args)
        # ---- per taxon: manifold step (host, in the reference's order) ...
        jobs = []  # (name, frame to sweep | None for a taxon that is only cached)
        for name, rank_df, taxon_cache in taxa:
            if name == resume_name:
                resumed = True
            if resume_name and not resumed:
                continue
            if rank_df is None:
                continue
            # contigs binned earlier IN THIS RANK belong to other taxa, so rank_df planned above is still exact
            n = rank_df.shape[0]
            taxon_counts = counts.loc[counts.index.isin(rank_df.index)]
            if n > max_partition_size:
                logger.debug(f"{name} > max_partition_size ({n:,}>{max_partition_size:,}). skipping [and caching embedding]")
                get_kmer_embedding(taxon_counts, cache_fpath=taxon_cache, pca_scores=scores.get(name), **embed_args)
                jobs.append((name, None))
                continue
            if n >= min_contigs:
                taxon_embedding = get_kmer_embedding(taxon_counts, cache_fpath=taxon_cache, pca_scores=scores.get(name),
                                                     **embed_args)
                logger.debug(f"{name} shape={taxon_embedding.shape}")
            else:
                taxon_embedding = rank_embedding.loc[rank_embedding.index.isin(rank_df.index)]
                logger.debug(f"using {rank} embedding for {name}. shape={taxon_embedding.shape}")
            rank_df = pd.merge(rank_df, taxon_embedding, how="left", left_index=True, right_index=True)
            logger.debug(f"Examining taxonomy: {rank} : {name} : {rank_df.shape}")
            jobs.append((name, rank_df))
        # ---- ... then the sweeps of all the rank's taxa in flight together on the device (the taxa of a rank
        # are disjoint, so no sweep changes another's input) ...
        binned_frames = iter(get_clusters_many(
            [df for _, df in jobs if df is not None], markers_df=markers, completeness=completeness, purity=purity,
            coverage_stddev=coverage_stddev, gc_content_stddev=gc_content_stddev, method=method, n_jobs=n_jobs,
            verbose=verbose, device=device,
        ))
        # ---- ... and the bookkeeping (bin names, checkpoints) taxon by taxon in the reference's order
        for name, rank_df in jobs:
            if rank_df is None:
                binning_checkpoints[name] = pd.NA
                continue
            binned = next(binned_frames)
            clustered = binned.loc[binned["cluster"].notnull()]
            if clustered.empty:
                continue
            clustered_contigs.update(clustered.index)
            clustered = reindex_bin_names(df=clustered, initial_index=num_clusters)
            num_clusters += clustered.cluster.nunique()
            clusters.append(clustered)
            if binning_checkpoints_fpath:
                binning_checkpoints = checkpoint(
                    checkpoints_df=binning_checkpoints, clustered=clustered, rank=rank, rank_name_txt=name,
                    completeness=completeness, purity=purity, coverage_stddev=coverage_stddev,
                    gc_content_stddev=gc_content_stddev, cluster_method=method, norm_method=norm_method,
                    pca_dimensions=pca_dimensions, embed_dimensions=embed_dimensions, embed_method=embed_method,
                    min_contigs=min_contigs, max_partition_size=max_partition_size,
                    binning_checkpoints_fpath=binning_checkpoints_fpath,
                )
    if not clusters:
        raise BinningError("Failed to recover any clusters from dataset")
    clustered_df = pd.concat(clusters, sort=True)
    unclustered_df = main.loc[~main.index.isin(clustered_df.index)].copy()
    unclustered_df["cluster"] = pd.NA
    return zero_pad_bin_names(pd.concat([clustered_df, unclustered_df], sort=True))
