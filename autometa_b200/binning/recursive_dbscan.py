"""Drop-in for the DBSCAN eps step and sweep of ``autometa.binning.recursive_dbscan``.

``run_dbscan`` (reference recursive_dbscan.py:47-111) and ``recursive_dbscan``
(:114-234) keep their signatures, in-place column drop, cluster numbering,
schedule, tie rule and termination; the clustering itself runs on the GPU
(``am_eps_union`` / ``am_labels_from_parent``).  The reference uses
``DBSCAN(min_samples=1)``: every point is a core point, there are no border or
noise points, and each eps is the connected components of the eps-graph — so the
sweep keeps ONE union-find forest on the device and only adds merges as eps grows.
The sweep's loop state (eps, step, best median so far, ``clustering_rounds``, termination)
lives on the device too (``device.DeviceSweep``, csrc/am_sweep.cu): one step is a CUDA
graph, the host queues batches of replays and reads one word per batch.

The callers of the sweep are mirrored on top of it: ``get_clusters``
(:440-545, re-sweeps the leftovers until nothing more is recovered) and
``taxon_guided_binning`` (:548-705, the same per taxon from superkingdom down to
species); ``get_clusters_many`` runs the independent ``get_clusters`` problems of one rank's
taxa with their sweeps in flight together.  Only ``method="dbscan"`` is on the device; HDBSCAN is out of scope
(``NotImplementedError``).  The eps step takes at most 4 clustering features (the
columns containing ``x_`` plus ``coverage``: ``embed_dimensions`` <= 3), more raise
``ValueError``; the reference's default is 2 + coverage.
"""
from __future__ import annotations

import logging
from typing import List, Tuple

import numpy as np
import pandas as pd

from .. import _lib as L
from .. import device as _dev
from ..common.exceptions import BinningError, TableFormatError
from .utilities import (
    METRICS,
    add_metrics,
    apply_binning_metrics_filter,
    markers_coo,
    passing_mask,
    reindex_bin_names,
    zero_pad_bin_names,
)

#: autometa/taxonomy/database.py:65-74
CANONICAL_RANKS = ["species", "genus", "family", "order", "class", "phylum", "superkingdom", "root"]

logger = logging.getLogger(__name__)

_DROPCOLS = ["cluster", "purity", "completeness", "coverage_stddev", "gc_content_stddev"]


def _feature_matrix(df: pd.DataFrame) -> np.ndarray:
    # recursive_dbscan.py:105-107: every column containing "x_" plus "coverage", in frame order
    cols = [col for col in df.columns if "x_" in col or col == "coverage"]
    if not cols or not df.columns.is_unique:
        return np.ascontiguousarray(df.loc[:, cols].to_numpy(dtype=np.float64))
    return np.ascontiguousarray(np.stack([df[col].to_numpy(dtype=np.float64) for col in cols], axis=1))


def _upload(X: np.ndarray, device=None):
    import torch

    dev = L.require_cuda(device)
    if X.shape[1] < 1 or X.shape[1] > 4:
        raise ValueError(f"the GPU eps step supports 1..4 clustering features, got {X.shape[1]}")
    return torch.from_numpy(X).to(dev)


def run_dbscan(
    df: pd.DataFrame, eps: float, n_jobs: int = -1, dropcols: List[str] = _DROPCOLS, device=None
) -> pd.DataFrame:
    """Cluster ``df`` at ``eps`` (reference recursive_dbscan.py:47-111).

    Drops ``dropcols`` from the caller's frame in place, raises
    ``TableFormatError`` on nulls, returns ``df`` + int64 ``cluster`` column
    (``pd.NA`` when there is a single contig).  ``n_jobs`` is accepted and ignored.
    """
    df.drop(columns=dropcols, inplace=True, errors="ignore")
    n_samples = df.shape[0]
    if n_samples == 1:
        clusters = pd.Series([pd.NA], index=df.index, name="cluster")
        return pd.merge(df, clusters, how="left", left_index=True, right_index=True)
    if np.any(df.isnull()):
        raise TableFormatError(f"df is missing {df.isnull().sum().sum()} kmer/coverage annotations")
    X = _feature_matrix(df)
    if n_samples == 0:
        labels = np.zeros(0, dtype=np.int64)
    else:
        sweep = _dev.EpsSweep(_upload(X, device))
        labels = sweep.step(float(eps)).cpu().numpy()
    clusters = pd.Series(labels, index=df.index, name="cluster")
    return pd.merge(df, clusters, how="left", left_index=True, right_index=True)


#: how a sweep is driven: "device" (default) = the loop state lives on the GPU (``device.DeviceSweep``,
#: csrc/am_sweep.cu), the host queues batches of graph replays; "host" = one synchronisation per eps with the
#: schedule in Python (the round-1 loop, kept as a cross-check of the device controller — both are CUDA paths)
SWEEP_MODES = ("device", "host")

#: device memory the sweeps in flight may hold together (bytes), and the most sweeps in flight
_IN_FLIGHT_BYTES = 4 << 30
_IN_FLIGHT_MAX = 16

_streams = {}


def _stream_pool(dev, k: int):
    """``k`` streams of the device's pool (created on first use): the sweeps in flight get one each."""
    import torch

    pool = _streams.setdefault(dev.index, [])
    while len(pool) < k:
        pool.append(torch.cuda.Stream(device=dev))
    return pool[:k]


def _drive(starters, limit: int = 1):
    """Run generator-style jobs to completion, at most ``limit`` in flight.  ``starters[i](slot)`` returns a
    generator that yields an object with ``ready()`` / ``wait()`` each time it has queued device work and needs
    it finished before it can go on; its return value is the job's result.  With one job this is a plain loop;
    with several, the host-side work of one (pandas frames, queueing) overlaps the device work of the others."""
    results = [None] * len(starters)
    live = {}  # slot -> (job index, generator, what it waits for)
    nxt = 0
    free = list(range(max(1, limit)))[::-1]

    def advance(slot, i, gen):
        try:
            live[slot] = (i, gen, next(gen))
        except StopIteration as stop:
            results[i] = stop.value
            live.pop(slot, None)
            free.append(slot)

    while nxt < len(starters) or live:
        while free and nxt < len(starters):
            slot = free.pop()
            advance(slot, nxt, starters[nxt](slot))
            nxt += 1
        progressed = False
        for slot in list(live):
            i, gen, w = live[slot]
            if w is None or w.ready():
                advance(slot, i, gen)
                progressed = True
        if not progressed and live:
            w = next(iter(live.values()))[2]
            if w is not None:
                w.wait()
    return results


def _in_flight(n_rows, n_markers: int) -> int:
    """How many sweeps of up to ``n_rows`` points may be in flight within the memory budget."""
    per = max(1, int(n_rows)) * (4 * max(1, int(n_markers)) + 256)
    return int(max(1, min(_IN_FLIGHT_MAX, _IN_FLIGHT_BYTES // per)))


def _sweep_mode() -> str:
    import os

    mode = os.environ.get("AUTOMETA_B200_SWEEP", "device")
    if mode not in SWEEP_MODES:
        raise ValueError(f"AUTOMETA_B200_SWEEP must be one of {SWEEP_MODES}, got {mode!r}")
    return mode


def recursive_dbscan(
    table: pd.DataFrame,
    markers_df: pd.DataFrame,
    completeness_cutoff: float,
    purity_cutoff: float,
    coverage_stddev_cutoff: float,
    gc_content_stddev_cutoff: float,
    n_jobs: int = -1,
    verbose: bool = False,
    device=None,
) -> Tuple[pd.DataFrame, pd.DataFrame]:
    """The eps sweep (reference recursive_dbscan.py:114-234): eps = 0.3, += step (0.1,
    x10 each time the current cluster count has been seen > 10 times), keep the
    last eps whose median completeness of passing clusters is >= the best so far,
    stop at <= 1 cluster (or median == 0 ten times).  Returns
    ``(clustered_df, unclustered_df)``.
    """
    cut = (completeness_cutoff, purity_cutoff, coverage_stddev_cutoff, gc_content_stddev_cutoff)
    return _drive([lambda slot: _recursive_dbscan_gen(table, markers_df, cut, verbose, device, None)])[0]


def _recursive_dbscan_gen(table, markers_df, cut, verbose, device, stream):
    """``recursive_dbscan`` as a generator (see ``_drive``): yields each time a batch of eps steps has been
    queued on the device."""
    n_samples = table.shape[0]
    if n_samples <= 1 or "coverage" not in table.columns or "gc_content" not in table.columns:
        return _recursive_dbscan_frames(table, markers_df, *cut, verbose, device)
    # run_dbscan's side effects and checks (:95-103), once
    table.drop(columns=_DROPCOLS, inplace=True, errors="ignore")
    if np.any(table.isnull()):
        raise TableFormatError(f"df is missing {table.isnull().sum().sum()} kmer/coverage annotations")
    X = _feature_matrix(table)
    if X.shape[1] < 1 or X.shape[1] > 4:
        raise ValueError(f"the GPU eps step supports 1..4 clustering features, got {X.shape[1]}")
    dev = L.require_cuda(device)
    coverage = table["coverage"].to_numpy(dtype=np.float64)
    gc = table["gc_content"].to_numpy(dtype=np.float64)
    rows, cols, vals = markers_coo(markers_df, table.index)
    n_ref = markers_df.shape[1]
    if _sweep_mode() == "host":
        labels, m, best_median = _sweep_hostloop(X, coverage, gc, rows, cols, vals, n_ref, cut, dev, verbose)
    else:
        # The loop (schedule, best-eps rule, termination :189-215) runs on the device; the host queues
        # batches of steps and looks at one word per batch.
        if stream is None:
            stream = _stream_pool(dev, 1)[0]
        run = _dev.DeviceSweep(X, coverage, gc, rows, cols, vals, n_ref, cut, dev, stream=stream,
                               trace_cap=4096 if verbose else 0)
        while True:
            run.launch()
            yield run
            if run.poll():
                break
        labels, m, best_median, _, trace = run.result()
        run.close()
        if verbose and trace is not None:
            zero_rounds = 0
            for k, (eps, n_clusters, median, best) in enumerate(trace):
                zero_rounds += int(median == 0)
                if zero_rounds >= 10:  # the reference leaves the loop before it logs this eps (:208-214)
                    break
                logger.debug(
                    f"EPS: {eps:3.2f} Clusters: {int(n_clusters):,}"
                    f" Completeness: Median={median:4.2f} Best={best:4.2f}"
                )
    # the frame add_metrics returns (utilities.py:173-206): the table without earlier metric columns, then
    # ``cluster`` and the six metrics; built in one concat, and split by the passing mask of the CLUSTERS
    # (= apply_binning_metrics_filter on the rows, :257-262, and its complement by index)
    stale = [c for c in METRICS if c in table.columns]
    base = table.drop(columns=stale) if stale else table
    extra = pd.DataFrame({"cluster": labels, **{name: m[name][labels] for name in METRICS}}, index=table.index)
    best_df = pd.concat([base, extra], axis=1)
    if table.index.is_unique:
        passing = passing_mask(m, *cut)[labels]
        clustered_df = best_df[passing]
        unclustered_df = best_df[~passing]
    else:
        clustered_df = apply_binning_metrics_filter(
            df=best_df,
            completeness_cutoff=cut[0],
            purity_cutoff=cut[1],
            coverage_stddev_cutoff=cut[2],
            gc_content_stddev_cutoff=cut[3],
        )
        unclustered_df = best_df.loc[~best_df.index.isin(clustered_df.index)]
    if verbose:
        logger.debug(f"Best completeness median: {best_median:4.2f}")
    logger.debug(f"clustered: {clustered_df.shape[0]:,} unclustered: {unclustered_df.shape[0]:,}")
    return clustered_df, unclustered_df


def _sweep_hostloop(X, coverage, gc, rows, cols, vals, n_ref, cut, dev, verbose):
    """The sweep with the schedule on the host: labels and metrics stay on the device, per eps only the
    3-number summary (cluster count, median completeness of passing clusters, number passing) comes back."""
    import torch

    n_samples = X.shape[0]
    sweep = _dev.EpsSweep(_upload(X, dev))
    metrics = _dev.ClusterMetrics(n_samples, rows, cols, vals, n_ref, coverage, gc, cut, dev)
    eps, step = 0.3, 0.1
    clustering_rounds = {0: 0}
    n_clusters = float("inf")
    best_median = float("-inf")
    best_labels_d = torch.empty(n_samples, dtype=torch.int64, device=dev)
    best_keep = None
    while n_clusters > 1:
        labels_d = sweep.step(eps)
        summ = metrics.compute(labels_d, sweep.n_clusters).cpu().numpy()  # the only sync per eps
        n_clusters = int(summ[0])
        median_completeness = float(summ[1])
        if median_completeness >= best_median:
            best_median = median_completeness
            best_labels_d.copy_(labels_d)
            best_keep = {k2: v.clone() for k2, v in dict(
                completeness=metrics.completeness[:n_clusters], purity=metrics.purity[:n_clusters],
                coverage_stddev=metrics.cov_std[:n_clusters], gc_content_stddev=metrics.gc_std[:n_clusters],
                present_marker_count=metrics.present[:n_clusters], single_copy_marker_count=metrics.single[:n_clusters],
            ).items()}
        clustering_rounds[n_clusters] = clustering_rounds.get(n_clusters, 0) + 1
        if clustering_rounds[n_clusters] > 10:
            step *= 10
        if median_completeness == 0:
            clustering_rounds[0] += 1
        if clustering_rounds[0] >= 10:
            break
        if verbose:
            logger.debug(
                f"EPS: {eps:3.2f} Clusters: {n_clusters:,}"
                f" Completeness: Median={median_completeness:4.2f} Best={best_median:4.2f}"
            )
        eps += step
    m = {k2: v.cpu().numpy() for k2, v in best_keep.items()}
    for k2 in ("present_marker_count", "single_copy_marker_count"):
        m[k2] = m[k2].astype(np.int64)
    return best_labels_d.cpu().numpy(), m, best_median


def _recursive_dbscan_frames(table, markers_df, completeness_cutoff, purity_cutoff,
                             coverage_stddev_cutoff, gc_content_stddev_cutoff, verbose, device):
    """Literal frame-by-frame form of the loop for degenerate inputs (a single
    contig, missing columns): same control flow as recursive_dbscan.py:172-234."""
    eps, step = 0.3, 0.1
    clustering_rounds = {0: 0}
    n_clusters = float("inf")
    best_median = float("-inf")
    best_df = pd.DataFrame()
    while n_clusters > 1:
        binned_df = run_dbscan(df=table, eps=eps, device=device)
        df, metrics_df = add_metrics(df=binned_df, markers_df=markers_df)
        filtered_df = apply_binning_metrics_filter(
            metrics_df, completeness_cutoff, purity_cutoff, coverage_stddev_cutoff, gc_content_stddev_cutoff
        )
        median_completeness = float("-inf") if filtered_df.empty else filtered_df.completeness.median()
        if median_completeness >= best_median:
            best_median = median_completeness
            best_df = df
        n_clusters = df["cluster"].nunique()
        clustering_rounds[n_clusters] = clustering_rounds.get(n_clusters, 0) + 1
        if clustering_rounds[n_clusters] > 10:
            step *= 10
        if median_completeness == 0:
            clustering_rounds[0] += 1
        if clustering_rounds[0] >= 10:
            break
        eps += step
    if best_df.empty:
        return pd.DataFrame(), table
    clustered_df = apply_binning_metrics_filter(
        best_df, completeness_cutoff, purity_cutoff, coverage_stddev_cutoff, gc_content_stddev_cutoff
    )
    unclustered_df = best_df.loc[~best_df.index.isin(clustered_df.index)]
    return clustered_df, unclustered_df


def get_clusters(
    main: pd.DataFrame,
    markers_df: pd.DataFrame,
    completeness: float,
    purity: float,
    coverage_stddev: float,
    gc_content_stddev: float,
    method: str,
    n_jobs: int = -1,
    verbose: bool = False,
    device=None,
) -> pd.DataFrame:
    """Sweep, keep the passing clusters, sweep the leftovers again ... until a round recovers
    nothing or nothing is left (reference recursive_dbscan.py:440-545).  Returns ``main`` with
    ``cluster`` (``bin_NN`` / ``pd.NA``) and the four metric columns."""
    return get_clusters_many([main], markers_df, completeness, purity, coverage_stddev, gc_content_stddev, method,
                             n_jobs=n_jobs, verbose=verbose, device=device)[0]


def get_clusters_many(
    mains: List[pd.DataFrame],
    markers_df: pd.DataFrame,
    completeness: float,
    purity: float,
    coverage_stddev: float,
    gc_content_stddev: float,
    method: str,
    n_jobs: int = -1,
    verbose: bool = False,
    device=None,
) -> List[pd.DataFrame]:
    """``get_clusters`` of several independent contig tables (the taxa of one rank: SURVEY 8f-3), their
    sweeps in flight together on the device — each on its own stream, each a chain of CUDA-graph replays
    (``device.DeviceSweep``), so the small sweeps of large-data mode are bound neither by one host
    synchronisation per eps nor by one sweep's launch latency.  Results are those of ``get_clusters`` on each
    table, in order."""
    if method not in ("dbscan", "hdbscan"):
        raise ValueError(f"Method: {method} not a choice. choose between dict_keys(['dbscan', 'hdbscan'])")
    if method != "dbscan":
        raise NotImplementedError("only method='dbscan' runs on the device; HDBSCAN is out of scope (DESIGN.md 9)")
    if not mains:
        return []
    cut = (completeness, purity, coverage_stddev, gc_content_stddev)
    limit = 1
    pool = [None]
    if _sweep_mode() == "device":
        dev = L.require_cuda(device)
        if len(mains) > 1:
            limit = min(len(mains), _in_flight(max(m.shape[0] for m in mains), markers_df.shape[1]))
        pool = _stream_pool(dev, limit)
    starters = [
        (lambda slot, main=main: _get_clusters_gen(main, markers_df, cut, verbose, device, pool[slot]))
        for main in mains
    ]
    return _drive(starters, limit)


def _get_clusters_gen(main, markers_df, cut, verbose, device, stream):
    n_bins = 0
    rounds = []
    while True:
        clustered_df, unclustered_df = yield from _recursive_dbscan_gen(main, markers_df, cut, verbose, device,
                                                                        stream)
        if clustered_df.empty:  # nothing recovered: the rest stays unbinned
            rounds.append(unclustered_df.assign(cluster=pd.NA))
            break
        clustered_df = reindex_bin_names(df=clustered_df, initial_index=n_bins)
        rounds.append(clustered_df)
        if unclustered_df.empty:
            break
        n_bins += clustered_df.cluster.nunique()
        main = unclustered_df
    df = pd.concat(rounds, axis="index", sort=True)
    for metric in ["purity", "completeness", "coverage_stddev", "gc_content_stddev"]:
        if metric not in df.columns:  # no cluster at all in the first round
            df[metric] = pd.NA
    return zero_pad_bin_names(df)


def taxon_guided_binning(
    main: pd.DataFrame,
    markers: pd.DataFrame,
    completeness: float = 20.0,
    purity: float = 95.0,
    coverage_stddev: float = 25.0,
    gc_content_stddev: float = 5.0,
    starting_rank: str = "superkingdom",
    method: str = "dbscan",
    reverse_ranks: bool = False,
    n_jobs: int = -1,
    verbose: bool = False,
    device=None,
) -> pd.DataFrame:
    """``get_clusters`` per taxon, rank by rank, on the contigs not binned yet
    (reference recursive_dbscan.py:548-705).  ``main`` carries the canonical-rank columns."""
    if main.loc[main.index.isin(markers.index)].empty:
        raise TableFormatError("No markers for contigs in table. Unable to assess binning quality")
    if main.shape[0] <= 1:
        raise BinningError("Not enough contigs in table for binning")
    logger.info(f"Using {method} clustering method")
    ranks = [r for r in (CANONICAL_RANKS if reverse_ranks else reversed(CANONICAL_RANKS)) if r != "root"]
    ranks = ranks[ranks.index(starting_rank):]
    logger.debug(f"Using ranks: {', '.join(ranks)}")
    binned = set()
    n_bins = 0
    found = []
    for rank in ranks:
        by_taxon = main.loc[main[rank] != "unclassified"].groupby(rank)
        sizes = by_taxon[rank].count()
        logger.info(f"Examining {rank}: {sizes.index.nunique():,} unique taxa ({sizes.sum():,} contigs)")
        # a contig has one taxon per rank, so the taxa of a rank are independent problems (bins found in
        # one never remove contigs from another): their sweeps are queued together (get_clusters_many)
        todos = []
        for taxon, members in by_taxon:
            if members.empty:
                continue
            todo = members.loc[members["cluster"].isna()] if "cluster" in members.columns else members
            if binned:
                todo = todo.loc[~todo.index.isin(binned)]
            if todo.empty:
                continue
            logger.debug(f"Examining taxonomy: {rank} : {taxon} : {todo.shape}")
            todos.append(todo)
        results = get_clusters_many(
            todos, markers_df=markers, completeness=completeness, purity=purity, coverage_stddev=coverage_stddev,
            gc_content_stddev=gc_content_stddev, method=method, n_jobs=n_jobs, verbose=verbose, device=device,
        )
        for clusters_df in results:
            clustered = clusters_df.loc[clusters_df["cluster"].notnull()]
            if clustered.empty:
                continue
            binned.update(clustered.index)
            names = {c: f"bin_{1 + i + n_bins:04d}" for i, c in enumerate(clustered.cluster.unique())}
            clustered.cluster = clustered.cluster.map(names.__getitem__)
            n_bins += clustered.cluster.nunique()
            found.append(clustered)
    if not found:
        raise BinningError("Failed to recover any clusters from dataset")
    clustered_df = pd.concat(found, sort=True)
    unclustered_df = main.loc[~main.index.isin(clustered_df.index)]
    unclustered_df["cluster"] = pd.NA
    return pd.concat([clustered_df, unclustered_df], sort=True)
