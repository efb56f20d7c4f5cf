"""Tensor-level wrappers over the C ABI: every function takes/returns torch CUDA
tensors, launches on torch's current stream and never synchronises unless it
has to read a scalar back.  These are what the reference-shaped functions in
``common/kmers.py`` and ``binning/recursive_dbscan.py`` (and ``bench.py``) call.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

from . import _lib as L
from .assembly import PackedAssembly, count_plan

_num_columns_cache: Dict[int, int] = {}


def num_columns(k: int) -> int:
    if k not in _num_columns_cache:
        n = L.lib.am_kmer_num_columns(int(k))
        if n < 0:
            L.check(n)
        _num_columns_cache[k] = n
    return _num_columns_cache[k]


def column_names(k: int):
    n = num_columns(k)
    buf = np.zeros(n * (k + 1), dtype=np.uint8)
    L.check(L.lib.am_kmer_column_names(k, L.np_ptr(buf)))
    raw = buf.tobytes()
    return [raw[i * (k + 1) : i * (k + 1) + k].decode() for i in range(n)]


def column_lut(k: int) -> np.ndarray:
    lut = np.zeros(4 ** k, dtype=np.int32)
    L.check(L.lib.am_kmer_column_lut(k, L.np_ptr(lut)))
    return lut


class CountPlan:
    """Device copy of the count work list for one (assembly, k)."""

    def __init__(self, packed: PackedAssembly, k: int, max_item_bases: int = 32768):
        import torch

        items, multi = count_plan(packed.lengths, k, max_item_bases)
        if len(items) > 1:
            # longest work items first: warps claim items from an atomic counter, so the tail of the kernel is
            # one item long — 32 kbp claimed last costs a 25k-contig shard a third of its run time
            items = np.ascontiguousarray(items[np.argsort(-items[:, 2], kind="stable")])
        dev = packed.device
        self.n_items = len(items)
        self.n_multi = len(multi)
        self.items = torch.from_numpy(items if len(items) else np.zeros((1, 4), np.int32)).to(dev)
        self.multi = torch.from_numpy(multi if len(multi) else np.zeros(1, np.int32)).to(dev)
        self.work = torch.zeros(64, dtype=torch.int32, device=dev)


def get_count_plan(packed: PackedAssembly, k: int) -> CountPlan:
    if k not in packed._plans:
        packed._plans[k] = CountPlan(packed, k)
    return packed._plans[k]


def count_packed(packed: PackedAssembly, k: int = 5, counts=None, col_present=None, row_sum=None):
    """Kernel 1. Returns (counts[int32 N x ncols], col_present[int32 ncols], row_sum[int32 N])."""
    import torch

    dev = packed.device
    ncols = num_columns(k)
    n = packed.n_contigs
    plan = get_count_plan(packed, k)
    with L.on_device(dev):
        if counts is None:
            counts = torch.empty((n, ncols), dtype=torch.int32, device=dev)
        if col_present is None:
            col_present = torch.zeros(ncols, dtype=torch.int32, device=dev)
        else:
            col_present.zero_()
        if row_sum is None:
            row_sum = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        L.check(
            L.lib.am_count_kmers(
                L.t_ptr(packed.codes), L.t_ptr(packed.exc_bitmap), L.t_ptr(packed.exc_rank),
                L.t_ptr(packed.exc_mask), L.t_ptr(packed.starts), L.t_ptr(plan.items), plan.n_items,
                L.t_ptr(plan.multi), plan.n_multi, n, k, L.t_ptr(counts), L.t_ptr(col_present),
                L.t_ptr(row_sum), L.t_ptr(plan.work), L.stream_ptr(dev),
            )
        )
    return counts, col_present, row_sum[:n]


_work_cache: Dict[Tuple[str, int, int], "torch.Tensor"] = {}


def _work(dev, nbytes: int, tag: str):
    import torch

    # scratch is per (kernel family, device, stream): passes running on different streams (pipeline.FeatureStream,
    # per-partition streams of the large-data-mode driver) must not share partial-result buffers
    key = (tag, dev.index, L.stream_ptr(dev))
    t = _work_cache.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=dev)
        _work_cache[key] = t
    return t


def normalize_counts(counts, method: str = "am_clr", row_index=None, col_index=None, out=None,
                     colsum=None, colmaxabs=None):
    """Kernel 2. ``counts`` int32 [N x D] on device; optional gathers fuse am_clr's
    row/column drops.  Returns float64 [N_out x D_out]."""
    import torch

    dev = counts.device
    if method not in L.NORM_METHODS:
        raise ValueError(f"Normalize Method not available! {method}. choices: ilr, clr, am_clr")
    m = L.NORM_METHODS[method]
    n_in, D = counts.shape
    n_out = n_in if row_index is None else int(row_index.numel())
    Dk = D if col_index is None else int(col_index.numel())
    Dout = Dk - 1 if method == "ilr" else Dk
    with L.on_device(dev):
        if out is None:
            out = torch.empty((n_out, Dout), dtype=torch.float64, device=dev)
        work = None
        if colsum is not None:
            work = _work(dev, L.lib.am_normalize_work_bytes(D), "norm")
        L.check(
            L.lib.am_normalize(
                L.t_ptr(counts), n_out, D, L.t_ptr(row_index), L.t_ptr(col_index), Dk, m,
                L.t_ptr(out), L.t_ptr(colsum), L.t_ptr(colmaxabs), L.t_ptr(work), L.stream_ptr(dev),
            )
        )
    return out


def colsum(X, out=None):
    import torch

    dev = X.device
    n, D = X.shape
    with L.on_device(dev):
        if out is None:
            out = torch.zeros(D, dtype=torch.float64, device=dev)
        work = _work(dev, L.lib.am_colsum_work_bytes(D), "colsum")
        L.check(L.lib.am_colsum(L.t_ptr(X), n, D, L.t_ptr(out), L.t_ptr(work), L.stream_ptr(dev)))
    return out


def gram_centered(X, mu, out=None):
    """G += (X - mu)^T (X - mu) (fp64).  ``out`` must be zeroed by the caller when given."""
    import torch

    dev = X.device
    n, D = X.shape
    with L.on_device(dev):
        if out is None:
            out = torch.zeros((D, D), dtype=torch.float64, device=dev)
        work = _work(dev, L.lib.am_gram_work_bytes(D), "gram")
        L.check(L.lib.am_gram(L.t_ptr(X), n, D, L.t_ptr(mu), L.t_ptr(out), L.t_ptr(work), L.stream_ptr(dev)))
    return out


def colstats(X, colsum_out=None, colmax_out=None):
    """Column sums and column max|x| of ``X`` in one pass (both accumulate into the outputs)."""
    import torch

    dev = X.device
    n, D = X.shape
    with L.on_device(dev):
        if colsum_out is None:
            colsum_out = torch.zeros(D, dtype=torch.float64, device=dev)
        if colmax_out is None:
            colmax_out = torch.zeros(D, dtype=torch.float64, device=dev)
        work = _work(dev, L.lib.am_colstats_work_bytes(D), "colstats")
        L.check(L.lib.am_colstats(L.t_ptr(X), n, D, L.t_ptr(colsum_out), L.t_ptr(colmax_out), L.t_ptr(work),
                                  L.stream_ptr(dev)))
    return colsum_out, colmax_out


GRAM_SLICES = 6


def gram_centered_i8(X, mu, colmaxabs, out=None, slices: int = GRAM_SLICES):
    """G += (X - mu)^T (X - mu) on the tensor cores (tcgen05 kind::i8 over ``slices`` exact
    base-256 digit planes; see csrc/am_gram_i8.cu).  ``colmaxabs[j] >= max_i |X[i, j]|``."""
    import torch

    dev = X.device
    n, D = X.shape
    with L.on_device(dev):
        if out is None:
            out = torch.zeros((D, D), dtype=torch.float64, device=dev)
        work = _work(dev, L.lib.am_gram_i8_work_bytes(n, D, slices), "gram_i8")
        L.check(L.lib.am_gram_i8(L.t_ptr(X), n, D, L.t_ptr(mu), L.t_ptr(colmaxabs), slices, L.t_ptr(out),
                                 L.t_ptr(work), L.stream_ptr(dev)))
    return out


def plane_scales(center, colmaxabs=None, bound: float = -1.0, slices: int = GRAM_SLICES):
    """Fixed-point scale of the digit planes: ``e[j]`` with ``2^e[j] > max|x_j| + |center_j|`` (per-column
    ``colmaxabs`` or one scalar ``bound``) and ``sc[j] = 2^(8S-2-e[j])``."""
    import torch

    dev = center.device
    D = center.numel()
    with L.on_device(dev):
        e = torch.empty(D, dtype=torch.int32, device=dev)
        sc = torch.empty(D, dtype=torch.float64, device=dev)
        L.check(L.lib.am_plane_scales(L.t_ptr(center), L.t_ptr(colmaxabs), float(bound), D, slices, L.t_ptr(e),
                                      L.t_ptr(sc), L.stream_ptr(dev)))
    return e, sc


def alloc_planes(n: int, D: int, dev, slices: int = GRAM_SLICES):
    import torch

    return torch.empty(max(256, L.lib.am_planes_bytes(n, D, slices)), dtype=torch.int8, device=dev)


def normalize_planes(counts, method: str, row_index, center, sc, colsum, out=None, planes=None,
                     slices: int = GRAM_SLICES, want_out: bool = True, want_planes: bool = True):
    """Fused kernel 2 + quantisation (D = 512): fp64 rows, column sums and the int8 digit planes of
    ``round((x - center) * sc)`` in one pass.  ``center``/``sc``/``colsum`` have 512 entries; for ilr
    the returned matrix is the 511-column view of a pitch-512 buffer (last column zero).

    ``want_out=False`` writes only the planes (the Gram's operand, on the critical path),
    ``want_planes=False`` only the fp64 rows into ``out``; the two launches compute bit-identical
    values, so the rows can be written while the eigensolve runs (``pipeline.featurise``)."""
    import torch

    dev = counts.device
    m = L.NORM_METHODS[method]
    n_in, D = counts.shape
    n_out = n_in if row_index is None else int(row_index.numel())
    with L.on_device(dev):
        if out is None and want_out:
            out = torch.empty((n_out, D), dtype=torch.float64, device=dev)
        if planes is None and want_planes:
            planes = alloc_planes(n_out, D, dev, slices)
        work = _work(dev, L.lib.am_normalize_work_bytes(D), "norm") if colsum is not None else None
        L.check(
            L.lib.am_normalize_planes(L.t_ptr(counts), n_out, D, L.t_ptr(row_index), m,
                                      L.t_ptr(out) if want_out else None, L.t_ptr(colsum),
                                      L.t_ptr(center), L.t_ptr(sc), slices,
                                      L.t_ptr(planes) if want_planes else None, L.t_ptr(work), L.stream_ptr(dev))
        )
    X = None
    if want_out:
        X = out[:, : D - 1] if method == "ilr" else out
    return X, (planes if want_planes else None)


def quantize_planes(X, center, sc, planes=None, slices: int = GRAM_SLICES):
    import torch

    dev = X.device
    n, D = X.shape
    with L.on_device(dev):
        if planes is None:
            planes = alloc_planes(n, D, dev, slices)
        L.check(L.lib.am_quantize_planes(L.t_ptr(X), n, D, L.t_ptr(center), L.t_ptr(sc), slices, L.t_ptr(planes),
                                         L.stream_ptr(dev)))
    return planes


def gram_planes(planes, n: int, D: int, e, out=None, slices: int = GRAM_SLICES):
    """G += sum_r (x_r - center)(x_r - center)^T from the digit planes (tcgen05 kind::i8)."""
    import torch

    dev = planes.device
    with L.on_device(dev):
        if out is None:
            out = torch.zeros((D, D), dtype=torch.float64, device=dev)
        work = _work(dev, L.lib.am_gram_planes_work_bytes(n, D, slices), "gram_planes")
        L.check(L.lib.am_gram_planes(L.t_ptr(planes), n, D, slices, L.t_ptr(e), L.t_ptr(out), L.t_ptr(work),
                                     L.stream_ptr(dev)))
    return out


def cov_finish(G, colsum, center, n_t):
    """In place: G <- (G - n (mu - center)(mu - center)^T) / (n - 1); returns mu = colsum / n."""
    import torch

    dev = G.device
    D = G.shape[0]
    with L.on_device(dev):
        mu = torch.empty(D, dtype=torch.float64, device=dev)
        L.check(L.lib.am_cov_finish(L.t_ptr(G), D, L.t_ptr(colsum), L.t_ptr(center), L.t_ptr(n_t), L.t_ptr(mu),
                                    L.stream_ptr(dev)))
    return mu


def count_flags(present, row_sum, n: int):
    """int32[2] on the device: [min(col_present), min(row_sum)]."""
    import torch

    dev = present.device
    with L.on_device(dev):
        flags = torch.empty(2, dtype=torch.int32, device=dev)
        L.check(L.lib.am_count_flags(L.t_ptr(present), int(present.numel()), L.t_ptr(row_sum), int(n), L.t_ptr(flags),
                                     L.stream_ptr(dev)))
    return flags


def stats_pack(present, row_sum, n: int, out):
    """``out[:ncols]`` = 1.0 where the column occurs, ``out[ncols]`` = number of contigs without counts
    (the SUM-reducible form of ``count_flags``; rides in the Gram all-reduce)."""
    import torch

    dev = present.device
    with L.on_device(dev):
        L.check(L.lib.am_stats_pack(L.t_ptr(present), int(present.numel()), L.t_ptr(row_sum), int(n), L.t_ptr(out),
                                    L.stream_ptr(dev)))
    return out


def gram_recenter(G, colsum, c_old, c_new, n_t):
    """In place: the Gram about ``c_old`` becomes the Gram about ``c_new`` (``None`` = the origin)."""
    import torch

    dev = G.device
    with L.on_device(dev):
        L.check(L.lib.am_gram_recenter(L.t_ptr(G), int(G.shape[0]), L.t_ptr(colsum), L.t_ptr(c_old), L.t_ptr(c_new),
                                       L.t_ptr(n_t), L.stream_ptr(dev)))
    return G


def eigh_desc(A, max_sweeps: int = 30, tol: float = 0.0):
    """All eigenpairs of symmetric A, descending; rows of the returned matrix are
    eigenvectors with sklearn's svd_flip(u_based_decision=False) sign."""
    import torch

    dev = A.device
    D = A.shape[0]
    with L.on_device(dev):
        evals = torch.empty(D, dtype=torch.float64, device=dev)
        evecs = torch.empty((D, D), dtype=torch.float64, device=dev)
        work = _work(dev, L.lib.am_eigh_work_bytes(D), "eigh")
        L.check(
            L.lib.am_eigh(L.t_ptr(A), D, L.t_ptr(evals), L.t_ptr(evecs), L.t_ptr(work), max_sweeps,
                          float(tol), L.stream_ptr(dev))
        )
    return evals, evecs


_topk_state: Dict[str, bool] = {}
TOPK_MAX_D = 576
TOPK_MAX_P = 128


def eigh_topk(A, P: int, started=None, started_value: int = 0):
    """Leading ``P`` eigenpairs of symmetric ``A`` (descending) and the trace.

    ``started`` (int32 device tensor, 1 element) receives ``started_value`` when the tridiagonalisation cluster is
    resident (``am_eigh_topk_signal``) — and in any case once the solve has been queued behind it, so that a
    stream waiting on it can never hang when another solver path is taken.

    Cluster/DSMEM tridiagonalisation path for 3 <= D <= 576 and P <= 128; other shapes
    go through the full Jacobi solver (``eigh_desc``) and are cut to P."""
    import torch

    dev = A.device
    D = A.shape[0]
    if _topk_state.get("disabled") or not (3 <= D <= TOPK_MAX_D and 1 <= P <= min(D, TOPK_MAX_P)):
        if started is not None:
            started.fill_(int(started_value))
        evals, evecs = eigh_desc(A)
        return evals[:P].contiguous(), evecs[:P].contiguous(), torch.clamp(evals, min=0.0).sum().reshape(1)
    with L.on_device(dev):
        evals = torch.empty(P, dtype=torch.float64, device=dev)
        evecs = torch.empty((P, D), dtype=torch.float64, device=dev)
        trace = torch.empty(1, dtype=torch.float64, device=dev)
        work = _work(dev, L.lib.am_eigh_topk_work_bytes(D, P), "eigtop")
        rc = L.lib.am_eigh_topk_signal(L.t_ptr(A), D, P, L.t_ptr(evals), L.t_ptr(evecs), L.t_ptr(trace),
                                       L.t_ptr(work), L.t_ptr(started), int(started_value), L.stream_ptr(dev))
        if rc == L.AM_ECUDA and not _topk_state.get("disabled"):
            # e.g. a 16-CTA cluster cannot be scheduled on this part.  The device Jacobi solver (am_eigh) gives the
            # same eigenpairs 65x slower (82 ms against 1.25 ms at D = 512): a silent switch would be a performance
            # cliff, so it has to be asked for (still a CUDA path; there is no CPU fallback)
            import os
            import warnings

            msg = L.lib.am_last_error().decode(errors="replace")
            if os.environ.get("AUTOMETA_B200_ALLOW_JACOBI", "") != "1":
                raise L.AutometaB200Error(rc, "am_eigh_topk (cluster tridiagonalisation) is unavailable on this device: "
                                          + msg + "; set AUTOMETA_B200_ALLOW_JACOBI=1 to use the 65x slower am_eigh")
            warnings.warn("am_eigh_topk unavailable (%s); using am_eigh (AUTOMETA_B200_ALLOW_JACOBI=1)" % msg)
            _topk_state["disabled"] = True
        elif rc != L.AM_OK:
            L.check(rc)
        if _topk_state.get("disabled"):
            if started is not None:
                started.fill_(int(started_value))
            evals_all, evecs_all = eigh_desc(A)
            return (evals_all[:P].contiguous(), evecs_all[:P].contiguous(),
                    torch.clamp(evals_all, min=0.0).sum().reshape(1))
    return evals, evecs, trace


def project(X, mu, comps, out=None):
    import torch

    dev = X.device
    n, D = X.shape
    P = comps.shape[0]
    with L.on_device(dev):
        if out is None:
            out = torch.empty((n, P), dtype=torch.float64, device=dev)
        L.check(
            L.lib.am_project(L.t_ptr(X), n, D, L.t_ptr(mu), L.t_ptr(comps), P, L.t_ptr(out), L.stream_ptr(dev))
        )
    return out


def project_planes(planes, n: int, D: int, e, center, mu, comps, out=None, slices: int = GRAM_SLICES):
    """scores = (X - mu) @ comps^T from the int8 digit planes (tcgen05 kind::i8); P <= 64."""
    import torch

    dev = planes.device
    P = comps.shape[0]
    with L.on_device(dev):
        if out is None:
            out = torch.empty((n, P), dtype=torch.float64, device=dev)
        work = _work(dev, L.lib.am_project_planes_work_bytes(D, slices), "project_planes")
        L.check(L.lib.am_project_planes(L.t_ptr(planes), n, D, slices, L.t_ptr(e), L.t_ptr(center), L.t_ptr(mu),
                                        L.t_ptr(comps), P, L.t_ptr(out), L.t_ptr(work), L.stream_ptr(dev)))
    return out


class EpsSweep:
    """Incremental DBSCAN(min_samples=1) over increasing eps on one point set."""

    def __init__(self, X):
        import torch

        assert X.dtype == torch.float64 and X.dim() == 2 and X.is_contiguous()
        self.X = X
        self.dev = X.device
        self.n, self.F = X.shape
        with L.on_device(self.dev):
            self.parent = torch.empty(max(self.n, 1), dtype=torch.int32, device=self.dev)
            L.check(L.lib.am_iota_i32(L.t_ptr(self.parent), self.n, L.stream_ptr(self.dev)))
            self.work = torch.empty(max(256, L.lib.am_dbscan_work_bytes(self.n)), dtype=torch.uint8, device=self.dev)
            self.labels = torch.empty(max(self.n, 1), dtype=torch.int64, device=self.dev)
            self.n_clusters = torch.zeros(1, dtype=torch.int32, device=self.dev)
            if self.n:
                self.h_min = np.ascontiguousarray(X.min(dim=0).values.cpu().numpy(), dtype=np.float64)
                self.h_max = np.ascontiguousarray(X.max(dim=0).values.cpu().numpy(), dtype=np.float64)
            else:
                self.h_min = np.zeros(self.F)
                self.h_max = np.zeros(self.F)
        self.last_eps = 0.0

    def reset(self):
        import torch

        with L.on_device(self.dev):
            L.check(L.lib.am_iota_i32(L.t_ptr(self.parent), self.n, L.stream_ptr(self.dev)))
        self.last_eps = 0.0

    def step(self, eps: float):
        """Labels (int64 device tensor, valid until the next step) at ``eps``.
        eps may not decrease without ``reset()``."""
        import torch

        if eps < self.last_eps:
            self.reset()
        self.last_eps = eps
        with L.on_device(self.dev):
            st = L.stream_ptr(self.dev)
            L.check(
                L.lib.am_eps_union(L.t_ptr(self.X), self.n, self.F, float(eps), L.np_ptr(self.h_min),
                                   L.np_ptr(self.h_max), L.t_ptr(self.parent), L.t_ptr(self.work), st)
            )
            L.check(
                L.lib.am_labels_from_parent(L.t_ptr(self.parent), self.n, L.t_ptr(self.labels),
                                            L.t_ptr(self.n_clusters), L.t_ptr(self.work), st)
            )
        return self.labels[: self.n]


_metrics_work_cache: Dict[int, "torch.Tensor"] = {}


def _metrics_work(device, nbytes: int):
    """The marker scratch of ``am_cluster_metrics`` is a dense [points x markers] table that the kernels leave
    zeroed (every entry they touch is cleared again with atomicExch), so ONE zero-filled buffer per device is
    allocated, grown when a larger sweep comes, and reused by every ``ClusterMetrics`` — ``get_clusters`` and the
    taxon drivers build one per sweep, and a fresh n x M memset per call would scale with n * M although only
    the non-zero markers are ever touched.  (Sweeps on one device run one after the other on one stream.)"""
    import torch

    t = _metrics_work_cache.get(device.index)
    if t is None or t.numel() < nbytes:
        t = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        _metrics_work_cache[device.index] = t
    return t


class ClusterMetrics:
    """Device-side add_metrics + apply_binning_metrics_filter + median completeness for the sweep
    (reference binning/utilities.py:125-262, recursive_dbscan.py:180-192)."""

    def __init__(self, n_points: int, marker_rows, marker_cols, marker_vals, n_ref_markers: int, coverage, gc,
                 cutoffs, device):
        import torch

        self.dev = device
        self.n = int(n_points)
        self.M = int(n_ref_markers)
        self.cutoffs = tuple(float(c) for c in cutoffs)
        with L.on_device(device):
            # np.array(..., copy=True): pandas hands out read-only views, which torch.from_numpy warns about
            t = lambda a, dt: torch.from_numpy(np.array(a, dtype=dt, order="C", copy=True)).to(device)
            self.nnz = int(len(marker_rows))
            self.rows = t(marker_rows if self.nnz else np.zeros(1), np.int32)
            self.cols = t(marker_cols if self.nnz else np.zeros(1), np.int32)
            self.vals = t(marker_vals if self.nnz else np.zeros(1), np.int32)
            self.cov = t(coverage, np.float64)
            self.gc = t(gc, np.float64)
            n1 = max(self.n, 1)
            self.completeness = torch.empty(n1, dtype=torch.float64, device=device)
            self.purity = torch.empty(n1, dtype=torch.float64, device=device)
            self.cov_std = torch.empty(n1, dtype=torch.float64, device=device)
            self.gc_std = torch.empty(n1, dtype=torch.float64, device=device)
            self.present = torch.empty(n1, dtype=torch.int32, device=device)
            self.single = torch.empty(n1, dtype=torch.int32, device=device)
            self.summary = torch.zeros(4, dtype=torch.float64, device=device)
            self.work = _metrics_work(device, max(256, L.lib.am_metrics_work_bytes(self.n, self.M)))

    def compute(self, labels, n_clusters):
        """Queue the metrics of one eps; returns the device summary [n_clusters, median, n_pass]."""
        import torch

        c = self.cutoffs
        with L.on_device(self.dev):
            L.check(
                L.lib.am_cluster_metrics(
                    L.t_ptr(labels), self.n, L.t_ptr(n_clusters), L.t_ptr(self.rows), L.t_ptr(self.cols),
                    L.t_ptr(self.vals), self.nnz, self.M, L.t_ptr(self.cov), L.t_ptr(self.gc), c[0], c[1], c[2], c[3],
                    L.t_ptr(self.completeness), L.t_ptr(self.purity), L.t_ptr(self.cov_std), L.t_ptr(self.gc_std),
                    L.t_ptr(self.present), L.t_ptr(self.single), L.t_ptr(self.summary), L.t_ptr(self.work),
                    L.stream_ptr(self.dev),
                )
            )
        return self.summary

    def arrays(self, n_clusters: int):
        """Host copies of the per-cluster metrics of the LAST compute() call."""
        k = int(n_clusters)
        return dict(
            completeness=self.completeness[:k].cpu().numpy(),
            purity=self.purity[:k].cpu().numpy(),
            coverage_stddev=self.cov_std[:k].cpu().numpy(),
            gc_content_stddev=self.gc_std[:k].cpu().numpy(),
            present_marker_count=self.present[:k].cpu().numpy().astype(np.int64),
            single_copy_marker_count=self.single[:k].cpu().numpy().astype(np.int64),
        )


class DeviceSweep:
    """The whole eps sweep of ``recursive_dbscan`` (reference recursive_dbscan.py:172-215) resident on the
    device (``csrc/am_sweep.cu``): eps step, labels, metrics, the loop's decisions (best eps so far, step x 10,
    termination) and the copy of the best labels / metrics all run as device work; one step is captured into a
    CUDA graph and replayed.  The host queues batches of steps on the sweep's OWN stream and reads one word per
    batch, so several sweeps (the taxa of a rank) can be in flight at once.

    ``launch(steps)`` queues; ``ready()`` / ``wait()`` tell when the batch has run; ``poll()`` -> has the sweep
    ended; ``result()`` -> labels and metrics of the best eps on the host."""

    #: steps queued per batch (a batch may overshoot the end of the sweep: extra steps change nothing)
    BATCH = 16

    def __init__(self, X: np.ndarray, coverage: np.ndarray, gc: np.ndarray, marker_rows, marker_cols, marker_vals,
                 n_ref_markers: int, cutoffs, device, stream=None, trace_cap: int = 0, graph: bool = True):
        import ctypes as C

        import torch

        X = np.ascontiguousarray(X, dtype=np.float64)
        self.n, self.F = X.shape
        if self.n < 1:
            raise ValueError("DeviceSweep needs at least one point")
        if not 1 <= self.F <= 4:
            raise ValueError(f"the GPU eps step supports 1..4 clustering features, got {self.F}")
        self.dev = device
        self.M = int(n_ref_markers)
        self.trace_cap = int(trace_cap)
        self.stream = stream if stream is not None else torch.cuda.Stream(device=device)
        n = self.n
        with torch.cuda.device(device), torch.cuda.stream(self.stream):
            up = lambda a, dt: torch.from_numpy(np.array(a, dtype=dt, order="C", copy=True)).to(device)
            e = lambda shape, dt: torch.empty(shape, dtype=dt, device=device)
            self.X = up(X, np.float64)
            self.cov = up(coverage, np.float64)
            self.gc = up(gc, np.float64)
            self.nnz = int(len(marker_rows))
            self.rows = up(marker_rows if self.nnz else np.zeros(1), np.int32)
            self.cols = up(marker_cols if self.nnz else np.zeros(1), np.int32)
            self.vals = up(marker_vals if self.nnz else np.zeros(1), np.int32)
            self.parent = e(n, torch.int32)
            self.labels = e(n, torch.int64)
            self.n_clusters = e(1, torch.int32)
            self.summary = e(4, torch.float64)
            self.cur = [e(n, torch.float64) for _ in range(4)] + [e(n, torch.int32) for _ in range(2)]
            self.best_labels = e(n, torch.int64)
            self.best = [e(n, torch.float64) for _ in range(4)] + [e(n, torch.int32) for _ in range(2)]
            self.w_dbscan = e(max(256, L.lib.am_dbscan_work_bytes(n)), torch.uint8)
            # private marker scratch (zero-filled once): concurrent sweeps cannot share the per-device one
            self.w_metrics = torch.zeros(max(256, L.lib.am_metrics_work_bytes(n, self.M)), dtype=torch.uint8,
                                         device=device)
            self.ctl = e(max(256, L.lib.am_sweep_ctl_bytes(n, self.trace_cap)), torch.uint8)
            self.h_ctl = torch.zeros(L.SWEEP_CTL_WORDS, dtype=torch.float64).pin_memory()
            a = L.SweepArgs()
            a.d_X, a.n_points, a.F, a.n_markers = self.X.data_ptr(), n, self.F, self.M
            mn, mx = X.min(axis=0), X.max(axis=0)
            for d in range(self.F):
                a.h_min[d], a.h_max[d] = float(mn[d]), float(mx[d])
            a.d_parent, a.d_dbscan_work = self.parent.data_ptr(), self.w_dbscan.data_ptr()
            a.d_labels, a.d_n_clusters = self.labels.data_ptr(), self.n_clusters.data_ptr()
            a.d_marker_rows, a.d_marker_cols, a.d_marker_vals = (self.rows.data_ptr(), self.cols.data_ptr(),
                                                                 self.vals.data_ptr())
            a.nnz = self.nnz
            a.d_coverage, a.d_gc = self.cov.data_ptr(), self.gc.data_ptr()
            for i, c in enumerate(cutoffs):
                a.cutoffs[i] = float(c)
            (a.d_completeness, a.d_purity, a.d_cov_std, a.d_gc_std, a.d_present,
             a.d_single) = [t.data_ptr() for t in self.cur]
            a.d_summary, a.d_metrics_work, a.d_ctl = (self.summary.data_ptr(), self.w_metrics.data_ptr(),
                                                      self.ctl.data_ptr())
            a.d_best_labels = self.best_labels.data_ptr()
            (a.d_best_completeness, a.d_best_purity, a.d_best_cov_std, a.d_best_gc_std, a.d_best_present,
             a.d_best_single) = [t.data_ptr() for t in self.best]
            a.trace_cap = self.trace_cap
            self.args = a
            self._argp = C.addressof(a)
            st = self.stream.cuda_stream
            L.check(L.lib.am_sweep_init(self._argp, st))
            self.graph = None
            if graph:
                g = C.c_void_p()
                L.check(L.lib.am_sweep_graph_create(self._argp, st, C.addressof(g)))
                self.graph = g.value
        self.event = torch.cuda.Event()
        self.steps_queued = 0

    def launch(self, steps: int = 0) -> None:
        """Queue ``steps`` eps steps and the copy of the loop state to pinned host memory."""
        steps = int(steps) or self.BATCH
        st = self.stream.cuda_stream
        if self.graph:
            L.check(L.lib.am_sweep_graph_launch(self.graph, self._argp, steps, self.h_ctl.data_ptr(), st))
        else:
            L.check(L.lib.am_sweep_steps(self._argp, steps, self.h_ctl.data_ptr(), st))
        self.event.record(self.stream)
        self.steps_queued += steps

    def ready(self) -> bool:
        return self.event.query()

    def wait(self) -> None:
        self.event.synchronize()

    def poll(self) -> bool:
        """After the batch has run: True once the sweep has ended."""
        return bool(self.h_ctl[L.SWEEP_DONE].item() != 0.0)

    def state(self) -> np.ndarray:
        return self.h_ctl.numpy().copy()

    def result(self):
        """``(labels int64[n], per-cluster metrics dict, best median, n_clusters at the best eps, trace)`` —
        labels is None when no step was ever kept (cannot happen: the first step's median >= -inf)."""
        import torch

        self.wait()
        ctl = self.state()
        nc = int(ctl[L.SWEEP_BEST_NC])
        with torch.cuda.device(self.dev), torch.cuda.stream(self.stream):
            labels = self.best_labels.cpu().numpy()
            names = ("completeness", "purity", "coverage_stddev", "gc_content_stddev", "present_marker_count",
                     "single_copy_marker_count")
            m = {k: t[:nc].cpu().numpy() for k, t in zip(names, self.best)}
            trace = None
            if self.trace_cap:
                k = min(int(ctl[L.SWEEP_NSTEPS]), self.trace_cap)
                off = (L.SWEEP_CTL_WORDS * 8 + 255) // 256 * 256 + ((self.n + 1) * 4 + 255) // 256 * 256
                trace = self.ctl[off : off + k * 32].cpu().numpy().view(np.float64).reshape(k, 4)
        for k2 in ("present_marker_count", "single_copy_marker_count"):
            m[k2] = m[k2].astype(np.int64)
        return labels, m, float(ctl[L.SWEEP_BEST]), nc, trace

    def close(self) -> None:
        if self.graph:
            self.wait()
            L.check(L.lib.am_sweep_graph_destroy(self.graph))
            self.graph = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
