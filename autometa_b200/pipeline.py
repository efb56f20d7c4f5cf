"""Device-resident featurisation: count -> normalise -> PCA in one pass.

This is the path ``bench.py`` times and what ``autometa-kmers`` amounts to once
the TSV hand-offs between count / normalize / embed (kmers.py:709-851) are
removed: the packed contigs stay in HBM, every stage reads the previous stage's
tensor, and with several GPUs each rank owns a bp-balanced contig shard and only
the tiny reductions in ``dist.py`` cross NVLink.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import numpy as np

from . import _lib as L
from . import device as _dev
from . import dist as _dist
from .assembly import HostAssembly, PackedAssembly, pack_device_arrays, pack_plan


def _sample_rows(n_out: int, dev, cache: dict):
    """Strided row sample (<= 2048 rows) for the provisional centre; cached per (n_out, device)."""
    import torch

    key = ("sample", n_out, dev.index)
    t = cache.get(key)
    if t is None:
        ns = min(n_out, 2048)
        t = torch.div(torch.arange(ns, device=dev, dtype=torch.int64) * n_out, ns, rounding_mode="floor")
        cache[key] = t
    return t


def _fit_project(packed, counts, method, n_components, row_index, col_index, group, distributed, mark):
    """normalise -> PCA fit -> scores for the given row/column selection (device only, no host sync)."""
    import torch

    dev = counts.device
    D = counts.shape[1]
    Dk = D if col_index is None else int(col_index.numel())
    Dout = Dk - 1 if method == "ilr" else Dk
    n_out = counts.shape[0] if row_index is None else int(row_index.numel())
    zbuf = torch.zeros(2 * D + D + 2, dtype=torch.float64, device=dev)  # cs | cm | sample sums + count
    cs = zbuf[:Dout]
    fused = method in ("am_clr", "clr", "ilr") and D == 512 and col_index is None and n_out > 0
    # fused path: Gram, column sums and row count share one buffer so that ONE all-reduce carries all three
    gbuf = torch.zeros(Dout * Dout + D + 1, dtype=torch.float64, device=dev) if fused else None
    if fused:
        cs = gbuf[Dout * Dout : Dout * Dout + Dout]
    planes = e = center = None
    if fused:
        # provisional centre from a strided row sample (the exact mean needs the full pass);
        # fixed-point range from the hard bound of the transform: |am_clr| <= log1p(longest
        # contig), |clr|, |ilr| <= log(max(D^2, longest contig)) + 1
        sample = _sample_rows(n_out, dev, packed._plans)
        ns = int(sample.numel())
        if row_index is not None:
            sample = row_index[sample]
        cs_s = zbuf[2 * D : 2 * D + D + 1]
        if method == "ilr":  # the generic ilr kernel has no fused column sums; 2048 x 511 is tiny
            cs_s[:Dout] = _dev.normalize_counts(counts, method, sample, None).sum(dim=0)
        else:
            _dev.normalize_counts(counts, method, sample, None, colsum=cs_s[:D])
        if distributed:
            cs_s[D] = float(ns)
            _dist.allreduce_sum_([cs_s], group)
            center = cs_s[:D] / cs_s[D]
        else:
            center = cs_s[:D] * (1.0 / ns)
        if method == "am_clr":
            bound = packed._plans.get("log1p_maxlen")
            if bound is None:
                bound = float(np.log1p(float(packed.lengths.max()) if packed.n_contigs else 1.0))
                packed._plans["log1p_maxlen"] = bound
        else:
            # |clr|, |ilr| <= max log q - min log q <= log(1 / min q), min q >= min(1/D^2, 1/(row sum))
            bound = packed._plans.get("log_range")
            if bound is None:
                longest = float(packed.lengths.max()) if packed.n_contigs else 1.0
                bound = float(max(2.0 * np.log(D), np.log(max(longest, 1.0))) + 1.0)
                packed._plans["log_range"] = bound
        e, sc = _dev.plane_scales(center, None, bound)
        X, planes = _dev.normalize_planes(counts, method, row_index, center, sc, gbuf[Dout * Dout : Dout * Dout + D])
        center = center[:Dout]
    elif method == "ilr" or Dout > 1024:
        X = _dev.normalize_counts(counts, method, row_index, col_index)
        cm = zbuf[Dout : 2 * Dout]
        _dev.colstats(X, cs, cm)
    else:
        cm = zbuf[Dout : 2 * Dout]  # column max|x|: fixed-point scale
        X = _dev.normalize_counts(counts, method, row_index, col_index, colsum=cs, colmaxabs=cm)
    mark("normalise")
    if fused:
        n_t = gbuf[Dout * Dout + D :]
        n_t.fill_(float(X.shape[0]))
    else:
        n_t = torch.full((1,), float(X.shape[0]), dtype=torch.float64, device=dev)
        if distributed:
            _dist.allreduce_sum_([cs, n_t], group)
    mu = None
    if not fused and X.shape[0] > 0:
        mu = cs / n_t
        # generic route to the same tensor-core kernels: quantise the fp64 matrix about its exact mean
        center = mu
        e, sc = _dev.plane_scales(center, cm)
        planes = _dev.quantize_planes(X, center, sc)
    if fused:
        G = _dev.gram_planes(planes, X.shape[0], Dout, e, out=gbuf[: Dout * Dout].view(Dout, Dout))
        if distributed:
            _dist.allreduce_sum_([gbuf], group)  # G | column sums | row count
    else:
        G = _dev.gram_planes(planes, X.shape[0], Dout, e) if planes is not None else torch.zeros(
            (Dout, Dout), dtype=torch.float64, device=dev)
        if distributed:
            _dist.allreduce_sum_([G], group)
    # mu and C = (G - n (mu - centre)(mu - centre)^T) / (n - 1) in one launch (in place on G)
    mu = _dev.cov_finish(G, cs, center if fused else None, n_t)
    C = G
    mark("gram")
    P = min(n_components, Dout)
    evals, comps, total = _dev.eigh_topk(C, P)
    mark("eigh")
    evals = torch.clamp(evals, min=0.0)
    if planes is not None and P <= 64:
        scores = _dev.project_planes(planes, X.shape[0], Dout, e, center, mu, comps)
    else:
        scores = _dev.project(X, mu, comps)
    mark("project")
    return dict(scores=scores, explained_variance=evals, explained_variance_ratio=evals / total,
                components=comps, mean=mu, row_index=row_index, col_index=col_index, normalized=X)


def featurise(packed: PackedAssembly, k: int = 5, method: str = "am_clr", n_components: int = 50,
              group=None, distributed: bool = False, keep: bool = True, marks=None) -> Dict[str, Any]:
    """counts, normalised matrix and PCA scores of this rank's contig shard.

    am_clr's global column drop and all-NA row drop (kmers.py:369) are handled optimistically:
    the whole pass runs on the stream assuming every column occurs somewhere and every contig has
    counts (true for any real assembly at k = 5); the two flags are read back ONCE, after the last
    kernel has been queued, and the rare other case re-runs with explicit row/column gathers.
    """
    import torch

    dev = packed.device
    mark = marks if marks is not None else (lambda name: None)
    with torch.cuda.device(dev):
        mark("start")
        counts, present, row_sum = _dev.count_packed(packed, k)
        mark("count")
        if distributed:
            _dist.allreduce_max_(present, group)
        if packed.n_contigs:
            flags = _dev.count_flags(present, row_sum, packed.n_contigs)
        else:
            flags = torch.ones(2, dtype=torch.int32, device=dev)
        res = None
        if packed.n_contigs:
            res = _fit_project(packed, counts, method, n_components, None, None, group, distributed, mark)
        flags_h = flags.cpu().numpy()  # the only host sync: after everything has been queued
        cols_ok, rows_ok = bool(flags_h[0] > 0), bool(flags_h[1] > 0)
        if method != "am_clr":
            if packed.n_contigs and not rows_ok:
                raise ValueError("Input matrix cannot have rows with all zeros")
        elif res is None or not (cols_ok and rows_ok):
            col_index = None if cols_ok else torch.nonzero(present > 0).flatten().to(torch.int32)
            row_index = None if rows_ok else torch.nonzero(row_sum > 0).flatten()
            res = _fit_project(packed, counts, method, n_components, row_index, col_index, group, distributed,
                               lambda name: None)
        out = res
        X = out.pop("normalized")
        if keep:
            out.update(counts=counts, normalized=X, col_present=present, row_sum=row_sum)
        return out


class HostFeaturiser:
    """End-to-end entry for host buffers: ASCII contig bytes in pinned host memory
    -> H2D -> device pack -> featurise -> scores (+ explained variance) back on the
    host.  Buffers are allocated once and reused across calls."""

    def __init__(self, asm: HostAssembly, device=None, k: int = 5, method: str = "am_clr", n_components: int = 50):
        import torch

        self.dev = L.require_cuda(device)
        self.k, self.method, self.P = k, method, n_components
        offsets = np.ascontiguousarray(asm.offsets, dtype=np.int64)
        base = int(offsets[0])
        self.n = asm.n_contigs
        self.total_bp = asm.total_bp
        self.starts_h, self.lengths, self.total = pack_plan(offsets)
        with torch.cuda.device(self.dev):
            self.h_seq = torch.empty(self.total_bp, dtype=torch.uint8).pin_memory()
            self.h_seq.numpy()[:] = asm.seq[base : base + self.total_bp]
            self.h_off = torch.from_numpy(offsets - base).pin_memory()
            self.h_starts = torch.from_numpy(self.starts_h if self.n else np.zeros(1, np.int64)).pin_memory()
            self.d_seq = torch.empty(self.total_bp + 64, dtype=torch.uint8, device=self.dev)
            self.d_off = torch.empty(self.h_off.shape, dtype=torch.int64, device=self.dev)
            self.d_starts = torch.empty(self.h_starts.shape, dtype=torch.int64, device=self.dev)
            self.h_scores = None
        self.h2d_bytes = self.total_bp + self.h_off.numel() * 8 + self.h_starts.numel() * 8
        self.d2h_bytes = 0
        self.exc_cap = 0

    def run(self, group=None, distributed: bool = False):
        """One assembly, strictly serial: H2D -> pack -> featurise -> D2H (latency path)."""
        import torch

        with torch.cuda.device(self.dev):
            self.d_seq[: self.total_bp].copy_(self.h_seq, non_blocking=True)
            self.d_off.copy_(self.h_off, non_blocking=True)
            self.d_starts.copy_(self.h_starts, non_blocking=True)
            return self._compute(self.d_seq, self.d_off, self.d_starts, group, distributed)

    def run_many(self, n_steps: int, group=None, distributed: bool = False):
        """``n_steps`` assemblies back to back with double buffering: the H2D copy of step i+1 runs on
        a copy stream while step i is packed and featurised.  Every step still copies its input from
        pinned host memory and reads its scores back; only the PCIe time of the NEXT step is hidden
        behind the compute of the current one.  Returns the last step's host results."""
        import torch

        with torch.cuda.device(self.dev):
            if getattr(self, "d_seq2", None) is None:
                self.d_seq2 = torch.empty_like(self.d_seq)
                self.d_off2 = torch.empty_like(self.d_off)
                self.d_starts2 = torch.empty_like(self.d_starts)
                self.copy_stream = torch.cuda.Stream(device=self.dev)
                self.ev_copied = [torch.cuda.Event(), torch.cuda.Event()]
            bufs = [(self.d_seq, self.d_off, self.d_starts), (self.d_seq2, self.d_off2, self.d_starts2)]
            cur = torch.cuda.current_stream(self.dev)

            def submit(slot):
                # ALL host->device traffic of a step goes through the copy stream: a small copy queued on
                # the compute stream would sit behind the next step's 1 GB transfer on the same DMA engine
                self.copy_stream.wait_stream(cur)  # the buffer's previous consumer has been queued before
                with torch.cuda.stream(self.copy_stream):
                    bufs[slot][0][: self.total_bp].copy_(self.h_seq, non_blocking=True)
                    bufs[slot][1].copy_(self.h_off, non_blocking=True)
                    bufs[slot][2].copy_(self.h_starts, non_blocking=True)
                    self.ev_copied[slot].record(self.copy_stream)

            out = None
            submit(0)
            for i in range(n_steps):
                slot = i & 1
                cur.wait_event(self.ev_copied[slot])
                if i + 1 < n_steps:
                    submit((i + 1) & 1)  # buffer (i+1)&1 was last used by step i-1, which has completed
                out = self._compute(bufs[slot][0], bufs[slot][1], bufs[slot][2], group, distributed)
            return out

    def _compute(self, d_seq, d_off, d_starts, group, distributed):
        import torch

        with torch.cuda.device(self.dev):
            packed = pack_device_arrays(d_seq, d_off, d_starts, self.n, self.total,
                                        self.lengths, self.starts_h, self.total_bp, exc_capacity=self.exc_cap)
            self.exc_cap = max(self.exc_cap, packed.n_exc + 1024)
            if getattr(self, "_plans", None):
                packed._plans = self._plans  # the work list depends only on the contig lengths
            res = featurise(packed, self.k, self.method, self.P, group=group, distributed=distributed, keep=False)
            self._plans = packed._plans
            scores = res["scores"]
            if self.h_scores is None or self.h_scores.shape != scores.shape:
                self.h_scores = torch.empty(scores.shape, dtype=scores.dtype).pin_memory()
                self.h_ev = torch.empty(res["explained_variance"].shape, dtype=torch.float64).pin_memory()
            self.h_scores.copy_(scores, non_blocking=True)
            self.h_ev.copy_(res["explained_variance"], non_blocking=True)
            torch.cuda.current_stream(self.dev).synchronize()
            self.d2h_bytes = self.h_scores.numel() * 8 + self.h_ev.numel() * 8
        return self.h_scores, self.h_ev
