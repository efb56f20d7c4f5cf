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
    """Strided row sample (n/32 rows, between 256 and 2048) for the provisional centre; cached per (n_out, device).
    The centre only has to be near the mean: the Gram correction n (mu - c)(mu - c)^T is 1/ns of G."""
    import torch

    key = ("sample", n_out, dev.index)
    t = cache.get(key)
    if t is None:
        ns = min(n_out, max(256, min(2048, n_out // 32)))
        t = torch.div(torch.arange(ns, device=dev, dtype=torch.int64) * n_out, ns, rounding_mode="floor")
        cache[key] = t
    return t


def _ctx(dev):
    import contextlib

    import torch

    return torch.cuda.device(dev) if dev.type == "cuda" else contextlib.nullcontext()


class _Side:
    """Second (low-priority) stream per device for work that is off the critical path: the fp64
    normalised rows are an OUTPUT of the pass but not an input of the Gram / eigensolve / projection
    (those read the digit planes), so they are written while the eigensolve — 16 SMs, latency-bound —
    leaves the rest of the chip idle."""

    _streams: Dict[int, Any] = {}

    @classmethod
    def stream(cls, dev):
        import torch

        st = cls._streams.get(dev.index)
        if st is None:
            st = torch.cuda.Stream(device=dev, priority=0)
            cls._streams[dev.index] = st
        return st


def _bound_for(packed, method: str, D: int) -> float:
    """Hard bound of the transform (fixed-point range of the digit planes, no pass over the data):
    |am_clr| <= log1p(longest contig); |clr|, |ilr| <= max log q - min log q <= log(1 / min q) with
    min q >= min(1/D^2, 1/(row sum))."""
    key = "log1p_maxlen" if method == "am_clr" else "log_range"
    bound = packed._plans.get(key)
    if bound is None:
        longest = float(packed.lengths.max()) if packed.n_contigs else 1.0
        if method == "am_clr":
            bound = float(np.log1p(longest))
        else:
            bound = float(max(2.0 * np.log(D), np.log(max(longest, 1.0))) + 1.0)
        packed._plans[key] = bound
    return bound


def _fit(packed, counts, stats, method, row_index, col_index, group, distributed, mark, overlap=False):
    """Stage 1 — normalise -> digit planes -> Gram -> (all-reduce) -> covariance, for the given row/column
    selection (device only, no host sync).  Everything whose work grows with the assembly.

    Every rank issues the same collectives with the same sizes whatever its shard holds (an empty
    shard, all-NA rows): the k = 5 path has exactly ONE all-reduce, of the buffer
    ``[Gram | column sums | row count | column-presence flags | all-NA row count]``; the generic path
    has two (``[column sums | row count]``, then the Gram).  ``stats`` = (col_present, row_sum) on the
    first pass, ``None`` on the re-run after rows / columns were dropped."""
    import torch

    dev = counts.device
    D = counts.shape[1]
    Dk = D if col_index is None else int(col_index.numel())
    Dout = Dk - 1 if method == "ilr" else Dk
    n_out = counts.shape[0] if row_index is None else int(row_index.numel())
    fused = method in ("am_clr", "clr", "ilr") and D == 512 and col_index is None
    GG = Dout * Dout
    # R: G | cs[D] | n | present[D] | bad rows — one SUM all-reduce carries all of it
    R = torch.zeros(GG + 2 * D + 2, dtype=torch.float64, device=dev)
    G = R[:GG].view(Dout, Dout)
    cs = R[GG : GG + Dout]
    n_t = R[GG + D : GG + D + 1]
    tail = R[GG + D + 1 :]
    zbuf = torch.zeros(2 * D + 1, dtype=torch.float64, device=dev)  # column max | centre sample sums
    planes = e = center = X = None
    side = None
    if stats is not None and packed.n_contigs:
        _dev.stats_pack(stats[0], stats[1], packed.n_contigs, tail)
    if fused:
        if n_out > 0:
            # provisional centre from a strided row sample of THIS shard (the exact mean needs the full
            # pass).  Ranks need not agree on it: each turns its Gram into one about the origin before
            # the all-reduce (am_gram_recenter), so no exchange is needed for the centre.
            sample = _sample_rows(n_out, dev, packed._plans)
            ns = int(sample.numel())
            if row_index is not None:
                sample = row_index[sample]
            cs_s = zbuf[D : 2 * D]
            if method == "ilr":  # the generic ilr kernel has no fused column sums; 2048 x 511 is tiny
                cs_s[:Dout] = _dev.normalize_counts(counts, method, sample, None).sum(dim=0)
            else:
                _dev.normalize_counts(counts, method, sample, None, colsum=cs_s)
            center = cs_s * (1.0 / ns)
            e, sc = _dev.plane_scales(center, None, _bound_for(packed, method, D))
            Xfull = torch.empty((n_out, D), dtype=torch.float64, device=dev)
            X = Xfull[:, : D - 1] if method == "ilr" else Xfull
            _, planes = _dev.normalize_planes(counts, method, row_index, center, sc, R[GG : GG + D],
                                              out=Xfull, want_out=not overlap)
            center = center[:Dout]
        else:
            X = torch.empty((0, Dout), dtype=torch.float64, device=dev)
        mark("normalise")
        n_t.fill_(float(n_out))
        if n_out > 0:
            _dev.gram_planes(planes, n_out, Dout, e, out=G)
            if overlap:
                # the fp64 rows: same kernel, planes switched off, queued behind the Gram on the side
                # stream so that it shares the chip with the eigensolve and not with the Gram
                cur = torch.cuda.current_stream(dev)
                side = _Side.stream(dev)
                ev = torch.cuda.Event()
                ev.record(cur)
                side.wait_event(ev)
                with torch.cuda.stream(side):
                    _dev.normalize_planes(counts, method, row_index, None, None, None, out=Xfull, want_planes=False)
        c_fin = center
        if distributed:
            if n_out > 0:
                _dev.gram_recenter(G, cs, center, None, n_t)
            _dist.allreduce_sum_([R], group)
            c_fin = zbuf[:Dout]  # zeros: the all-reduced Gram is about the origin
        elif n_out == 0:
            c_fin = None
    else:
        cm = zbuf[:Dout]  # column max|x|: fixed-point scale
        if method == "ilr" or Dout > 1024:
            X = _dev.normalize_counts(counts, method, row_index, col_index)
            if n_out > 0:
                _dev.colstats(X, cs, cm)
        else:
            X = _dev.normalize_counts(counts, method, row_index, col_index, colsum=cs, colmaxabs=cm)
        mark("normalise")
        n_t.fill_(float(n_out))
        if distributed:
            _dist.allreduce_sum_([R[GG:]], group)  # column sums | row count | flags
        if n_out > 0:
            # generic route to the same tensor-core kernels: quantise the fp64 matrix about the exact mean
            center = cs / n_t
            e, sc = _dev.plane_scales(center, cm)
            planes = _dev.quantize_planes(X, center, sc)
            _dev.gram_planes(planes, n_out, Dout, e, out=G)
        if distributed:
            _dist.allreduce_sum_([R[:GG]], group)
        c_fin = None
    # mu and C = (G - n (mu - centre)(mu - centre)^T) / (n - 1) in one launch (in place on G)
    mu = _dev.cov_finish(G, cs, c_fin, n_t)
    mark("gram")
    return dict(C=G, mu=mu, planes=planes, e=e, center=center, X=X, n_out=n_out, Dout=Dout, side=side,
                row_index=row_index, col_index=col_index, stats=tail, n_total=n_t, R=R, counts=counts,
                n_and_stats=R[GG + D :])  # [row count | presence flags | all-NA rows]: one read-back


def _eig(st, n_components, started=None, started_value=0, pack=False):
    """Stage 2a — top-P eigenpairs of the covariance (N-independent, latency-bound).  ``pack``: also return them
    as one buffer ``[eigenvalues | components | trace]`` (P + P D + 1 doubles, 0.2 MB) ready to be broadcast."""
    import torch

    C, Dout = st["C"], st["Dout"]
    P = min(n_components, Dout)
    evals, comps, total = _dev.eigh_topk(C, P, started, started_value)
    if not pack:
        return evals, comps, total
    buf = torch.empty(P + P * Dout + 1, dtype=torch.float64, device=C.device)
    buf[:P] = evals
    buf[P : P + P * Dout] = comps.reshape(-1)
    buf[P + P * Dout :] = total.reshape(-1)
    return buf


def _unpack_eig(buf, P, Dout):
    return buf[:P], buf[P : P + P * Dout].view(P, Dout), buf[P + P * Dout :]


def _scores(st, evals, comps, total, mark):
    """Stage 2b — the scores from the digit planes (or the fp64 rows) and the result record."""
    import torch

    mu, planes, e, center, X = st["mu"], st["planes"], st["e"], st["center"], st["X"]
    n_out, Dout, side = st["n_out"], st["Dout"], st["side"]
    dev = st["C"].device
    P = comps.shape[0]
    evals = torch.clamp(evals, min=0.0)
    if n_out == 0:
        scores = torch.empty((0, P), dtype=torch.float64, device=dev)
    elif planes is not None and P <= 64:
        scores = _dev.project_planes(planes, n_out, Dout, e, center, mu, comps)
    else:
        if side is not None:
            torch.cuda.current_stream(dev).wait_stream(side)
            side = None
        scores = _dev.project(X, mu, comps)
    if side is not None:
        torch.cuda.current_stream(dev).wait_stream(side)  # everything is ordered on the caller's stream again
    mark("project")
    return dict(scores=scores, explained_variance=evals, explained_variance_ratio=evals / total,
                components=comps, mean=mu, row_index=st["row_index"], col_index=st["col_index"], normalized=X,
                stats=st["stats"], n_total=st["n_total"], n_and_stats=st["n_and_stats"])


def _project(st, n_components, mark):
    """Stage 2 on one stream: eigensolve, then scores (the serial pass; every rank solves the same, bit-identical
    covariance, so no exchange is needed)."""
    evals, comps, total = _eig(st, n_components)
    mark("eigh")
    return _scores(st, evals, comps, total, mark)


def _fit_project(packed, counts, stats, method, n_components, row_index, col_index, group, distributed, mark,
                 overlap=False):
    return _project(_fit(packed, counts, stats, method, row_index, col_index, group, distributed, mark, overlap),
                    n_components, mark)


def featurise(packed: PackedAssembly, k: int = 5, method: str = "am_clr", n_components: int = 50,
              group=None, distributed: bool = False, keep: bool = True, marks=None,
              overlap: bool = False) -> Dict[str, Any]:
    """counts, normalised matrix and PCA scores of this rank's contig shard.

    am_clr's global column drop and all-NA row drop (kmers.py:369) are handled optimistically:
    the whole pass runs on the stream assuming every column occurs somewhere and every contig has
    counts (true for any real assembly at k = 5); the two tests are read back ONCE, after the last
    kernel has been queued, and the rare other case re-runs with explicit row/column gathers.
    With several GPUs the tests travel inside the Gram all-reduce, so every rank reads the same
    (global) answer and takes the same branch.

    ``overlap=True`` writes the fp64 normalised rows from a second launch on a side stream, behind the
    Gram, instead of from the pass that writes the digit planes.  Measured on B200 (profiles/r2c): the
    planes-only pass is ALU-bound (0.29 ms against 0.33 ms fused), and the row writer then shares SMs with
    the latency-critical eigensolve (+0.2 ms) — a net loss, so it is off by default.
    """
    import torch

    dev = packed.device
    mark = marks if marks is not None else (lambda name: None)
    if dev.type != "cuda":
        overlap = False  # only tests/test_dist_gloo.py gets here, with a stand-in for ``device``
    with _ctx(dev):
        mark("start")
        counts, present, row_sum = _dev.count_packed(packed, k)
        mark("count")
        res = _fit_project(packed, counts, (present, row_sum), method, n_components, None, None, group,
                           distributed, mark, overlap)
        D = counts.shape[1]
        back = res["n_and_stats"].cpu().numpy()  # the only host sync: after everything has been queued
        n_total, st = float(back[0]), back[1:]
        present_g = st[:D] > 0
        cols_ok, rows_ok = bool(present_g.all()), bool(st[D] == 0)
        if method != "am_clr":
            if n_total > 0 and not rows_ok:
                raise ValueError("Input matrix cannot have rows with all zeros")
        elif n_total > 0 and not (cols_ok and rows_ok):
            col_index = None if cols_ok else torch.from_numpy(np.flatnonzero(present_g).astype(np.int32)).to(dev)
            row_index = None if rows_ok else torch.nonzero(row_sum > 0).flatten()
            res = _fit_project(packed, counts, None, method, n_components, row_index, col_index, group, distributed,
                               lambda name: None, overlap)
        out = res
        out.pop("stats")
        out.pop("n_total")
        out.pop("n_and_stats")
        X = out.pop("normalized")
        if keep:
            if distributed:
                present = torch.from_numpy(present_g.astype(np.int32)).to(dev)
            out.update(counts=counts, normalized=X, col_present=present, row_sum=row_sum)
        return out


class _PassSlot:
    """Every device buffer one pass of the k = 5 path needs for a given shard shape, allocated once and reused
    every ``depth`` passes, with the raw pointers resolved once.  A pass is then queued with ~15 C calls and two
    NCCL calls — no allocation, no tensor-library kernel, no stream context — which is what bounds small shards
    (1/8 of C2 per GPU: 0.5 ms of GPU work per pass against 0.7 ms of host work through the general wrappers)."""

    def __init__(self, dev, n: int, D: int, method: str, P: int, ns: int):
        import torch

        self.n, self.D, self.method, self.P, self.ns = n, D, method, P, ns
        self.Dout = D - 1 if method == "ilr" else D
        self.GG = self.Dout * self.Dout
        f64, i32 = torch.float64, torch.int32
        e = lambda shape, dt: torch.empty(shape, dtype=dt, device=dev)
        self.counts = e((n, D), i32)
        self.present = e(D, i32)
        self.row_sum = e(max(n, 1), i32)
        self.R = e(self.GG + 2 * D + 2, f64)        # G | cs[D] | n | present[D] | all-NA rows
        self.zb = e(2 * D, f64)                      # zeros[D] | centre (sample sums, then means)[D]
        self.sample_out = e((max(ns, 1), D), f64)
        self.e = e(D, i32)
        self.sc = e(D, f64)
        self.X = e((max(n, 1), D), f64)
        self.planes = e(max(256, L.lib.am_planes_bytes(max(n, 1), D, 6)), torch.int8)
        self.mu = e(D, f64)
        self.eig = e(P + P * self.Dout + 1 + P, f64)  # eigenvalues | components | trace | ratio
        self.scores = e((n, P), f64)
        wb = lambda nbytes: e(max(256, int(nbytes)), torch.uint8)
        self.w_norm = wb(L.lib.am_normalize_work_bytes(D))
        self.w_colsum = wb(L.lib.am_colsum_work_bytes(D))
        self.w_gram = wb(L.lib.am_gram_planes_work_bytes(max(n, 1), self.Dout, 6))
        self.w_eig = wb(L.lib.am_eigh_topk_work_bytes(self.Dout, P))
        self.w_proj = wb(L.lib.am_project_planes_work_bytes(self.Dout, 6))
        p = lambda t: t.data_ptr()
        GG, Dout = self.GG, self.Dout
        self.p_counts, self.p_present, self.p_row_sum = p(self.counts), p(self.present), p(self.row_sum)
        self.p_R = p(self.R)
        self.p_cs = self.p_R + 8 * GG
        self.p_nt = self.p_R + 8 * (GG + D)
        self.p_tail = self.p_R + 8 * (GG + D + 1)
        self.p_zeros = p(self.zb)
        self.p_center = p(self.zb) + 8 * D
        self.p_sample, self.p_e, self.p_sc, self.p_X, self.p_planes, self.p_mu = (
            p(self.sample_out), p(self.e), p(self.sc), p(self.X), p(self.planes), p(self.mu))
        self.p_evals = p(self.eig)
        self.p_comps = self.p_evals + 8 * P
        self.p_total = self.p_comps + 8 * P * Dout
        self.p_ratio = self.p_total + 8
        self.p_scores = p(self.scores)
        self.p_w_norm, self.p_w_colsum, self.p_w_gram, self.p_w_eig, self.p_w_proj = (
            p(self.w_norm), p(self.w_colsum), p(self.w_gram), p(self.w_eig), p(self.w_proj))
        self.bcast = self.eig[: P + P * Dout + 1]
        self.n_and_stats = self.R[GG + D :]

    def result(self, keep: bool, clone: bool):
        P, Dout = self.P, self.Dout
        out = dict(scores=self.scores, explained_variance=self.eig[:P],
                   explained_variance_ratio=self.eig[P + P * Dout + 1 :],
                   components=self.eig[P : P + P * Dout].view(P, Dout), mean=self.mu[:Dout], row_index=None,
                   col_index=None)
        if keep:
            out.update(counts=self.counts, normalized=self.X[: self.n, :Dout], col_present=self.present,
                       row_sum=self.row_sum[: self.n])
        if clone:  # the slot is reused `depth` passes later: results that outlive it are copies
            out = {k2: (v.clone() if hasattr(v, "clone") else v) for k2, v in out.items()}
        return out


class FeatureStream:
    """Throughput mode: assemblies (or partitions of one) featurised back to back with up to ``depth`` of them
    in flight on two streams.

    A pass has parts of very different shape: stage 1 (count -> normalise -> Gram; all the work that grows
    with the assembly, every SM busy), the eigensolve (top-P eigenpairs of the 512 x 512 covariance; a
    latency-bound chain of small kernels, 16 SMs for most of its 1.2 ms) and the projection.  Run one after the
    other the chip idles through the eigensolve; here stage 1 of pass i+1 is queued on stream A while the
    eigensolve of pass i runs on stream B and its projection on stream C (both higher priority), so the
    eigensolve hides behind the next pass's counting.  Results are
    identical to ``featurise`` (same kernels, same order within a pass); the optimistic am_clr tests are read
    back when a pass retires and the rare failing pass is redone by ``featurise``."""

    def __init__(self, device, k: int = 5, method: str = "am_clr", n_components: int = 50, group=None,
                 distributed: bool = False, depth: int = 2, keep: bool = False, spread_eigh: bool = True,
                 on_result=None, lean: bool = True):
        """``spread_eigh`` (several GPUs): the eigensolve of pass i is done by rank i mod N only and broadcast,
        instead of being replicated on every rank — with N passes in flight each GPU solves one eigenproblem
        per N passes, which removes the replicated, N-independent 1.2 ms from every rank's critical path
        (the Amdahl term of strong scaling).  The broadcasts use their own communicator so that they never
        queue behind the next pass's Gram all-reduce."""
        import torch

        self.dev = L.require_cuda(device)
        self.k, self.method, self.P = k, method, n_components
        self.group, self.distributed, self.depth, self.keep = group, distributed, max(1, depth), keep
        # results are handed to ``on_result(result)`` as each pass retires (and not kept) when it is given;
        # otherwise they accumulate until ``drain()`` — device memory of every pass stays allocated until then
        self.on_result = on_result
        self.n_submitted = 0
        self.eig_group = None
        self.world = 1
        self.rank = 0
        # "the tridiagonalisation cluster of pass i is resident": stage 1 of pass i+1 waits for it (a stream-level
        # wait on this word), because its first kernel fills every SM for its whole duration and the cluster —
        # ready a few microseconds later, behind an event — would otherwise queue behind it
        self.started = None
        self.wait_value = 0
        self.use_wait = True
        if distributed and spread_eigh:
            import torch.distributed as dist

            self.world = dist.get_world_size(group)
            self.rank = dist.get_rank(group)
            if self.world > 1:
                ranks = list(range(dist.get_world_size())) if group is None else dist.get_process_group_ranks(group)
                self.eig_group = dist.new_group(ranks=ranks)  # collective call: every rank constructs the stream
        lo, hi = torch.cuda.Stream.priority_range()  # (lowest, highest): numerically hi <= lo
        self.sA = torch.cuda.Stream(device=self.dev, priority=lo)   # stage 1: count -> normalise -> Gram
        self.sB = torch.cuda.Stream(device=self.dev, priority=hi)   # this rank's eigensolves
        self.sC = torch.cuda.Stream(device=self.dev, priority=hi)   # (broadcast +) projection, in pass order
        self.inflight = []
        self.done = []
        self.lean = lean        # k = 5 passes queued from pre-allocated slots with raw C calls (_PassSlot)
        self.slots = {}
        self._events, self._ev_i = [], 0
        # host-side accounting: seconds spent queueing passes (submit, without the waits) and waiting for the
        # oldest pass in flight to finish (back-pressure): tells a launch-bound stream from a GPU-bound one
        self.host_queue_s = 0.0
        self.host_wait_s = 0.0

    def _lean_ok(self, packed) -> bool:
        return (self.lean and self.method in ("am_clr", "clr", "ilr") and self.k == 5 and self.P <= 64
                and packed.n_contigs > 0)

    def _submit_lean(self, packed):
        """Queue one pass of the k = 5 path from pre-allocated slot buffers with raw C calls (see _PassSlot);
        same kernels, same order, same results as ``_fit`` + ``_eig`` + ``_scores``."""
        import torch

        lib, dev = L.lib, self.dev
        n, D = packed.n_contigs, 512
        m = L.NORM_METHODS[self.method]
        sample = _sample_rows(n, dev, packed._plans)
        ns = int(sample.numel())
        slot_i = self.n_submitted % self.depth
        key = (n, self.method, self.P, ns)
        slot = self.slots.get(slot_i)
        if slot is None or slot.key != key:
            slot = _PassSlot(dev, n, D, self.method, self.P, ns)
            slot.key = key
            self.slots[slot_i] = slot
            torch.cuda.current_stream(dev).synchronize()
        S = slot
        plan = _dev.get_count_plan(packed, self.k)
        bound = _bound_for(packed, self.method, D)
        Dout, P = S.Dout, min(self.P, S.Dout)
        sa, sb, sc = self.sA.cuda_stream, self.sB.cuda_stream, self.sC.cuda_stream
        ck = L.check
        self.sA.wait_stream(torch.cuda.current_stream(dev))  # inputs produced on the caller's stream
        if self.wait_value and self.use_wait:
            self._hold_stage1(sa)
        # ---- stage 1 (stream A)
        ck(lib.am_zero_async(S.p_R, S.R.numel() * 8, sa))
        ck(lib.am_zero_async(S.p_zeros, S.zb.numel() * 8, sa))
        ck(lib.am_zero_async(S.p_present, D * 4, sa))
        ck(lib.am_count_kmers(L.t_ptr(packed.codes), L.t_ptr(packed.exc_bitmap), L.t_ptr(packed.exc_rank),
                              L.t_ptr(packed.exc_mask), L.t_ptr(packed.starts), L.t_ptr(plan.items), plan.n_items,
                              L.t_ptr(plan.multi), plan.n_multi, n, self.k, S.p_counts, S.p_present, S.p_row_sum,
                              L.t_ptr(plan.work), sa))
        ck(lib.am_stats_pack(S.p_present, D, S.p_row_sum, n, S.p_tail, sa))
        if self.method == "ilr":  # the generic ilr kernel has no fused column sums
            ck(lib.am_normalize(S.p_counts, ns, D, L.t_ptr(sample), None, D, m, S.p_sample, None, None, None, sa))
            ck(lib.am_colsum(S.p_sample, ns, Dout, S.p_center, S.p_w_colsum, sa))
        else:
            ck(lib.am_normalize(S.p_counts, ns, D, L.t_ptr(sample), None, D, m, S.p_sample, S.p_center, None,
                                S.p_w_norm, sa))
        ck(lib.am_axpb_f64(S.p_center, D, 1.0 / ns, 0.0, sa))  # sample sums -> provisional centre
        ck(lib.am_plane_scales(S.p_center, None, float(bound), D, 6, S.p_e, S.p_sc, sa))
        ck(lib.am_normalize_planes(S.p_counts, n, D, None, m, S.p_X, S.p_cs, S.p_center, S.p_sc, 6, S.p_planes,
                                   S.p_w_norm, sa))
        ck(lib.am_axpb_f64(S.p_nt, 1, 0.0, float(n), sa))       # row count
        ck(lib.am_gram_planes(S.p_planes, n, Dout, 6, S.p_e, S.p_R, S.p_w_gram, sa))
        c_fin = S.p_center
        if self.distributed:
            ck(lib.am_gram_recenter(S.p_R, Dout, S.p_cs, S.p_center, None, S.p_nt, sa))
            with torch.cuda.stream(self.sA):
                _dist.allreduce_sum_([S.R], self.group)
            c_fin = S.p_zeros  # the all-reduced Gram is about the origin
        ck(lib.am_cov_finish(S.p_R, Dout, S.p_cs, c_fin, S.p_nt, S.p_mu, sa))
        evA = self._event()
        evA.record(self.sA)
        # ---- stage 2a (stream B): this rank's eigensolves only
        owner = (self.n_submitted % self.world) if self.eig_group is not None else None
        self.n_submitted += 1
        solves_here = owner is None or owner == self.rank
        if solves_here:
            self.sB.wait_event(evA)
            ck(lib.am_eigh_topk_signal(S.p_R, Dout, P, S.p_evals, S.p_comps, S.p_total, S.p_w_eig,
                                       L.t_ptr(self.started) if self.depth > 1 else None, self.n_submitted, sb))
            evE = self._event()
            evE.record(self.sB)
            self.sC.wait_event(evE)
        else:
            self.sC.wait_event(evA)
        # ---- stage 2b (stream C): (broadcast +) ratio, projection
        if owner is not None:
            import torch.distributed as dist

            with torch.cuda.stream(self.sC):
                dist.broadcast(S.bcast, src=dist.get_global_rank(self.eig_group, owner), group=self.eig_group)
        ck(lib.am_pca_finish(S.p_evals, S.p_total, P, S.p_ratio, sc))
        ck(lib.am_project_planes(S.p_planes, n, Dout, 6, S.p_e, S.p_center, S.p_mu, S.p_comps, P, S.p_scores,
                                 S.p_w_proj, sc))
        evB = self._event()
        evB.record(self.sC)
        self.wait_value = self.n_submitted if (solves_here and self.depth > 1) else 0
        self.inflight.append((packed, S, None, evB, None, None))

    def _hold_stage1(self, raw_stream):
        """Stream-level wait until the previous pass's tridiagonalisation cluster is resident.  A scheduling aid
        only: where the driver refuses stream memory operations the stream simply runs without it."""
        rc = L.lib.am_stream_wait_geq(raw_stream, L.t_ptr(self.started), self.wait_value)
        if rc != L.AM_OK:
            import warnings

            warnings.warn("stream memory wait unavailable (%s): throughput mode runs without the start-order hold"
                          % L.lib.am_last_error().decode(errors="replace"))
            self.use_wait = False

    def _event(self):
        import torch

        # a small ring of reusable events (creating one costs more than recording it)
        if len(self._events) < 6 * self.depth + 6:
            ev = torch.cuda.Event()
            self._events.append(ev)
            return ev
        self._ev_i = (self._ev_i + 1) % len(self._events)
        return self._events[self._ev_i]

    #: SMs the statically split kernels of stage 1 (Gram, projection) leave to the tridiagonalisation cluster of
    #: the pass before (16 CTAs, one per SM): without it the Gram's last CTAs wait for the cluster to end and
    #: the two stages run one after the other (measured: no overlap at all on 2 GPUs)
    RESERVED_SMS = 16

    def submit(self, packed: PackedAssembly, marks=None):
        import torch

        import time

        mark = marks if marks is not None else (lambda name: None)
        t_in = time.perf_counter()
        w_in = self.host_wait_s
        if self.depth > 1 and L.lib.am_get_reserved_sms() != self.RESERVED_SMS:
            L.check(L.lib.am_set_reserved_sms(self.RESERVED_SMS))
        if self.started is None:
            with torch.cuda.device(self.dev):
                self.started = torch.zeros(1, dtype=torch.int32, device=self.dev)
                torch.cuda.current_stream(self.dev).synchronize()
        if marks is None and self._lean_ok(packed):
            while len(self.inflight) >= self.depth:
                self._retire()
            with L.on_device(self.dev):
                self._submit_lean(packed)
            self.host_queue_s += (time.perf_counter() - t_in) - (self.host_wait_s - w_in)
            return
        while len(self.inflight) >= self.depth:
            self._retire()
        with torch.cuda.device(self.dev):
            cur = torch.cuda.current_stream(self.dev)
            self.sA.wait_stream(cur)  # inputs produced on the caller's stream
            if self.started is None:
                self.started = torch.zeros(1, dtype=torch.int32, device=self.dev)
                torch.cuda.current_stream(self.dev).synchronize()
            with torch.cuda.stream(self.sA):
                if self.wait_value and self.use_wait:
                    self._hold_stage1(L.stream_ptr(self.dev))
                mark("start")
                counts, present, row_sum = _dev.count_packed(packed, self.k)
                mark("count")
                st = _fit(packed, counts, (present, row_sum), self.method, None, None, self.group, self.distributed,
                          mark, False)
                evA = torch.cuda.Event()
                evA.record(self.sA)
            owner = (self.n_submitted % self.world) if self.eig_group is not None else None
            self.n_submitted += 1
            solves_here = owner is None or owner == self.rank
            P, Dout = min(self.P, st["Dout"]), st["Dout"]
            eig = None
            if solves_here:
                # stream B holds nothing but this rank's own eigensolves, back to back: it never waits for a peer
                self.sB.wait_event(evA)
                with torch.cuda.stream(self.sB):
                    eig = _eig(st, self.P, self.started if self.depth > 1 else None, self.n_submitted,
                               pack=owner is not None)
                    evE = torch.cuda.Event()
                    evE.record(self.sB)
                self.sC.wait_event(evE)
            else:
                self.sC.wait_event(evA)
            with torch.cuda.stream(self.sC):
                if owner is not None:
                    import torch.distributed as dist

                    buf = eig if solves_here else torch.empty(P + P * Dout + 1, dtype=torch.float64, device=self.dev)
                    dist.broadcast(buf, src=dist.get_global_rank(self.eig_group, owner), group=self.eig_group)
                    eig = _unpack_eig(buf, P, Dout)
                mark("eigh")
                res = _scores(st, eig[0], eig[1], eig[2], mark)
                evB = torch.cuda.Event()
                evB.record(self.sC)
            # the next pass's stage 1 is held until this pass's cluster is resident (only where it runs)
            self.wait_value = self.n_submitted if (solves_here and self.depth > 1) else 0
            self.inflight.append((packed, st, res, evB, row_sum, present))
        self.host_queue_s += (time.perf_counter() - t_in) - (self.host_wait_s - w_in)

    def _retire(self):
        import torch

        import time

        packed, st, res, evB, row_sum, present = self.inflight.pop(0)
        t_w = time.perf_counter()
        evB.synchronize()
        self.host_wait_s += time.perf_counter() - t_w
        lean = isinstance(st, _PassSlot)
        D = st.D if lean else st["counts"].shape[1]
        back = (st.n_and_stats if lean else st["n_and_stats"]).cpu().numpy()  # the one host read of the pass
        n_total, flags = float(back[0]), back[1:]
        cols_ok, rows_ok = bool((flags[:D] > 0).all()), bool(flags[D] == 0)
        if n_total > 0 and not (rows_ok and (cols_ok or self.method != "am_clr")):
            # rare: redo this pass on the general path (row / column gathers, or the clr / ilr error).  The
            # library's scratch buffers are shared per device: let the passes in flight finish first
            with torch.cuda.device(self.dev):
                torch.cuda.synchronize(self.dev)
                res = featurise(packed, self.k, self.method, self.P, self.group, self.distributed, keep=self.keep)
        elif lean:
            res = st.result(self.keep, clone=self.on_result is None)
        else:
            res.pop("stats")
            res.pop("n_total")
            res.pop("n_and_stats")
            X = res.pop("normalized")
            if self.keep:
                res.update(counts=st["counts"], normalized=X, col_present=present, row_sum=row_sum)
        if self.on_result is not None:
            self.on_result(res)
        else:
            self.done.append(res)

    def drain(self):
        """Wait for everything submitted; returns the results in submission order."""
        while self.inflight:
            self._retire()
        L.check(L.lib.am_set_reserved_sms(0))
        out, self.done = self.done, []
        return out


class HostFeaturiser:
    """End-to-end entry for host buffers -> scores (+ explained variance) back on the host.
    Buffers are allocated once and reused across calls.

    ``host_format="packed"`` (default): the pinned host buffers hold the 2-bit codes and the sparse
    exception structure written by the threaded host packer (``am_pack_host`` — in a real run the
    FASTA parser fills them directly), 0.25 byte per base over PCIe.  ``host_format="ascii"``: pinned
    ASCII bytes, 1 byte per base over PCIe, packed on the device (``am_pack_device``)."""

    def __init__(self, asm: HostAssembly, device=None, k: int = 5, method: str = "am_clr", n_components: int = 50,
                 host_format: str = "packed", threads: int = 0):
        import time

        import torch

        if host_format not in ("packed", "ascii"):
            raise ValueError("host_format must be 'packed' or 'ascii'")
        self.dev = L.require_cuda(device)
        self.k, self.method, self.P = k, method, n_components
        self.host_format = host_format
        offsets = np.ascontiguousarray(asm.offsets, dtype=np.int64)
        base = int(offsets[0])
        self.n = asm.n_contigs
        self.total_bp = asm.total_bp
        self.ids = asm.ids
        self.host_pack_ms = None
        with torch.cuda.device(self.dev):
            if host_format == "ascii":
                self.starts_h, self.lengths, self.total = pack_plan(offsets)
                h_seq = torch.empty(self.total_bp, dtype=torch.uint8).pin_memory()
                h_seq.numpy()[:] = asm.seq[base : base + self.total_bp]
                h_off = torch.from_numpy(offsets - base).pin_memory()
                h_starts = torch.from_numpy(self.starts_h if self.n else np.zeros(1, np.int64)).pin_memory()
                self.host = [h_seq, h_off, h_starts]
                pad = [64, 0, 0]
            else:
                from .assembly import pack_host_arrays

                pack_host_arrays(asm.slice(0, min(self.n, 64)), threads)  # thread pool / page warm-up
                t0 = time.perf_counter()
                starts, lengths, total, codes, bitmap, rank, masks = pack_host_arrays(asm, threads)
                self.host_pack_ms = (time.perf_counter() - t0) * 1e3
                self.starts_h, self.lengths, self.total = starts, lengths, total
                self.n_exc = len(masks)
                pin = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a).view(dt)).pin_memory()
                self.host = [pin(codes, np.int32), pin(bitmap, np.int32), pin(rank, np.int32),
                             pin(masks if len(masks) else np.zeros(1, np.uint64), np.int64),
                             pin(starts if self.n else np.zeros(1, np.int64), np.int64)]
                pad = [0, 0, 0, 0, 0]
            self.bufs = [[torch.empty(h.numel() + p, dtype=h.dtype, device=self.dev) for h, p in zip(self.host, pad)]]
            self.h_scores = None
        self.h2d_bytes = sum(h.numel() * h.element_size() for h in self.host)
        self.d2h_bytes = 0
        self.exc_cap = 0
        self._plans = None

    def describe(self) -> str:
        if self.host_format == "packed":
            return ("pinned host 2-bit codes + exception masks (threaded host packer am_pack_host, 0.25 B/bp) -> H2D "
                    "-> count -> normalise -> PCA -> D2H scores; back-to-back steps through FeatureStream: the H2D "
                    "copy of step i+1 on a copy stream, up to 3 steps in flight on the compute streams, the scores "
                    "of each step read back as it retires")
        return ("pinned host ASCII (1 B/bp) -> H2D -> am_pack_device -> count -> normalise -> PCA -> D2H scores; "
                "back-to-back steps, the H2D copy of step i+1 overlaps the compute of step i (PCIe-bound)")

    def _upload(self, slot: int):
        for h, d in zip(self.host, self.bufs[slot]):
            d[: h.numel()].copy_(h, non_blocking=True)

    def run(self, group=None, distributed: bool = False):
        """One assembly, strictly serial: H2D -> (pack) -> featurise -> D2H (latency path)."""
        import torch

        with torch.cuda.device(self.dev):
            self._upload(0)
            return self._compute(self.bufs[0], group, distributed)

    def run_stream(self, n_steps: int, group=None, distributed: bool = False, depth: int = 3):
        """``n_steps`` assemblies back to back through ``FeatureStream`` (packed host format): the H2D copy of step
        i+1 runs on a copy stream while steps i, i-1, ... are in flight on the compute streams (the eigensolve of one
        beside the counting of the next; on several GPUs one eigensolve per N steps per GPU), and the scores of
        every step are read back to pinned host memory as the step retires.  Every step copies its own input from
        pinned host memory and returns its own scores; nothing is shared between steps but the allocations.
        Returns the last step's host results."""
        import torch

        if self.host_format != "packed":
            return self.run_many(n_steps, group, distributed)
        with torch.cuda.device(self.dev):
            nbuf = depth + 1
            while len(self.bufs) < nbuf:
                self.bufs.append([torch.empty_like(b) for b in self.bufs[0]])
            if getattr(self, "copy_stream", None) is None:
                self.copy_stream = torch.cuda.Stream(device=self.dev)
            evs = [torch.cuda.Event() for _ in range(nbuf)]
            cur = torch.cuda.current_stream(self.dev)
            if self.h_scores is None:
                self.run(group, distributed)  # sizes the pinned result buffers

            def collect(res):
                # runs when a step retires (its device work is complete): D2H of its scores on the caller's stream
                self.h_scores.copy_(res["scores"], non_blocking=True)
                self.h_ev.copy_(res["explained_variance"], non_blocking=True)

            key = (id(group), distributed, depth)
            if getattr(self, "_fs_key", None) != key:
                # built once per featuriser: with several GPUs it owns a communicator (a collective to create)
                self._fs = FeatureStream(self.dev, self.k, self.method, self.P, group=group, distributed=distributed,
                                         depth=depth)
                self._fs_key = key
            fs = self._fs
            fs.on_result = collect
            packed = [self._packed(b) for b in self.bufs]
            for p in packed:
                if self._plans:
                    p._plans = self._plans

            def upload(i):
                slot = i % nbuf
                # slot's previous user (step i - nbuf) has retired: FeatureStream keeps at most `depth` in flight
                with torch.cuda.stream(self.copy_stream):
                    self._upload(slot)
                    evs[slot].record(self.copy_stream)

            upload(0)
            for i in range(n_steps):
                cur.wait_event(evs[i % nbuf])
                fs.submit(packed[i % nbuf])
                self._plans = packed[i % nbuf]._plans
                if i + 1 < n_steps:
                    upload(i + 1)  # its slot's previous user, step i - depth, retired inside submit(i)
            fs.drain()
            cur.synchronize()
            self.d2h_bytes = self.h_scores.numel() * 8 + self.h_ev.numel() * 8
        return self.h_scores, self.h_ev

    def run_many(self, n_steps: int, group=None, distributed: bool = False):
        """``n_steps`` assemblies back to back with double buffering: the H2D copy of step i+1 runs on
        a copy stream while step i is featurised.  Every step still copies its input from pinned host
        memory and reads its scores back; only the PCIe time of the NEXT step is hidden behind the
        compute of the current one.  Returns the last step's host results."""
        import torch

        with torch.cuda.device(self.dev):
            if len(self.bufs) == 1:
                self.bufs.append([torch.empty_like(b) for b in self.bufs[0]])
                self.copy_stream = torch.cuda.Stream(device=self.dev)
                self.ev_copied = [torch.cuda.Event(), torch.cuda.Event()]
            cur = torch.cuda.current_stream(self.dev)

            def submit(slot):
                # ALL host->device traffic of a step goes through the copy stream: a small copy queued on
                # the compute stream would sit behind the next step's transfer on the same DMA engine
                self.copy_stream.wait_stream(cur)  # the buffer's previous consumer has been queued before
                with torch.cuda.stream(self.copy_stream):
                    self._upload(slot)
                    self.ev_copied[slot].record(self.copy_stream)

            out = None
            submit(0)
            for i in range(n_steps):
                slot = i & 1
                cur.wait_event(self.ev_copied[slot])
                if i + 1 < n_steps:
                    submit((i + 1) & 1)  # buffer (i+1)&1 was last used by step i-1, which has completed
                out = self._compute(self.bufs[slot], group, distributed)
            return out

    def _packed(self, bufs):
        if self.host_format == "ascii":
            d_seq, d_off, d_starts = bufs
            packed = pack_device_arrays(d_seq, d_off, d_starts, self.n, self.total,
                                        self.lengths, self.starts_h, self.total_bp, exc_capacity=self.exc_cap)
            self.exc_cap = max(self.exc_cap, packed.n_exc + 1024)
            return packed
        codes, bitmap, rank, masks, starts = bufs
        return PackedAssembly(n_contigs=self.n, total_bp=self.total_bp, total_packed_bases=self.total,
                              lengths=self.lengths, starts_host=self.starts_h, codes=codes, exc_bitmap=bitmap,
                              exc_rank=rank, exc_mask=masks, starts=starts, n_exc=self.n_exc, ids=self.ids)

    def _compute(self, bufs, group, distributed):
        import torch

        with torch.cuda.device(self.dev):
            packed = self._packed(bufs)
            if self._plans:
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
