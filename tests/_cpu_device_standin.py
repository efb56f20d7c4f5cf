"""TEST-ONLY stand-in for ``autometa_b200.device`` on CPU tensors (oracle arithmetic).

``pipeline.featurise`` is host logic (which kernels and which collectives are issued, in which order,
with which sizes) around device calls.  The world_size-2 gloo tests swap this module in for
``pipeline._dev`` to exercise that host logic — in particular that every rank posts the same collectives
whatever its shard holds — without a GPU.  Nothing in the product imports it.
"""
from types import SimpleNamespace

import numpy as np
import torch

from oracle import oracle_np as O

CPU = torch.device("cpu")


def fake_packed(asm):
    """``asm``: a HostAssembly (slice); counted by the C restatement of the reference loop."""
    return SimpleNamespace(device=CPU, n_contigs=asm.n_contigs, lengths=asm.lengths.astype(np.int32), _plans={},
                           asm=asm)


def count_packed(packed, k=5):
    if packed.n_contigs:
        from oracle import oracle_c

        counts = oracle_c.count(packed.asm.seq, packed.asm.offsets, k, 1)
    else:
        counts = np.zeros((0, len(O.kmer_columns(k)[0])), dtype=np.int32)
    counts = torch.from_numpy(np.ascontiguousarray(counts.astype(np.int32)))
    present = (counts.sum(dim=0) > 0).to(torch.int32)
    return counts, present, counts.sum(dim=1).to(torch.int32)


def stats_pack(present, row_sum, n, out):
    ncols = present.numel()
    out[:ncols] = (present > 0).to(torch.float64)
    out[ncols] = float((row_sum[:n] <= 0).sum())
    return out


def _norm(counts, method, row_index, col_index):
    c = counts.numpy().astype(np.float64)
    if row_index is not None:
        c = c[row_index.numpy()]
    if col_index is not None:
        c = c[:, col_index.numpy()]
    if c.shape[0] == 0:
        return np.zeros((0, c.shape[1] - (method == "ilr")))
    if method == "am_clr":
        l = np.log(c + 1.0)  # the row sum of (x + 1) / sum(x) cancels in log(y / gmean(y)), as in the kernel
        return l - l.mean(axis=1, keepdims=True)
    # the kernels do not raise on an all-zero row (the pass is optimistic; the test is read back afterwards):
    # such rows come out non-finite here
    bad = c.sum(axis=1) <= 0
    c[bad] = 1.0
    X = O.clr(c) if method == "clr" else O.ilr(c)
    X[bad] = np.nan
    return X


def normalize_counts(counts, method="am_clr", row_index=None, col_index=None, out=None, colsum=None, colmaxabs=None):
    X = torch.from_numpy(np.ascontiguousarray(_norm(counts, method, row_index, col_index)))
    if colsum is not None:
        colsum[: X.shape[1]] += X.sum(dim=0)
    if colmaxabs is not None and X.shape[0]:
        colmaxabs[: X.shape[1]] = torch.maximum(colmaxabs[: X.shape[1]], X.abs().max(dim=0).values)
    return X


def colstats(X, cs, cm):
    cs += X.sum(dim=0)
    cm.copy_(torch.maximum(cm, X.abs().max(dim=0).values))
    return cs, cm


def plane_scales(center, colmaxabs=None, bound=-1.0, slices=6):
    return None, None


def normalize_planes(counts, method, row_index, center, sc, colsum, out=None, planes=None, slices=6,
                     want_out=True, want_planes=True):
    X = torch.from_numpy(np.ascontiguousarray(_norm(counts, method, row_index, None)))
    D = counts.shape[1]
    if colsum is not None:
        colsum[: X.shape[1]] += X.sum(dim=0)
    if want_out and out is not None:
        out[:, : X.shape[1]] = X
    pl = None
    if want_planes:
        pl = X - center[: X.shape[1]]
    return (out[:, : D - 1] if (want_out and method == "ilr") else (out if want_out else None)), pl


def quantize_planes(X, center, sc, planes=None, slices=6):
    return X - center


def gram_planes(planes, n, D, e, out=None, slices=6):
    out += planes.T @ planes
    return out


def gram_recenter(G, colsum, c_old, c_new, n_t):
    n = float(n_t.item())
    if n > 0:
        mu = colsum / n
        o = mu - (c_old if c_old is not None else 0.0)
        w = mu - (c_new if c_new is not None else 0.0)
        G += n * (torch.outer(w, w) - torch.outer(o, o))
    return G


def cov_finish(G, colsum, center, n_t):
    n = float(n_t.item())
    mu = colsum / n
    d = mu - (center if center is not None else mu)
    G.copy_((G - n * torch.outer(d, d)) / (n - 1.0))
    return mu.clone()


def eigh_topk(C, P, started=None, started_value=0):
    if not np.all(np.isfinite(C.numpy())):
        D = C.shape[0]
        return torch.zeros(P, dtype=torch.float64), torch.zeros((P, D), dtype=torch.float64), torch.ones(1, dtype=torch.float64)
    w, V = np.linalg.eigh(C.numpy())
    w, V = w[::-1][:P].copy(), V[:, ::-1][:, :P].T.copy()
    s = np.sign(V[np.arange(P), np.argmax(np.abs(V), axis=1)])
    V *= s[:, None]
    return torch.from_numpy(w), torch.from_numpy(V), torch.tensor([float(np.trace(C.numpy()))], dtype=torch.float64)


def project_planes(planes, n, D, e, center, mu, comps, out=None, slices=6):
    return (planes + (center - mu)) @ comps.T


def project(X, mu, comps, out=None):
    return (X - mu) @ comps.T
