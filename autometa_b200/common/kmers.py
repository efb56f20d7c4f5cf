"""Drop-in for the k-mer featurisation functions of ``autometa.common.kmers``.

Same names, arguments, DataFrame schemas, TSV formats and error behaviour as the
reference (autometa/common/kmers.py), with the arithmetic done by the sm_100a
kernels behind ``libautometa_b200.so``:

================  =====================================  ======================
function          reference                              kernel
================  =====================================  ======================
``init_kmers``    kmers.py:56-89                         host table (C ABI)
``count``         kmers.py:268-330 (+122-265 workers)    am_count_kmers
``autometa_clr``  kmers.py:333-373                       am_normalize (am_clr)
``normalize``     kmers.py:376-432                       am_normalize
``pca``           kmers.py:588-607 (sklearn PCA inside)  am_gram / am_eigh / am_project
``embed``         kmers.py:435-706                       PCA stage on GPU; manifold step passed through
================  =====================================  ======================

There is no CPU fallback: without a CUDA device these functions raise.

Where this module is narrower than, or differs from, the reference (INTEGRATION.md 3b has the reasons):
``count`` takes ``size`` 1..7 only (larger k: ``AutometaB200Error``, "k must be in 1..7") and returns nullable
``Int32`` columns instead of object columns (same values, same ``pd.NA``, same TSV bytes); the PCA stage of
``embed`` always computes the exact leading eigenpairs, also for the table sizes (500 < N < 10 D) at which
sklearn's ``svd_solver="auto"`` takes the randomized branch — there the scores agree with the reference's only to
the randomized solver's own accuracy and the shared ``RandomState`` is not advanced by the PCA.
"""
from __future__ import annotations

import logging
import multiprocessing as mp
import os
from typing import Any, Dict, Union

import numpy as np
import pandas as pd

from .. import _lib as L
from .. import assembly as _asm
from .. import device as _dev
from .. import tsv as _tsv
from .exceptions import TableFormatError

logger = logging.getLogger(__name__)


def _revcomp(string: str) -> str:
    """Reverse complement of an upper-case ACGT string (reference kmers.py:38-53)."""
    comp = {"A": "T", "T": "A", "C": "G", "G": "C"}
    return "".join(comp.get(ch) for ch in reversed(string))


def init_kmers(kmer_size: int = 5) -> Dict[str, int]:
    """``{kmer: column}`` in the reference's column order (kmers.py:56-89): k-mers
    enumerated over the letters A,T,C,G, one representative per reverse-complement
    pair (the first one met).  The table comes from ``am_kmer_column_names``."""
    if not isinstance(kmer_size, int) or isinstance(kmer_size, bool):
        raise TypeError(f"kmer_size must be an int! Given: {type(kmer_size)}")
    return {name: i for i, name in enumerate(_dev.column_names(kmer_size))}


def load(kmers_fpath: str) -> pd.DataFrame:
    """Load a k-mer table written by :func:`count` / :func:`normalize` (kmers.py:92-119)."""
    if not os.path.exists(kmers_fpath) or not os.path.getsize(kmers_fpath):
        raise FileNotFoundError(kmers_fpath)
    try:
        df = _tsv.read_table(kmers_fpath, index_col="contig")
    except (ValueError, TypeError):
        raise TableFormatError(f"contig column not found in {kmers_fpath}!")
    return df


def _counts_frame(ids, counts: np.ndarray, countable: np.ndarray, names) -> pd.DataFrame:
    """Frame with the reference's schema (kmers.py:205,325-326): index 'contig',
    k-mer columns, zero counts stored as ``pd.NA``.  Nullable Int32 columns are
    used instead of object dtype; contigs with L <= k are all-NA rows."""
    # duplicate ids: the reference merges per-record dicts with dict.update (kmers.py:155-157),
    # so a later record overwrites the values but keeps the first record's position
    index = pd.Index(ids, name="contig")
    countable = np.ascontiguousarray(countable, dtype=bool)
    if not index.is_unique:
        first, last = {}, {}
        for pos, cid in enumerate(ids):
            first.setdefault(cid, pos)
            last[cid] = pos
        keep = sorted(first.values())
        src = [last[ids[p]] for p in keep]
        counts, countable = counts[src], countable[src]
        index = pd.Index([ids[p] for p in keep], name="contig")
    # one blocked, threaded pass (am_counts_to_columns) instead of a mask pass + 2 x 512 strided column gathers
    # (2.4 s -> 0.2 s at 200k x 512): row j of the transposed matrices is column j's value / mask array
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    n, D = counts.shape
    values_t = np.empty((D, n), dtype=np.int32)
    mask_t = np.empty((D, n), dtype=bool)
    L.check(L.lib.am_counts_to_columns(L.np_ptr(counts), n, D, L.np_ptr(countable), L.np_ptr(values_t),
                                       L.np_ptr(mask_t), 0))
    cols = {name: pd.arrays.IntegerArray(values_t[j], mask_t[j]) for j, name in enumerate(names)}
    return pd.DataFrame(cols, index=index, copy=False)


def count(
    assembly: str,
    size: int = 5,
    out: str = None,
    force: bool = False,
    verbose: bool = True,
    cpus: int = mp.cpu_count(),
    device=None,
) -> pd.DataFrame:
    """Count canonical k-mers of every contig of ``assembly`` (reference kmers.py:268-330).

    ``cpus`` only sizes the host packing threads; the counting runs on the GPU.
    Returns index='contig' (FASTA order), columns = :func:`init_kmers` order, zero ->
    ``pd.NA``; contigs with ``len <= size`` give an all-NA row and a warning
    (kmers.py:187-192).  ``out``/``force`` caching as in the reference (:310-314,327-329).
    """
    out_specified = out is not None
    out_exists = os.path.exists(out) if out else False
    if out_specified and out_exists and not force:
        logger.warning(f"counts already exist: {out} force to overwrite. [retrieving]")
        df = _tsv.read_table(out, index_col="contig")
    else:
        if not isinstance(size, int) or isinstance(size, bool):
            raise TypeError(f"size must be an int! Given: {type(size)}")
        names = _dev.column_names(size)
        logger.info(f"Counting {size}-mers.")
        host = _asm.read_fasta(assembly)
        packed = _asm.pack_to_device(host, device=device, threads=int(cpus) if cpus else 0)
        counts_d, _, _ = _dev.count_packed(packed, size)
        counts = counts_d.cpu().numpy()
        lengths = host.lengths
        countable = lengths > size
        for cid, ln in zip(np.asarray(host.ids, dtype=object)[~countable], lengths[~countable]):
            logger.warning(f"{cid} can not be counted! k-mer size exceeds length. {ln}")
        df = _counts_frame(host.ids, counts, countable, names)
        if out and df.shape[0] == len(host.ids):  # no duplicate ids: write straight from the count matrix
            _tsv.write_counts(out, host.ids, names, counts)
            logger.debug(f"Wrote {df.shape[0]} contigs {size}-mers frequencies to {out}.")
            return df
    if out:
        _tsv.write_frame(df, out)
        logger.debug(f"Wrote {df.shape[0]} contigs {size}-mers frequencies to {out}.")
    return df


def _nullable_int_matrix(df: pd.DataFrame):
    """``_counts_matrix`` for the frame :func:`count` returns (every column a nullable integer array): values
    and masks are stacked column by column and transposed once, instead of going through an object / float64
    conversion of the whole frame (2.4 s -> 0.2 s at 200k x 512; am_columns_to_counts).  ``None`` for any other frame."""
    if df.shape[1] == 0 or not df.columns.is_unique:
        return None
    arrays = [df[c].array for c in df.columns]
    if not all(isinstance(a, pd.arrays.IntegerArray) for a in arrays):
        return None
    n, D = df.shape
    itemsizes = {a._data.dtype.itemsize for a in arrays}
    if itemsizes not in ({4}, {8}) or any(a._data.dtype.kind != "i" for a in arrays):
        return None
    import ctypes as C

    vals = [np.ascontiguousarray(a._data) for a in arrays]
    masks = [np.ascontiguousarray(a._mask) for a in arrays]
    pv = (C.c_void_p * D)(*[v.ctypes.data for v in vals])
    pm = (C.c_void_p * D)(*[m.ctypes.data for m in masks])
    counts = np.empty((n, D), dtype=np.int32)
    isna = np.empty((n, D), dtype=bool)
    bad = C.c_int64(0)
    L.check(L.lib.am_columns_to_counts(C.addressof(pv), C.addressof(pm), itemsizes.pop(), n, D, L.np_ptr(counts),
                                       L.np_ptr(isna), C.addressof(bad), 0))
    if bad.value:
        raise TableFormatError("k-mer counts must be non-negative integers")
    return counts, isna


def _counts_matrix(df: pd.DataFrame):
    """(int32 counts with NA -> 0, isna mask) from an object/NA, nullable-int or float/NaN frame."""
    fast = _nullable_int_matrix(df)
    if fast is not None:
        return fast
    try:
        X = df.to_numpy(dtype=np.float64, na_value=np.nan)
    except (ValueError, TypeError):
        raise TableFormatError("k-mer table holds non-numeric values")
    isna = np.isnan(X)
    X = np.where(isna, 0.0, X)
    if X.size and (np.any(X < 0) or np.any(X > 2**31 - 1) or np.any(X != np.rint(X))):
        raise TableFormatError("k-mer counts must be non-negative integers")
    return np.ascontiguousarray(X.astype(np.int32)), isna


def _normalize_array(counts: np.ndarray, isna: np.ndarray, method: str, device=None):
    """Run kernel 2 on a host count matrix; returns (values, row_keep, col_keep)."""
    import torch

    dev = L.require_cuda(device)
    n, D = counts.shape
    if method == "am_clr":
        # kmers.py:369: drop all-NA columns, then all-NA rows, then fill with 0
        col_keep = ~isna.all(axis=0) if n else np.zeros(D, dtype=bool)
        row_keep = ~isna[:, col_keep].all(axis=1) if col_keep.any() else np.zeros(n, dtype=bool)
    else:
        col_keep = np.ones(D, dtype=bool)
        row_keep = np.ones(n, dtype=bool)
        if n and np.any(counts.sum(axis=1) == 0):
            # skbio closure() inside multiplicative_replacement (kmers.py:420)
            raise ValueError("Input matrix cannot have rows with all zeros")
    n_out, d_keep = int(row_keep.sum()), int(col_keep.sum())
    d_out = d_keep - 1 if method == "ilr" else d_keep
    if n_out == 0 or d_keep == 0 or d_out <= 0:
        return np.zeros((n_out, max(d_out, 0))), row_keep, col_keep
    with torch.cuda.device(dev):
        d_counts = torch.from_numpy(counts).to(dev)
        ridx = None if row_keep.all() else torch.from_numpy(np.nonzero(row_keep)[0].astype(np.int64)).to(dev)
        cidx = None if col_keep.all() else torch.from_numpy(np.nonzero(col_keep)[0].astype(np.int32)).to(dev)
        out = _dev.normalize_counts(d_counts, method, ridx, cidx)
        return out.cpu().numpy(), row_keep, col_keep


def autometa_clr(df: pd.DataFrame, device=None) -> pd.DataFrame:
    """Autometa's CLR (reference kmers.py:333-373): drop k-mers absent from every
    contig and contigs without counts, then ``log((x+1)/sum / gmean((x+1)/sum))``."""
    counts, isna = _counts_matrix(df)
    vals, row_keep, col_keep = _normalize_array(counts, isna, "am_clr", device)
    return pd.DataFrame(vals, index=df.index[row_keep], columns=df.columns[col_keep])


def normalize(
    df: pd.DataFrame, method: str = "am_clr", out: str = None, force: bool = False, device=None
) -> pd.DataFrame:
    """Normalise raw k-mer counts by ``am_clr``, ``clr`` or ``ilr`` (reference kmers.py:376-432)."""
    method = method.lower()
    out_specified = out is not None
    out_exists = os.path.exists(out) if out else False
    if out_specified and out_exists and not force:
        logger.debug(f"{out} already exists. Use force to overwrite. retrieving...")
        return _tsv.read_table(out, index_col="contig")
    logger.debug(f"Transforming k-mer counts using {method}")
    choices = {"ilr", "clr", "am_clr"}
    if method == "am_clr":
        counts, isna = _counts_matrix(df)
        vals, row_keep, col_keep = _normalize_array(counts, isna, "am_clr", device)
        norm_df = pd.DataFrame(vals, index=df.index[row_keep], columns=df.columns[col_keep])  # = autometa_clr(df)
    elif method in choices:
        counts, isna = _counts_matrix(df)
        vals, _, _ = _normalize_array(counts, isna, method, device)
        norm_df = pd.DataFrame(vals, index=df.index)  # integer column labels (kmers.py:422)
    else:
        raise ValueError(f"Normalize Method not available! {method}. choices: {', '.join(choices)}")
    if out_specified and (force or not out_exists):
        _tsv.write_frame(norm_df, out, values=vals)  # the matrix the frame was built from: no transposing copy
    return norm_df


def pca(X, n_components: int = 50, seed: int = 42, device=None, as_numpy: bool = True) -> Dict[str, Any]:
    """PCA scores of ``X`` [N x D] exactly as sklearn 1.9's
    ``PCA(n_components, svd_solver="covariance_eigh").fit_transform`` computes them
    (the branch reference kmers.py:605 takes when N >= 10*D; sklearn/decomposition/_pca.py:560,604-646):
    mean, covariance from the centred Gram, symmetric eigensolve, descending order,
    negative eigenvalues clipped, svd_flip sign rule, scores = (X - mean) @ components.T.
    ``seed`` is accepted for signature parity; this exact solver uses no randomness.
    """
    import torch

    dev = L.require_cuda(device)
    with torch.cuda.device(dev):
        if isinstance(X, torch.Tensor):
            Xd = X.to(device=dev, dtype=torch.float64).contiguous()
        else:
            Xd = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float64)).to(dev)
        n, D = Xd.shape
        if not 1 <= n_components <= min(n, D):
            raise ValueError(
                f"n_components={n_components!r} must be between 1 and min(n_samples, n_features)={min(n, D)}"
            )
        res = pca_device(Xd, n_components)
    if as_numpy:
        res = {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in res.items()}
    return res


def pca_device(Xd, n_components: int, group=None) -> Dict[str, Any]:
    """PCA on a device-resident fp64 matrix (rows may be sharded over the ranks of
    ``group``: column sums, row count and the Gram are all-reduced)."""
    import torch

    n_local, D = Xd.shape
    cs, cm = _dev.colstats(Xd)
    n_t = torch.tensor([float(n_local)], dtype=torch.float64, device=Xd.device)
    if group is not None:
        import torch.distributed as dist

        dist.all_reduce(cs, group=group if group is not True else None)
        dist.all_reduce(n_t, group=group if group is not True else None)
    n = n_t  # stays on device: no host sync in the fit
    mu = cs / n
    # tcgen05 int8 digit-plane product (int32 accumulation is exact; digit pairs of weight >= 7 are dropped:
    # <= 2^-48 of (column range)^2 per term, on top of the 2^-46 quantisation); the planes are reused by the projection
    e, sc = _dev.plane_scales(mu, cm)
    planes = _dev.quantize_planes(Xd, mu, sc)
    G = _dev.gram_planes(planes, n_local, D, e)
    if group is not None:
        import torch.distributed as dist

        dist.all_reduce(G, group=group if group is not True else None)
    C = G / (n - 1.0)
    # only the leading n_components pairs are solved; total variance = trace (sum of all eigenvalues)
    evals, comps, total = _dev.eigh_topk(C, n_components)
    evals = torch.clamp(evals, min=0.0)  # _pca.py:630
    total = total.reshape(())
    if n_components <= 64:
        scores = _dev.project_planes(planes, n_local, D, e, mu, mu, comps)
    else:
        scores = _dev.project(Xd, mu, comps)
    return dict(
        scores=scores,
        mean=mu,
        components=comps,
        explained_variance=evals.clone(),
        explained_variance_ratio=evals / total,
        total_variance=total,
        covariance=C,
    )


def embed(
    kmers: Union[str, pd.DataFrame],
    out: str = None,
    force: bool = False,
    embed_dimensions: int = 2,
    pca_dimensions: int = 50,
    method: str = "bhsne",
    perplexity: float = 30.0,
    seed: int = 42,
    n_jobs: int = -1,
    device=None,
    **method_kwargs: Dict[str, Any],
) -> pd.DataFrame:
    """Embed normalised k-mer frequencies (reference kmers.py:435-706).

    Only the PCA stage (kmers.py:588-607) runs on the GPU here; the manifold step
    is handed to whichever embedder is installed (sklearn's TSNE for ``sksne``;
    ``tsne``/``umap``/``trimap`` packages for the others) with the reference's arguments.
    """
    if isinstance(kmers, str) and os.path.exists(kmers) and os.path.getsize(kmers):
        try:
            df = _tsv.read_table(kmers, index_col="contig")
        except ValueError:
            raise TableFormatError(f"contig column not found in {kmers}")
    elif isinstance(kmers, pd.DataFrame):
        df = kmers
    else:
        raise TypeError(kmers)
    if out and os.path.exists(out) and os.path.getsize(out) and not force:
        logger.debug(f"k-mers frequency embedding already exists {out}")
        try:
            return _tsv.read_table(out, index_col="contig")
        except ValueError:
            raise TableFormatError(f"contig column not found in {out}")
    if df.empty:
        raise FileNotFoundError(
            f"kmers:{kmers} type:{type(kmers)} out:{out} type:{type(out)} Given pd.DataFrame is empty!"
        )
    method = method.lower()
    choices = {"umap", "sksne", "bhsne", "densmap", "trimap"}
    if method not in choices:
        raise ValueError(f"{method} not in embedding methods. Choices: {', '.join(choices)}")
    n_samples, n_components = df.shape
    X = df.dropna(axis="index", how="all").fillna(0).to_numpy(dtype=np.float64)
    random_state = np.random.RandomState(seed)
    if isinstance(pca_dimensions, str):
        try:
            pca_dimensions = int(pca_dimensions)
        except Exception:
            raise TypeError(f"pca_dimensions must be an integer! given: {pca_dimensions}")
    if n_components > pca_dimensions and pca_dimensions != 0:
        logger.debug(f"Performing decomposition with PCA (seed {seed}): {n_components} to {pca_dimensions} dims")
        X = pca(X, n_components=pca_dimensions, seed=seed, device=device)["scores"]
        n_samples, n_components = X.shape
    logger.debug(f"{method}: {n_samples} data points and {n_components} dimensions")
    n_rows = n_samples - 1
    scaler = 3.0
    if n_rows < (scaler * perplexity):
        perplexity = (n_rows / scaler) - 1

    def do_sksne():
        from sklearn.manifold import TSNE

        return TSNE(
            n_components=embed_dimensions, perplexity=perplexity, random_state=random_state,
            n_jobs=n_jobs, **method_kwargs,
        ).fit_transform(X)

    def do_bhsne():
        from tsne import bh_sne  # not part of this package: passed through when installed

        return bh_sne(data=X, d=embed_dimensions, perplexity=perplexity, random_state=random_state, **method_kwargs)

    def do_umap():
        from umap import UMAP

        return UMAP(
            n_neighbors=15, n_components=embed_dimensions, metric="euclidean", random_state=random_state,
            densmap=(method == "densmap"), n_jobs=n_jobs, **method_kwargs,
        ).fit_transform(X)

    def do_trimap():
        from trimap import TRIMAP

        return TRIMAP(n_dims=embed_dimensions, verbose=False, **method_kwargs).fit_transform(X)

    dispatcher = {"sksne": do_sksne, "bhsne": do_bhsne, "umap": do_umap, "densmap": do_umap, "trimap": do_trimap}
    logger.debug(f"Performing embedding with {method} (seed {seed})")
    try:
        X = dispatcher[method]()
    except ValueError as err:
        if method == "sksne":
            logger.warning(f"embed_dimensions ({embed_dimensions}) is too high for sksne. Reducing to 3.")
            embed_dimensions = 3
            X = dispatcher[method]()
        else:
            raise err
    embed_cols = [f"x_{col}" for col in range(1, embed_dimensions + 1)]
    if isinstance(X, tuple):
        cols = [embed_cols, ["original_local_radius"], ["embedded_local_radius"]]
        embedded_df = pd.concat(
            [pd.DataFrame(res, index=df.index, columns=c) for res, c in zip(X, cols)], axis=1
        )
    else:
        embedded_df = pd.DataFrame(X, index=df.index, columns=embed_cols)
    if out:
        _tsv.write_frame(embedded_df, out)
        logger.debug(f"embedded.shape {embedded_df.shape} : Written {out}")
    return embedded_df
