"""CPU oracle: a numpy restatement of Autometa's featurise + cluster hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline legs may import this module; the product package
``autometa_b200`` never does (it fails loudly when the CUDA library is missing).

Every function cites the reference lines it restates (paths relative to the
reference checkout; ``sklearn/`` = site-packages/sklearn 1.9.0).  Pinning: the
reference ships no golden vectors for this path (tests/data is not in the tree,
SURVEY.md fact 10), so the oracle is pinned against the reference code itself,
run through ``oracle/ref_shim.py`` in the build container; the outputs are
committed under ``tests/golden/`` by ``tests/golden/make_golden.py`` and
``tests/test_oracle.py`` re-checks this module against them everywhere.
"""
from __future__ import annotations

import gzip
from typing import Dict, List, Tuple

import numpy as np

# --------------------------------------------------------------------------- #
# k-mer column order                                                          #
# --------------------------------------------------------------------------- #
#: letter order used by the reference when it enumerates k-mers
#: (autometa/common/kmers.py:75 ``dna_letters = ["A","T","C","G"]``).
LETTERS = "ATCG"
_ENC = {c: i for i, c in enumerate(LETTERS)}


def revcomp_code(code: int, k: int) -> int:
    """Reverse complement in the A=0,T=1,C=2,G=3 code: reverse digits, XOR 1.

    Restates ``_revcomp`` (kmers.py:38-53): complement A<->T, C<->G is ``d ^ 1``.
    """
    rc = 0
    for _ in range(k):
        rc = (rc << 2) | ((code & 3) ^ 1)
        code >>= 2
    return rc


def kmer_columns(k: int = 5) -> Tuple[List[str], np.ndarray]:
    """Column names and the code -> column LUT of ``init_kmers`` (kmers.py:56-89).

    The reference enumerates all 4**k strings in base-4 order over A,T,C,G
    (kmers.py:76-82) and keeps a string iff neither it nor its reverse
    complement was kept before (kmers.py:84-88).  Enumeration order = ascending
    code, so the kept set is {min(code, rc(code))} and the column index is the
    rank of the canonical code.
    """
    if not isinstance(k, int):
        raise TypeError(f"kmer_size must be an int! Given: {type(k)}")
    n = 4 ** k
    codes = np.arange(n, dtype=np.int64)
    rc = np.array([revcomp_code(int(c), k) for c in codes], dtype=np.int64)
    canon = np.minimum(codes, rc)
    uniq = np.unique(canon)
    lut = np.searchsorted(uniq, canon).astype(np.int32)
    names = []
    for c in uniq:
        s = []
        c = int(c)
        for i in range(k):
            s.append(LETTERS[(c >> (2 * (k - 1 - i))) & 3])
        names.append("".join(s))
    return names, lut


_BYTE2CODE = np.full(256, 255, dtype=np.uint8)
for _c, _i in _ENC.items():
    _BYTE2CODE[ord(_c)] = _i  # upper-case ACGT only (kmers.py:197-204, dict keys)


def count_sequence(seq: bytes, k: int, lut: np.ndarray, ncols: int) -> Tuple[np.ndarray, bool]:
    """Canonical k-mer histogram of one contig.

    Restates ``record_counter`` (kmers.py:161-206): ``max_length = L - k``
    (:186), ``for i in range(max_length)`` (:193) -> only the first L-k windows
    (the last k-mer is never counted); a window counts iff the k-mer or its
    reverse complement is a dict key (:198-204), i.e. iff all k bytes are
    upper-case A/C/G/T.  Returns (counts[int64, ncols], countable) where
    ``countable`` is False when ``L <= k`` (row becomes all-NA, :187-192).
    """
    L = len(seq)
    out = np.zeros(ncols, dtype=np.int64)
    nwin = L - k
    if nwin <= 0:
        return out, False
    codes = _BYTE2CODE[np.frombuffer(seq, dtype=np.uint8)]
    bad = codes == 255
    c = np.where(bad, 0, codes).astype(np.int64)
    code = np.zeros(nwin, dtype=np.int64)
    anybad = np.zeros(nwin, dtype=bool)
    for j in range(k):
        code = (code << 2) | c[j : j + nwin]
        anybad |= bad[j : j + nwin]
    cols = lut[code[~anybad]]
    out += np.bincount(cols, minlength=ncols)
    return out, True


def read_fasta(path: str) -> Tuple[List[str], List[bytes]]:
    """Minimal FASTA reader with Biopython ``SeqIO.parse(..., "fasta")`` semantics
    for the fields the reference uses (kmers.py:143-148): id = first token of the
    header, sequence = concatenated lines, case preserved."""
    opener = gzip.open if path.endswith(".gz") else open
    ids, seqs, buf, cur = [], [], [], None
    with opener(path, "rb") as fh:
        for line in fh:
            line = line.rstrip(b"\r\n")
            if line.startswith(b">"):
                if cur is not None:
                    ids.append(cur)
                    seqs.append(b"".join(buf))
                tok = line[1:].split()
                cur = tok[0].decode() if tok else ""
                buf = []
            else:
                # Biopython: line.rstrip() per line; spaces and CRs removed from the joined string
                buf.append(line.rstrip().replace(b" ", b"").replace(b"\r", b""))
        if cur is not None:
            ids.append(cur)
            seqs.append(b"".join(buf))
    return ids, seqs


def count_sequences(seqs: List[bytes], k: int = 5) -> Tuple[np.ndarray, np.ndarray]:
    """counts[N, ncols] int64 and countable[N] bool for a list of contigs."""
    names, lut = kmer_columns(k)
    ncols = len(names)
    counts = np.zeros((len(seqs), ncols), dtype=np.int64)
    ok = np.zeros(len(seqs), dtype=bool)
    for i, s in enumerate(seqs):
        counts[i], ok[i] = count_sequence(s, k, lut, ncols)
    return counts, ok


def counts_to_float_frame(counts: np.ndarray) -> np.ndarray:
    """What ``kmers.load`` (kmers.py:116) gives back after ``count`` wrote its TSV
    (kmers.py:205,328): zeros were stored as NA -> empty cell -> NaN."""
    X = counts.astype(np.float64)
    X[counts == 0] = np.nan
    return X


# --------------------------------------------------------------------------- #
# normalisation                                                               #
# --------------------------------------------------------------------------- #
def am_clr(X: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """``autometa_clr`` (kmers.py:333-373) on a float/NaN matrix.

    Returns (normalised[N', D'], row_keep[N] bool, col_keep[D] bool).
    :369 drops columns that are NaN in every row, then rows that are NaN in
    every (remaining) column, then fills NaN with 0; :371 ``(x+1)/x.sum()``;
    :372 ``log(y / gmean(y))`` with scipy's gmean = exp(mean(log y)).
    The arithmetic below follows those two steps literally (not the simplified
    log1p form) so that rounding matches the reference to the last ulps.
    """
    col_keep = ~np.all(np.isnan(X), axis=0)
    Xc = X[:, col_keep]
    row_keep = ~np.all(np.isnan(Xc), axis=1)
    Xr = np.nan_to_num(Xc[row_keep], nan=0.0)
    y = (Xr + 1.0) / Xr.sum(axis=1, keepdims=True)
    g = np.exp(np.mean(np.log(y), axis=1, keepdims=True))
    return np.log(y / g), row_keep, col_keep


def multiplicative_replacement(X: np.ndarray) -> np.ndarray:
    """scikit-bio ``multiplicative_replacement`` (published definition; called at
    kmers.py:420 on ``df.fillna(0)``): closure, zeros -> delta = 1/D**2,
    non-zeros scaled by (1 - n_zeros*delta)."""
    X = np.asarray(X, dtype=np.float64)
    s = X.sum(axis=1, keepdims=True)
    if np.any(s == 0):
        raise ValueError("Input matrix cannot have rows with all zeros")
    p = X / s
    z = p == 0
    D = X.shape[1]
    delta = (1.0 / D) ** 2
    return np.where(z, delta, (1 - z.sum(axis=1, keepdims=True) * delta) * p)


def clr(X: np.ndarray) -> np.ndarray:
    """scikit-bio ``clr`` applied after multiplicative replacement (kmers.py:418-421)."""
    q = multiplicative_replacement(X)
    q = q / q.sum(axis=1, keepdims=True)
    l = np.log(q)
    return l - l.mean(axis=1, keepdims=True)


def ilr_basis(D: int) -> np.ndarray:
    """Default scikit-bio ilr basis (``_gram_schmidt_basis``): row j (i=j+1) is
    sqrt(i/(i+1)) * [1/i]*i + [-1] + [0]*(D-i-1).  Shape [D-1, D]."""
    B = np.zeros((D - 1, D))
    for j in range(D - 1):
        i = j + 1
        B[j, :i] = 1.0 / i
        B[j, i] = -1.0
        B[j] *= np.sqrt(i / (i + 1.0))
    return B


def ilr(X: np.ndarray) -> np.ndarray:
    """scikit-bio ``ilr`` = clr(x) @ basis.T (kmers.py:418-421)."""
    return clr(X) @ ilr_basis(X.shape[1]).T


# --------------------------------------------------------------------------- #
# PCA                                                                         #
# --------------------------------------------------------------------------- #
def pca_cov_eigh(X: np.ndarray, P: int = 50) -> Dict[str, np.ndarray]:
    """sklearn 1.9 ``PCA._fit_full`` with ``svd_solver="covariance_eigh"``
    (sklearn/decomposition/_pca.py:560,604-646), which is what kmers.py:605 runs
    when N >= 10*D; then ``fit_transform`` scores (sklearn/decomposition/_base.py:151-159
    via _pca.py fit_transform -> U*S path is NOT used for covariance_eigh; it
    calls ``_transform`` = X @ Vt.T - mean @ Vt.T).
    """
    n, D = X.shape
    mean = X.mean(axis=0)
    C = X.T @ X
    C -= n * np.outer(mean, mean)
    C /= n - 1
    w, V = np.linalg.eigh(C)
    w = w[::-1].copy()
    V = V[:, ::-1]
    w[w < 0.0] = 0.0
    total = w.sum()
    comps = V.T.copy()
    # svd_flip(u_based_decision=False) (sklearn/utils/extmath.py:973-981)
    idx = np.argmax(np.abs(comps), axis=1)
    signs = np.sign(comps[np.arange(D), idx])
    comps *= signs[:, None]
    comps = comps[:P]
    scores = X @ comps.T - (mean.reshape(1, -1) @ comps.T)
    return dict(
        mean=mean,
        explained_variance=w[:P].copy(),
        explained_variance_ratio=w[:P] / total,
        components=comps,
        scores=scores,
        total_variance=total,
    )


# --------------------------------------------------------------------------- #
# DBSCAN(min_samples=1) = connected components of the eps graph               #
# --------------------------------------------------------------------------- #
def _rdist_le(a: np.ndarray, b: np.ndarray, eps: float) -> np.ndarray:
    """sklearn KD-tree predicate (sklearn/neighbors/_binary_tree.pxi.tp:1952-1957 with
    ``euclidean_rdist``): d = sum_j (a_j - b_j)**2 accumulated in column order in
    fp64, compared ``<= eps*eps`` (inclusive)."""
    d = np.zeros(a.shape[0], dtype=np.float64)
    for j in range(a.shape[1]):
        t = a[:, j] - b[:, j]
        d = d + t * t
    return d <= eps * eps


def eps_edges(X: np.ndarray, eps: float) -> Tuple[np.ndarray, np.ndarray]:
    """All pairs i<j with rdist <= eps**2, found with a uniform cell list."""
    n, F = X.shape
    cell = np.floor((X - X.min(axis=0)) / eps).astype(np.int64)
    dims = cell.max(axis=0) + 1
    # hash cells to sort points
    key = np.zeros(n, dtype=np.int64)
    for j in range(F):
        key = key * int(dims[j] + 2) + (cell[:, j] + 1)
    order = np.argsort(key, kind="stable")
    skey = key[order]
    uniq, start = np.unique(skey, return_index=True)
    end = np.append(start[1:], n)
    ii, jj = [], []
    offs = np.array(np.meshgrid(*[[-1, 0, 1]] * F, indexing="ij")).reshape(F, -1).T
    mult = np.ones(F, dtype=np.int64)
    for j in range(F - 2, -1, -1):
        mult[j] = mult[j + 1] * int(dims[j + 1] + 2)
    for off in offs:
        dk = int((off * mult).sum())
        if dk < 0:
            continue
        tgt = uniq + dk
        pos = np.searchsorted(uniq, tgt)
        ok = (pos < len(uniq)) & (uniq[np.minimum(pos, len(uniq) - 1)] == tgt)
        for ca, cb in zip(np.nonzero(ok)[0], pos[ok]):
            A = order[start[ca] : end[ca]]
            B = order[start[cb] : end[cb]]
            if dk == 0:
                a, b = np.triu_indices(len(A), k=1)
                pa, pb = A[a], A[b]
            else:
                pa = np.repeat(A, len(B))
                pb = np.tile(B, len(A))
            if len(pa) == 0:
                continue
            m = _rdist_le(X[pa], X[pb], eps)
            ii.append(pa[m])
            jj.append(pb[m])
    if not ii:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    return np.concatenate(ii), np.concatenate(jj)


def labels_from_edges(n: int, i: np.ndarray, j: np.ndarray) -> np.ndarray:
    """Connected-component labels numbered by ascending minimum member index —
    the order sklearn's ``dbscan_inner`` (sklearn/cluster/_dbscan_inner.pyx) produces
    when every point is a core point (min_samples=1, recursive_dbscan.py:109)."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components

    g = coo_matrix((np.ones(len(i), dtype=np.int8), (i, j)), shape=(n, n))
    _, lab = connected_components(g, directed=False)
    # relabel by first occurrence
    _, first = np.unique(lab, return_index=True)
    rank = np.empty(len(first), dtype=np.int64)
    rank[np.argsort(first)] = np.arange(len(first))
    return rank[lab].astype(np.int64)


def dbscan_labels(X: np.ndarray, eps: float) -> np.ndarray:
    """``DBSCAN(eps, min_samples=1).fit(X).labels_`` (recursive_dbscan.py:109)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    i, j = eps_edges(X, eps)
    return labels_from_edges(X.shape[0], i, j)


def dbscan_labels_bruteforce(X: np.ndarray, eps: float) -> np.ndarray:
    """O(n^2) version for small n (used to validate the cell list)."""
    n = X.shape[0]
    a, b = np.triu_indices(n, k=1)
    m = _rdist_le(X[a], X[b], eps)
    return labels_from_edges(n, a[m], b[m])


def eps_schedule(n_values: int, eps0: float = 0.3, step: float = 0.1) -> List[float]:
    """The reference's float accumulation ``eps = 0.3; eps += 0.1`` (recursive_dbscan.py:172-173,215)."""
    out, e = [], eps0
    for _ in range(n_values):
        out.append(e)
        e += step
    return out


# --------------------------------------------------------------------------- #
# per-eps cluster metrics (binning/utilities.py:125-262)                       #
# --------------------------------------------------------------------------- #
def cluster_metrics(labels, markers, coverage, gc, n_ref_markers):
    """Per-cluster completeness / purity / coverage sigma / GC sigma
    (binning/utilities.py:176-192): marker sums per cluster, ``==1`` single copy,
    ``>=1`` present, completeness = present / M * 100, purity = single / present * 100
    (NaN when present == 0), std with ddof=1 (NaN for singleton clusters)."""
    nclu = int(labels.max()) + 1 if len(labels) else 0
    M = markers.shape[1]
    sums = np.zeros((nclu, M), dtype=np.int64)
    np.add.at(sums, labels, markers)
    single = (sums == 1).sum(axis=1)
    present = (sums >= 1).sum(axis=1)
    completeness = present / n_ref_markers * 100
    with np.errstate(invalid="ignore", divide="ignore"):
        purity = single / present * 100
    cnt = np.bincount(labels, minlength=nclu).astype(np.float64)

    def _std(v):
        s = np.bincount(labels, weights=v, minlength=nclu)
        mean = s / cnt
        d = v - mean[labels]
        ss = np.bincount(labels, weights=d * d, minlength=nclu)
        with np.errstate(invalid="ignore", divide="ignore"):
            return np.sqrt(ss / (cnt - 1))

    return dict(
        completeness=completeness,
        purity=purity,
        coverage_stddev=_std(coverage),
        gc_content_stddev=_std(gc),
        present_marker_count=present,
        single_copy_marker_count=single,
    )


# --------------------------------------------------------------------------- #
# FASTA-ingest by-products                                      #
# --------------------------------------------------------------------------- #
def gc_fraction_remove(seq: bytes) -> float:
    """``Bio.SeqUtils.gc_fraction(seq)`` with its default ``ambiguous="remove"`` as used by
    autometa/common/metagenome.py:139,314.  Biopython (>=1.82 per autometa-env.yml:9) is
    not installed in the build container, so this restates its published definition:
    gc = count of the letters "CGScgs", length = gc + count of "ATWUatwu", 0 for length 0,
    else gc / length.  PARITY UNPINNED against a live Biopython; checked only against the
    values its docstring documents (ACTG -> 0.5, ACTGN -> 0.5, ACTGSSSS -> 0.75)."""
    gc = sum(seq.count(x) for x in b"CGScgs")
    length = gc + sum(seq.count(x) for x in b"ATWUatwu")
    if length == 0:
        return 0
    return gc / length


def gc_content_table(seqs: List[bytes]) -> Tuple[np.ndarray, np.ndarray]:
    """(gc_content %, length) per record: Metagenome.gc_content, metagenome.py:303-320."""
    gc = np.array([gc_fraction_remove(s) * 100 for s in seqs], dtype=np.float64)
    return gc, np.array([len(s) for s in seqs], dtype=np.int64)


def length_weighted_gc(seqs: List[bytes]) -> float:
    """Metagenome.length_weighted_gc, metagenome.py:124-144."""
    size = sum(len(s) for s in seqs)
    weights = [len(s) / size for s in seqs]
    return float(np.average(a=[gc_fraction_remove(s) * 100 for s in seqs], weights=weights))


def fragmentation_metric(lengths, quality_measure: float = 0.5):
    """Metagenome.fragmentation_metric, metagenome.py:177-208."""
    target = sum(lengths) * quality_measure
    total = 0
    for length in sorted(lengths, reverse=True):
        total += length
        if total >= target:
            return length
