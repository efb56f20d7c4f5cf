"""Import shim that loads the UNMODIFIED reference modules from /root/reference.

TEST INFRASTRUCTURE ONLY.  Nothing under ``autometa_b200/`` may import this.
It exists so that (a) the numpy/C oracle in this directory can be pinned against
the reference's own code and (b) ``tests/golden/make_golden.py`` can generate the
committed golden vectors.  ``/root/reference`` does not exist on the GPU box, so
nothing that runs there (``-m gpu`` tests, ``smoke()``, ``bench.py``) imports it.

The shim contains no reference code: it fakes the ``autometa`` package object
(``autometa/__init__.py:3-8`` calls pkg_resources and fails unless pip-installed)
and the third-party modules that are absent from this image (Biopython,
scikit-bio, tsne, umap, trimap).  The Biopython stub restates the only behaviour
``autometa/common/kmers.py:145-148,194-197`` touches: FASTA parsing (id = first
header token, lines concatenated, case preserved), ``Seq`` slicing, ``len`` and
``reverse_complement`` (ACGT/acgt/N table; other letters never match the k-mer
dict so their complement is irrelevant).  The scikit-bio stub restates the
published definitions of ``multiplicative_replacement``/``clr``/``ilr``
(scikit-bio ``skbio/stats/composition.py``; parity for clr/ilr is therefore
"restated from the library's documentation").
"""
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("AUTOMETA_REFERENCE", "/root/reference")
REF_PKG = os.path.join(REF_ROOT, "autometa")


def available() -> bool:
    return os.path.isdir(REF_PKG)


class _Seq:
    _table = str.maketrans("ACGTacgtNn", "TGCAtgcaNn")

    def __init__(self, data):
        self._d = data

    def __len__(self):
        return len(self._d)

    def __getitem__(self, i):
        r = self._d[i]
        return _Seq(r) if isinstance(i, slice) else r

    def reverse_complement(self):
        return _Seq(self._d.translate(_Seq._table)[::-1])

    def __str__(self):
        return self._d


class _Record:
    def __init__(self, id_, seq):
        self.id = id_
        self.seq = _Seq(seq)


def _parse(handle, fmt):
    assert fmt == "fasta"
    close = isinstance(handle, str)
    fh = open(handle) if close else handle
    id_, buf = None, []
    for line in fh:
        line = line.rstrip("\n")
        if line.startswith(">"):
            if id_ is not None:
                yield _Record(id_, "".join(buf))
            tok = line[1:].split()
            id_ = tok[0] if tok else ""
            buf = []
        else:
            buf.append(line.strip())
    if id_ is not None:
        yield _Record(id_, "".join(buf))
    if close:
        fh.close()


# --- scikit-bio restatement (skbio.stats.composition) -------------------------
def _closure(mat):
    mat = np.atleast_2d(np.asarray(mat, dtype=np.float64))
    if np.any(mat < 0):
        raise ValueError("Cannot have negative proportions")
    if mat.ndim > 2:
        raise ValueError("Input matrix can only have two dimensions or less")
    if np.all(mat == 0, axis=1).sum() > 0:
        raise ValueError("Input matrix cannot have rows with all zeros")
    mat = mat / mat.sum(axis=1, keepdims=True)
    return mat.squeeze()


def _multiplicative_replacement(mat, delta=None):
    mat = _closure(mat)
    z_mat = mat == 0
    num_feats = mat.shape[-1]
    tot = z_mat.sum(axis=-1, keepdims=True)
    if delta is None:
        delta = (1.0 / num_feats) ** 2
    zcnts = 1 - tot * delta
    if np.any(zcnts) < 0:
        raise ValueError("The multiplicative replacement created negative proportions.")
    mat = np.where(z_mat, delta, zcnts * mat)
    return mat.squeeze()


def _clr(mat):
    mat = _closure(mat)
    lmat = np.log(mat)
    gm = lmat.mean(axis=-1, keepdims=True)
    return (lmat - gm).squeeze()


def _gram_schmidt_basis(n):
    basis = np.zeros((n, n - 1))
    for j in range(n - 1):
        i = j + 1.0
        e = np.array([(1.0 / i)] * int(i) + [-1] + [0] * int(n - i - 1)) * np.sqrt(i / (i + 1))
        basis[:, j] = e
    return basis.T


def _ilr(mat):
    mat = _closure(mat)
    basis = _gram_schmidt_basis(mat.shape[-1])
    return _clr(mat) @ basis.T


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_loaded = {}


def load():
    """Return (kmers, recursive_dbscan, binning_utilities) reference modules."""
    if _loaded:
        return _loaded["kmers"], _loaded["recursive_dbscan"], _loaded["utilities"]
    if not available():
        raise RuntimeError(f"reference not found at {REF_PKG}")
    pkg = types.ModuleType("autometa")
    pkg.__path__ = [REF_PKG]
    pkg.__version__ = "2.2.3"
    pkg.console_scripts = []
    sys.modules["autometa"] = pkg
    seqio = _stub("Bio.SeqIO", parse=_parse, SeqRecord=_Record)
    seqio.__path__ = []
    _stub("Bio", SeqIO=seqio)
    _stub("Bio.SeqIO.FastaIO", SimpleFastaParser=None)
    _stub("Bio.SeqRecord", SeqRecord=_Record)
    _stub("Bio.SeqUtils")
    _stub("skbio")
    _stub("skbio.stats")
    _stub(
        "skbio.stats.composition",
        ilr=_ilr,
        clr=_clr,
        multiplicative_replacement=_multiplicative_replacement,
    )
    _stub("tsne", bh_sne=lambda data, d, **kw: np.ascontiguousarray(data[:, :d]))
    _stub("umap", UMAP=None)
    _stub("trimap", TRIMAP=None)
    from autometa.common import kmers  # noqa: E402
    from autometa.binning import recursive_dbscan  # noqa: E402
    from autometa.binning import utilities  # noqa: E402

    _loaded.update(kmers=kmers, recursive_dbscan=recursive_dbscan, utilities=utilities)
    return kmers, recursive_dbscan, utilities
