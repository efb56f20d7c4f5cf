"""Fast TSV I/O for the k-mer tables.

``write_frame`` / ``write_counts`` / ``read_table`` replace
``df.to_csv(out, sep="\\t", index=True, header=True)`` and
``pd.read_csv(path, sep="\\t", index_col="contig")`` of the reference
(autometa/common/kmers.py:116,328,431,704) with the multi-threaded C routines
``am_tsv_*``: byte-identical files, identical frames.  ``.gz`` paths (the embedding
caches of large-data mode, large_data_mode.py:104,112) take the same fast formatter /
parser around a multi-member gzip stream compressed on all host threads (the decompressed
bytes are what pandas writes; any gzip reader, pandas included, reads the file).  Anything
outside the fast path (mixed dtypes, quoted labels, other compressions ...) goes through
pandas, so the result is the same either way.
"""
from __future__ import annotations

import gzip
import os
import tempfile
import zlib
from concurrent.futures import ThreadPoolExecutor
from typing import Optional, Sequence

import numpy as np
import pandas as pd

from . import _lib as L

_NEEDS_QUOTE = ("\t", '"', "\r", "\n")
#: spellings pandas turns into NaN when it parses an index column
_NA_LABELS = {"", "#N/A", "#N/A N/A", "#NA", "-1.#IND", "-1.#QNAN", "-NaN", "-nan", "1.#IND", "1.#QNAN", "<NA>",
              "N/A", "NA", "NULL", "NaN", "None", "n/a", "nan", "null"}


def _threads() -> int:
    return max(1, min(64, os.cpu_count() or 1))


def _cell(label) -> str:
    s = str(label)
    if any(ch in s for ch in _NEEDS_QUOTE):  # csv.QUOTE_MINIMAL with quotechar '"'
        return '"' + s.replace('"', '""') + '"'
    return s


def _labels_blob(labels: Sequence):
    labels = labels.tolist() if hasattr(labels, "tolist") else list(labels)
    if labels and all(type(x) is str for x in labels):
        joined = "".join(labels)
        if joined.isascii() and not any(ch in joined for ch in _NEEDS_QUOTE):
            # the common case (contig ids): no quoting, one byte per character — no per-label work beyond len()
            offsets = np.zeros(len(labels) + 1, dtype=np.int64)
            np.cumsum(np.fromiter(map(len, labels), dtype=np.int64, count=len(labels)), out=offsets[1:])
            blob = joined.encode("ascii")
            return np.frombuffer(blob, dtype=np.uint8) if blob else np.zeros(1, np.uint8), offsets
    cells = [_cell(x) for x in labels]
    blob = "".join(cells).encode("utf-8")
    lens = np.fromiter((len(c.encode("utf-8")) if not c.isascii() else len(c) for c in cells), dtype=np.int64,
                       count=len(cells))
    offsets = np.zeros(len(cells) + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    return np.frombuffer(blob, dtype=np.uint8) if blob else np.zeros(1, np.uint8), offsets


def _header(index_name, columns) -> bytes:
    first = "" if index_name is None else _cell(index_name)
    return ("\t".join([first] + [_cell(c) for c in columns]) + "\n").encode("utf-8")


def write_counts(path: str, ids: Sequence[str], names: Sequence[str], counts: np.ndarray,
                 index_name: str = "contig") -> None:
    """Write an int32 count matrix (0 = not observed = empty cell) as ``count`` would."""
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    n, d = counts.shape
    blob, offsets = _labels_blob(ids)
    hdr = _header(index_name, names)
    L.check(L.lib.am_tsv_write_i32(os.fsencode(path), hdr, len(hdr), L.np_ptr(blob), L.np_ptr(offsets), n, d,
                                   L.np_ptr(counts) if counts.size else None, _threads()))


_GZ_CHUNK = 8 << 20


def _gzip_members(raw: bytes, path: str, level: int = 6) -> None:
    """Write ``raw`` as a gzip file made of independent members of _GZ_CHUNK input bytes each, compressed in
    parallel (zlib releases the GIL).  Multi-member streams are standard gzip (RFC 1952 2.2)."""
    chunks = [raw[i : i + _GZ_CHUNK] for i in range(0, len(raw), _GZ_CHUNK)] or [b""]

    def one(c):
        co = zlib.compressobj(level, zlib.DEFLATED, 31)  # wbits 31: gzip container
        return co.compress(c) + co.flush()

    with ThreadPoolExecutor(max_workers=min(len(chunks), _threads())) as pool:
        parts = list(pool.map(one, chunks))
    with open(path, "wb") as fh:
        for part in parts:
            fh.write(part)


def write_frame(df: pd.DataFrame, path: str, values: Optional[np.ndarray] = None) -> None:
    """``df.to_csv(path, sep="\t", index=True, header=True)``; the fast writer when every column is float64
    (plain or ``.gz``).  ``values``: the C-contiguous float64 matrix ``df`` was built from, when the caller still
    holds it — a frame built from a row-major matrix hands back a column-major view, and turning that into rows
    again is a transposing copy of the whole table (0.5 s per 100k x 512)."""
    spath = str(path) if isinstance(path, (str, os.PathLike)) else ""
    simple = (
        bool(spath) and not spath.endswith((".bz2", ".zip", ".xz", ".zst"))
        and df.index.nlevels == 1 and df.columns.nlevels == 1 and df.shape[1] > 0
        and all(dt == np.float64 for dt in df.dtypes)
        and (df.index.dtype == object or pd.api.types.is_string_dtype(df.index.dtype))
        and not df.index.hasnans
    )
    if not simple:
        df.to_csv(path, sep="\t", index=True, header=True)
        return
    if (values is None or values.shape != df.shape or values.dtype != np.float64
            or not values.flags["C_CONTIGUOUS"]):
        values = np.ascontiguousarray(df.to_numpy(dtype=np.float64))
    blob, offsets = _labels_blob(df.index)
    hdr = _header(df.index.name, df.columns)
    target = spath
    tmp = None
    if spath.endswith(".gz"):
        fd, tmp = tempfile.mkstemp(suffix=".tsv", dir=os.path.dirname(os.path.abspath(spath)) or None)
        os.close(fd)
        target = tmp
    try:
        L.check(L.lib.am_tsv_write_f64(os.fsencode(target), hdr, len(hdr), L.np_ptr(blob), L.np_ptr(offsets),
                                       values.shape[0], values.shape[1], L.np_ptr(values) if values.size else None,
                                       _threads()))
        if tmp is not None:
            with open(tmp, "rb") as fh:
                _gzip_members(fh.read(), spath)
    finally:
        if tmp is not None and os.path.exists(tmp):
            os.remove(tmp)


def _looks_numeric(label: str) -> bool:
    try:
        float(label)
        return True
    except ValueError:
        return False


def read_table(path: str, index_col: str = "contig") -> pd.DataFrame:
    """``pd.read_csv(path, sep="\t", index_col=index_col)`` for a table of numbers (plain or ``.gz``)."""
    spath = str(path) if isinstance(path, (str, os.PathLike)) else ""
    fast = bool(spath) and not spath.endswith((".bz2", ".zip", ".xz", ".zst"))
    df = None
    if fast:
        if spath.endswith(".gz"):
            with open(spath, "rb") as fh:
                raw = gzip.decompress(fh.read())  # handles multi-member files
            buf = np.frombuffer(raw, dtype=np.uint8)
        else:
            buf = np.fromfile(spath, dtype=np.uint8)
        df = _read_fast(buf, index_col)
    if df is None:
        df = pd.read_csv(path, sep="\t", index_col=index_col)
    return df


def _read_fast(buf: np.ndarray, index_col: str) -> Optional[pd.DataFrame]:
    if buf.size == 0:
        return None
    # first newline: look at growing prefixes, not at a boolean image of the whole (GB-sized) file
    end, probe = buf.size, 1 << 16
    while True:
        hit = np.flatnonzero(buf[:probe] == 10)
        if hit.size:
            end = int(hit[0])
            break
        if probe >= buf.size:
            break
        probe <<= 4
    try:
        head = buf[:end].tobytes().decode("utf-8")
    except UnicodeDecodeError:
        return None
    head = head[:-1] if head.endswith("\r") else head
    fields = head.split("\t")
    if '"' in head or fields[0] != index_col or len(set(fields)) != len(fields) or len(fields) < 2:
        return None  # index column elsewhere, mangled duplicate names, quoting: pandas' job
    if any(f == "" or f.startswith("Unnamed") for f in fields):
        return None
    n_rows = np.zeros(1, np.int64)
    n_fields = np.zeros(1, np.int64)
    L.check(L.lib.am_tsv_shape(L.np_ptr(buf), buf.size, L.np_ptr(n_rows), L.np_ptr(n_fields)))
    n, f = int(n_rows[0]), int(n_fields[0])
    if n == 0 or f != len(fields):
        return None
    idb = np.empty(n, np.int64)
    ide = np.empty(n, np.int64)
    values = np.empty((n, f - 1), np.float64)
    all_int = np.zeros(f - 1, np.uint8)
    rc = L.lib.am_tsv_parse_f64(L.np_ptr(buf), buf.size, n, f, L.np_ptr(idb), L.np_ptr(ide), L.np_ptr(values),
                                L.np_ptr(all_int), _threads())
    if rc != 0:
        return None
    mv = memoryview(buf)  # the labels are sliced out of the file image in place (no second copy of the file)
    try:
        ids = [str(mv[b:e], "utf-8") for b, e in zip(idb.tolist(), ide.tolist())]
    except UnicodeDecodeError:
        return None
    # pandas would convert an all-numeric / NA-spelled index column; leave those tables to it
    if _looks_numeric(ids[0]) or not _NA_LABELS.isdisjoint(ids) or any(i != i.strip(" ") for i in ids[:1]):
        return None
    index = pd.Index(ids, name=index_col)
    cols = fields[1:]
    if not all_int.any():
        return pd.DataFrame(values, index=index, columns=cols, copy=False)  # `values` is ours: no defensive copy
    data = {c: (values[:, j].astype(np.int64) if all_int[j] else values[:, j]) for j, c in enumerate(cols)}
    return pd.DataFrame(data, index=index, columns=cols)
