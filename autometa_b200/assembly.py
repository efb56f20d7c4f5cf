"""FASTA ingest and 2-bit packing (host side of kernel 1).

Replaces the Biopython ``SeqIO.parse`` + ``multiprocessing`` fan-out of
``mp_counter`` / ``seq_counter`` (autometa/common/kmers.py:122-158, 209-265):
the whole assembly is indexed once in C (``am_fasta_index``), packed to 2 bits per
base (``am_pack_host`` on host threads, or ``am_pack_device`` on the GPU) and
handed to the count kernel as flat device arrays.
"""
from __future__ import annotations

import gzip
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _lib as L
from .hostdata import HostAssembly  # noqa: F401  (re-exported; lives in a module that needs no .so)


@dataclass
class PackedAssembly:
    """Device-resident 2-bit-packed contigs (layout: include/autometa_b200.h)."""

    n_contigs: int
    total_bp: int
    total_packed_bases: int
    lengths: np.ndarray  # int32 [N] (host)
    starts_host: np.ndarray  # int64 [N] (host)
    codes: "torch.Tensor"  # int32 words (bit pattern of uint32)
    exc_bitmap: "torch.Tensor"
    exc_rank: "torch.Tensor"
    exc_mask: "torch.Tensor"  # int64 (bit pattern of uint64)
    starts: "torch.Tensor"  # int64 [N] device
    n_exc: int = 0
    ids: Optional[List[str]] = None
    gc_content: Optional["torch.Tensor"] = None  # fp64 [N] percent, when packed with gc=True
    _plans: dict = field(default_factory=dict)

    @property
    def device(self):
        return self.codes.device


def read_raw(path: str) -> bytes:
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rb") as fh:
        return fh.read()


def read_fasta(path: str) -> HostAssembly:
    """Parse a (optionally gzipped) FASTA file; ids = first token of each header,
    sequences concatenated with case preserved (kmers.py:143-148 semantics)."""
    return parse_fasta_bytes(read_raw(path))


def record_titles(raw: bytes) -> List[bytes]:
    """Header lines without '>' and trailing whitespace, CR/LF inside replaced by blanks:
    what Bio.SeqIO's FastaWriter prints for a record that came from its FASTA parser."""
    buf = np.frombuffer(raw, dtype=np.uint8)
    gt = np.flatnonzero(buf == ord(">"))
    gt = gt[(gt == 0) | (buf[np.maximum(gt, 1) - 1] == ord("\n"))]
    titles = []
    for p in gt:
        e = raw.find(b"\n", int(p))
        titles.append(raw[int(p) + 1 : e if e >= 0 else len(raw)].rstrip().replace(b"\r", b" "))
    return titles


def parse_fasta_bytes(raw: bytes) -> HostAssembly:
    buf = np.frombuffer(raw, dtype=np.uint8)
    n = len(buf)
    seq = np.empty(n + 64, dtype=np.uint8)
    nrec = np.zeros(1, dtype=np.int64)
    if n:  # the number of records from the library's own (threaded) header count, not from a scan in Python
        L.check(L.lib.am_fasta_index_mt(L.np_ptr(buf), n, None, None, None, None, 0, L.np_ptr(nrec), 0))
    cap = max(16, int(nrec[0]))
    offsets = np.zeros(cap + 1, dtype=np.int64)
    idb = np.zeros(cap, dtype=np.int64)
    ide = np.zeros(cap, dtype=np.int64)
    L.check(
        L.lib.am_fasta_index(
            L.np_ptr(buf) if n else None, n, L.np_ptr(seq), L.np_ptr(offsets), L.np_ptr(idb),
            L.np_ptr(ide), cap, L.np_ptr(nrec),
        )
    )
    k = int(nrec[0])
    ids = [raw[int(b) : int(e)].decode("ascii", errors="replace") for b, e in zip(idb[:k], ide[:k])]
    return HostAssembly(ids, seq, offsets[: k + 1].copy())


def from_sequences(ids: Sequence[str], seqs: Sequence[bytes]) -> HostAssembly:
    lens = np.fromiter((len(s) for s in seqs), dtype=np.int64, count=len(seqs))
    offsets = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    seq = np.empty(int(offsets[-1]) + 64, dtype=np.uint8)
    seq[: int(offsets[-1])] = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    return HostAssembly(list(ids), seq, offsets)


def pack_plan(offsets: np.ndarray):
    n = len(offsets) - 1
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    starts = np.zeros(max(n, 1), dtype=np.int64)
    lengths = np.zeros(max(n, 1), dtype=np.int32)
    total = np.zeros(1, dtype=np.int64)
    L.check(L.lib.am_pack_plan(L.np_ptr(offsets), n, L.np_ptr(starts), L.np_ptr(lengths), L.np_ptr(total)))
    return starts[:n], lengths[:n], int(total[0])


def pack_host_arrays(asm: HostAssembly, threads: int = 0):
    """Pack on the host. Returns (starts, lengths, total, codes, bitmap, rank, masks) numpy arrays."""
    offsets = np.ascontiguousarray(asm.offsets, dtype=np.int64)
    starts, lengths, total = pack_plan(offsets)
    n = asm.n_contigs
    codes = np.empty(L.lib.am_pack_code_words(total), dtype=np.uint32)
    nbw = L.lib.am_pack_bitmap_words(total)
    bitmap = np.empty(nbw, dtype=np.uint32)
    rank = np.empty(nbw, dtype=np.uint32)
    cap = max(1024, total // 64 // 8)
    nexc = np.zeros(1, dtype=np.int64)
    if threads <= 0:
        threads = min(32, os.cpu_count() or 1)
    while True:
        masks = np.empty(cap, dtype=np.uint64)
        rc = L.lib.am_pack_host(
            L.np_ptr(asm.seq), L.np_ptr(offsets), n, L.np_ptr(starts) if n else None, total,
            L.np_ptr(codes), L.np_ptr(bitmap), L.np_ptr(rank), L.np_ptr(masks), cap,
            L.np_ptr(nexc), threads,
        )
        if rc == L.AM_ECAPACITY:
            cap = int(nexc[0])
            continue
        L.check(rc)
        break
    return starts, lengths, total, codes, bitmap, rank, masks[: int(nexc[0])]


def _as_i32(a: np.ndarray):
    return a.view(np.int32)


def pack_to_device(asm: HostAssembly, device=None, threads: int = 0, on_device: bool = False,
                   gc: bool = False) -> PackedAssembly:
    """Pack ``asm`` and place it in HBM.

    ``on_device=False``: 2-bit packing on host threads, then one H2D copy of the
    packed arrays (B/4 bytes).  ``on_device=True``: H2D of the ASCII bytes (B
    bytes) and packing by the ``am_pack_device`` kernel.
    """
    import torch

    device = L.require_cuda(device)
    n = asm.n_contigs
    if gc:
        on_device = True  # GC% is a by-product of the device packer's pass over the ASCII bytes
    if not on_device:
        starts, lengths, total, codes, bitmap, rank, masks = pack_host_arrays(asm, threads)
        with torch.cuda.device(device):
            t = lambda a, dt: torch.from_numpy(a.view(dt)).to(device, non_blocking=False)
            return PackedAssembly(
                n_contigs=n, total_bp=asm.total_bp, total_packed_bases=total, lengths=lengths,
                starts_host=starts, codes=t(codes, np.int32), exc_bitmap=t(bitmap, np.int32),
                exc_rank=t(rank, np.int32),
                exc_mask=t(masks if len(masks) else np.zeros(1, np.uint64), np.int64),
                starts=torch.from_numpy(starts if n else np.zeros(1, np.int64)).to(device),
                n_exc=len(masks), ids=asm.ids,
            )
    # device packing
    offsets = np.ascontiguousarray(asm.offsets, dtype=np.int64)
    base = int(offsets[0])
    starts, lengths, total = pack_plan(offsets)
    with torch.cuda.device(device):
        nbytes = asm.total_bp
        seq_h = torch.from_numpy(asm.seq[base : base + nbytes])
        d_seq = torch.empty(nbytes + 64, dtype=torch.uint8, device=device)
        d_seq[:nbytes].copy_(seq_h, non_blocking=True)
        d_off = torch.from_numpy(offsets - base).to(device)
        d_starts = torch.from_numpy(starts if n else np.zeros(1, np.int64)).to(device)
        return pack_device_arrays(d_seq, d_off, d_starts, n, total, lengths, starts, asm.total_bp, asm.ids, gc=gc)


def pack_device_arrays(d_seq, d_off, d_starts, n, total, lengths, starts_host, total_bp, ids=None,
                       exc_capacity: int = 0, gc: bool = False) -> PackedAssembly:
    """Run ``am_pack_device`` on ASCII bytes that already live in HBM; ``gc=True`` also
    returns ``Metagenome.gc_content``'s percentage per contig (metagenome.py:303-320)."""
    import torch

    device = d_seq.device
    nwords = L.lib.am_pack_code_words(total)
    nbw = L.lib.am_pack_bitmap_words(total)
    with torch.cuda.device(device):
        codes = torch.empty(nwords, dtype=torch.int32, device=device)
        bitmap = torch.empty(nbw, dtype=torch.int32, device=device)
        rank = torch.empty(nbw, dtype=torch.int32, device=device)
        work = torch.empty(L.lib.am_pack_work_words(total), dtype=torch.int32, device=device)
        n_exc = torch.zeros(1, dtype=torch.int64, device=device)
        gc_at = torch.empty((max(n, 1), 2), dtype=torch.int32, device=device) if gc else None
        cap = exc_capacity if exc_capacity > 0 else max(1024, total // 64 // 16)
        while True:
            masks = torch.empty(cap, dtype=torch.int64, device=device)
            L.check(
                L.lib.am_pack_device_gc(
                    L.t_ptr(d_seq), L.t_ptr(d_off), L.t_ptr(d_starts), n, total, L.t_ptr(codes),
                    L.t_ptr(bitmap), L.t_ptr(rank), L.t_ptr(masks), cap, L.t_ptr(n_exc),
                    L.t_ptr(gc_at) if gc else None, L.t_ptr(work), L.stream_ptr(device),
                )
            )
            found = int(n_exc.item())
            if found <= cap:
                break
            cap = found
        gc_content = None
        if gc:
            gc_content = torch.empty(max(n, 1), dtype=torch.float64, device=device)
            L.check(L.lib.am_gc_finish(L.t_ptr(gc_at), n, L.t_ptr(gc_content), L.stream_ptr(device)))
            gc_content = gc_content[:n]
        return PackedAssembly(
            n_contigs=n, total_bp=total_bp, total_packed_bases=total, lengths=lengths,
            starts_host=starts_host, codes=codes, exc_bitmap=bitmap, exc_rank=rank, exc_mask=masks,
            starts=d_starts, n_exc=found, ids=ids, gc_content=gc_content,
        )


def count_plan(lengths: np.ndarray, k: int, max_item_bases: int = 32768):
    """Host work list for the count kernel (``am_count_plan``)."""
    n = len(lengths)
    lengths = np.ascontiguousarray(lengths, dtype=np.int32)
    nwin = np.maximum(lengths.astype(np.int64) - k, 0)
    cap = int(np.sum((nwin + max_item_bases - 1) // max_item_bases) + np.sum(nwin == 0)) + 1
    items = np.empty((cap, 4), dtype=np.int32)
    multi = np.empty(max(1, int(np.sum(nwin > max_item_bases))), dtype=np.int32)
    ni = np.zeros(1, dtype=np.int64)
    nm = np.zeros(1, dtype=np.int64)
    L.check(
        L.lib.am_count_plan(
            L.np_ptr(lengths) if n else None, n, k, max_item_bases, L.np_ptr(items), cap,
            L.np_ptr(ni), L.np_ptr(multi), len(multi), L.np_ptr(nm),
        )
    )
    return items[: int(ni[0])], multi[: int(nm[0])]
