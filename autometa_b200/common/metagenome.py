"""Drop-in for the sequence-statistics part of ``autometa.common.metagenome.Metagenome``
(FASTA ingest with length and GC% as by-products of the packing pass).

=======================  ================================  ===========================
member                   reference                         where it is computed
=======================  ================================  ===========================
``gc_content()``         metagenome.py:303-320             am_pack_device_gc + am_gc_finish
``length_weighted_gc``   metagenome.py:124-144             device GC%, ``np.average`` on host
``nseqs`` / ``size``     metagenome.py:112-122,146-156     FASTA index (``am_fasta_index``)
``largest_seq``          metagenome.py:158-175             FASTA index
``fragmentation_metric`` metagenome.py:177-208             lengths from the FASTA index
``describe()``           metagenome.py:210-243             the above
``length_filter()``      metagenome.py:245-300             host (file in, file out)
=======================  ================================  ===========================

ORF calling (prodigal) and the CLI are out of scope.  There is no CPU fallback for
the GC pass: without a CUDA device ``gc_content`` raises.
"""
from __future__ import annotations

import logging
import numbers
import os
from typing import List, Optional

import numpy as np
import pandas as pd

from .. import assembly as _asm

logger = logging.getLogger(__name__)

_WRAP = 60  # Bio.SeqIO's FastaWriter default line width


class Metagenome:
    """Sequence statistics of a metagenome assembly (``Metagenome(assembly=path)``)."""

    def __init__(self, assembly, device=None):
        self.assembly = os.path.realpath(assembly)
        self._device = device
        self._host: Optional[_asm.HostAssembly] = None
        self._raw: Optional[bytes] = None
        self._packed: Optional[_asm.PackedAssembly] = None
        self._gc: Optional[np.ndarray] = None

    def __repr__(self) -> str:
        return str(self)

    def __str__(self) -> str:
        return self.assembly

    # ------------------------------------------------------------------ ingest
    def _load(self) -> _asm.HostAssembly:
        if self._host is None:
            self._raw = _asm.read_raw(self.assembly)
            self._host = _asm.parse_fasta_bytes(self._raw)
        return self._host

    @property
    def packed(self) -> _asm.PackedAssembly:
        """The 2-bit packed assembly in HBM (what ``kmers`` counting consumes); the same
        pass produced the GC percentages."""
        if self._packed is None:
            self._packed = _asm.pack_to_device(self._load(), device=self._device, gc=True)
            self._gc = self._packed.gc_content.cpu().numpy()
        return self._packed

    @property
    def lengths(self) -> np.ndarray:
        return np.asarray(self._load().lengths, dtype=np.int64)

    @property
    def ids(self) -> List[str]:
        return self._load().ids

    # --------------------------------------------------------- reference API
    @property
    def sequences(self) -> List[str]:
        asm = self._load()
        buf, off = asm.seq, asm.offsets
        return [buf[int(off[i]) : int(off[i + 1])].tobytes().decode("ascii", errors="replace")
                for i in range(asm.n_contigs)]

    @property
    def nseqs(self) -> int:
        return self._load().n_contigs

    @property
    def size(self) -> int:
        return int(self.lengths.sum())

    @property
    def length_weighted_gc(self) -> float:
        self.packed
        lengths = self.lengths
        size = int(lengths.sum())
        weights = [int(n) / size for n in lengths]
        return float(np.average(a=[float(g) for g in self._gc], weights=weights))

    @property
    def largest_seq(self) -> str:
        lengths = self.lengths
        if not len(lengths):
            raise AttributeError("'NoneType' object has no attribute 'id'")  # what the reference ends in
        return self.ids[int(np.argmax(lengths))]  # first of the longest, like the strict '>' scan

    def fragmentation_metric(self, quality_measure: float = 0.50) -> int:
        if not 0 < quality_measure < 1:
            raise ValueError(f"quality measure must be b/w 0 and 1! given: {quality_measure}")
        lengths = np.sort(self.lengths)[::-1]
        target_size = int(lengths.sum()) * quality_measure
        reached = np.flatnonzero(np.cumsum(lengths) >= target_size)
        return int(lengths[reached[0]]) if len(reached) else None

    def describe(self) -> pd.DataFrame:
        return pd.DataFrame(
            [
                {
                    "assembly": self.assembly,
                    "nseqs": self.nseqs,
                    "size (bp)": self.size,
                    "N50 (bp)": self.fragmentation_metric(0.5),
                    "N10 (bp)": self.fragmentation_metric(0.1),
                    "N90 (bp)": self.fragmentation_metric(0.9),
                    "length_weighted_gc_content (%)": self.length_weighted_gc,
                    "largest_seq": self.largest_seq,
                }
            ]
        ).set_index("assembly")

    def gc_content(self) -> pd.DataFrame:
        """index="contig", columns=["gc_content", "length"] in FASTA order."""
        self.packed
        if not self.nseqs:
            return pd.DataFrame([]).set_index("contig")  # the reference's KeyError on an empty assembly
        return pd.DataFrame(
            {"contig": self.ids, "gc_content": self._gc.astype(np.float64), "length": self.lengths}
        ).set_index("contig")

    def length_filter(self, out: str, cutoff: int = 3000, force: bool = False) -> "Metagenome":
        if not isinstance(cutoff, numbers.Number) or isinstance(cutoff, bool):
            raise TypeError(f"cutoff: {cutoff} must be a float or int")
        if cutoff <= 0:
            raise ValueError(f"cutoff: {cutoff} must be a positive real number")
        if os.path.exists(out) and not force:
            raise FileExistsError(out)
        logger.info(f"Getting contigs greater than or equal to {cutoff:,} bp")
        asm = self._load()
        keep = np.flatnonzero(self.lengths >= cutoff)
        outdir = os.path.dirname(os.path.abspath(out))
        if not os.path.exists(outdir):
            os.makedirs(outdir)
        titles = _asm.record_titles(self._raw)
        buf, off = asm.seq, asm.offsets
        chunks = []
        for i in keep:
            seq = buf[int(off[i]) : int(off[i + 1])].tobytes()
            chunks.append(b">" + titles[i] + b"\n")
            chunks.extend(seq[p : p + _WRAP] + b"\n" for p in range(0, len(seq), _WRAP))
        with open(out, "wb") as fh:
            fh.write(b"".join(chunks))
        if not len(keep):
            logger.warning(f"No contigs pass {cutoff:,}. Returning original Metagenome.")
            return self
        return Metagenome(assembly=out, device=self._device)
