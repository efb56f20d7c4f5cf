"""Host-side containers that need neither the shared library nor CUDA.

Kept apart from ``assembly.py`` (which binds ``libautometa_b200.so`` at import) so that the
synthetic-data generator and the CPU reference arm of ``bench.py`` can be imported
without loading the extension.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np


@dataclass
class HostAssembly:
    """Contig ids + ASCII sequence bytes (concatenated) + offsets, on the host."""

    ids: List[str]
    seq: np.ndarray  # uint8, all contigs concatenated (may carry slack at the end)
    offsets: np.ndarray  # int64 [N+1]
    genome: Optional[np.ndarray] = None  # synthetic assemblies only: source genome of every contig

    @property
    def n_contigs(self) -> int:
        return len(self.offsets) - 1

    @property
    def lengths(self) -> np.ndarray:
        return np.diff(self.offsets)

    @property
    def total_bp(self) -> int:
        return int(self.offsets[-1] - self.offsets[0])

    def slice(self, lo: int, hi: int) -> "HostAssembly":
        """Contigs [lo, hi) as a view (offsets stay absolute into ``seq``)."""
        return HostAssembly(self.ids[lo:hi], self.seq, self.offsets[lo : hi + 1],
                            None if self.genome is None else self.genome[lo:hi])
