"""Seeded synthetic assemblies and sweep inputs shared by the tests and bench.py.

Sizes follow BASELINE.json's configs (C1 10k contigs / 50 Mbp, C2 200k / 1 Gbp,
C5 1M / 5 Gbp; C4 = 200k points for the eps sweep).  The generator is vectorised
numpy so that 1 Gbp builds in tens of seconds on the host:

* G = max(20, N/500) genomes; genome g has its own hexamer distribution
  (Dirichlet around a genome-specific base composition), so k-mer spectra differ
  between genomes beyond GC content; a contig is a slice of its genome's stream of
  i.i.d. hexamers.  (SURVEY.md §8d sketches a first-order Markov chain; i.i.d.
  hexamer blocks give the same kind of signal and vectorise.)
* contig lengths 3000 + LogNormal(sigma=0.8), rescaled so that they sum to B
  exactly (3000 = Autometa's default length cutoff, workflows/autometa.sh:27).
* noise exercising the reference's skip rules (kmers.py:197-204): ~0.05 % of bases
  become 'N' in runs of 1-50, 0.5 % of contigs get one lower-case run of 10-200;
  ``tiny=True`` appends contigs of length 4, 5, 6 (the L <= k edge, kmers.py:187-192).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from .hostdata import HostAssembly

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def contig_lengths(n: int, total_bp: int, rng: np.random.Generator, min_len: int = 3000) -> np.ndarray:
    if total_bp < n * min_len:
        raise ValueError("total_bp too small for n contigs of the minimum length")
    extra = rng.lognormal(mean=7.0, sigma=0.8, size=n)
    budget = total_bp - n * min_len
    extra = extra / extra.sum() * budget
    lens = min_len + np.floor(extra).astype(np.int64)
    short = total_bp - int(lens.sum())
    if short > 0:
        idx = np.arange(short) % n
        np.add.at(lens, idx, 1)
    return lens


def make_assembly(n_contigs: int, total_bp: int, seed: int, noise: bool = True, tiny: bool = False) -> HostAssembly:
    rng = np.random.default_rng(seed)
    lens = contig_lengths(n_contigs, total_bp, rng)
    G = max(20, n_contigs // 500)
    weights = rng.dirichlet(np.full(G, 2.0))
    genome = np.sort(rng.choice(G, size=n_contigs, p=weights))
    seq = np.empty(total_bp + (15 if tiny else 0) + 64, dtype=np.uint8)
    offsets = np.zeros(n_contigs + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    # hexamer table: 4096 x 6 letters
    hexa = np.zeros((4096, 6), dtype=np.uint8)
    codes = np.arange(4096)
    for j in range(6):
        hexa[:, j] = _ACGT[(codes >> (2 * (5 - j))) & 3]
    for g in range(G):
        members = np.nonzero(genome == g)[0]
        if len(members) == 0:
            continue
        lo, hi = int(offsets[members[0]]), int(offsets[members[-1] + 1])
        nb = hi - lo
        base = rng.dirichlet(np.full(4, 8.0))
        p = np.ones(4096)
        for j in range(6):
            p *= base[(codes >> (2 * j)) & 3]
        p = rng.dirichlet(p / p.sum() * 4096 * 4.0 + 1e-3)
        cdf = np.cumsum(p)
        cdf[-1] = 1.0
        nhex = (nb + 5) // 6
        draws = np.searchsorted(cdf, rng.random(nhex), side="right").astype(np.int64)
        np.minimum(draws, 4095, out=draws)
        stream = hexa[draws].reshape(-1)
        seq[lo:hi] = stream[:nb]
    if noise:
        n_runs = max(1, int(total_bp * 0.0005 / 25.5))
        starts = rng.integers(0, total_bp, size=n_runs)
        run = rng.integers(1, 51, size=n_runs)
        idx = np.repeat(starts, run) + (np.arange(int(run.sum())) - np.repeat(np.cumsum(run) - run, run))
        idx = idx[idx < total_bp]
        seq[idx] = ord("N")
        n_lower = max(1, int(n_contigs * 0.005))
        for c in rng.choice(n_contigs, size=n_lower, replace=False):
            L = int(lens[c])
            ln = int(rng.integers(10, 201))
            s = int(rng.integers(0, max(1, L - ln)))
            a = int(offsets[c]) + s
            seq[a : a + ln] |= 0x20  # lower-case (also turns 'N' into 'n')
    cov = np.exp(rng.uniform(np.log(2.0), np.log(500.0), size=n_contigs))
    ids = [f"NODE_{i + 1}_length_{int(l)}_cov_{c:.3f}" for i, (l, c) in enumerate(zip(lens, cov))]
    if tiny:
        pos = total_bp
        for k, L in enumerate((4, 5, 6)):
            seq[pos : pos + L] = _ACGT[rng.integers(0, 4, size=L)]
            pos += L
            offsets = np.append(offsets, pos)
            ids.append(f"NODE_tiny{k}_length_{L}_cov_1.000")
        genome = np.concatenate([genome, np.full(3, -1, dtype=genome.dtype)])
    return HostAssembly(ids, seq, offsets, genome)


def write_fasta(asm: HostAssembly, path: str, width: int = 60) -> None:
    with open(path, "wb") as fh:
        for i, cid in enumerate(asm.ids):
            fh.write(b">" + cid.encode() + b"\n")
            s = asm.seq[int(asm.offsets[i]) : int(asm.offsets[i + 1])].tobytes()
            for a in range(0, len(s), width):
                fh.write(s[a : a + width] + b"\n")
            if not s:
                fh.write(b"\n")


def sequences(asm: HostAssembly) -> List[bytes]:
    return [asm.seq[int(asm.offsets[i]) : int(asm.offsets[i + 1])].tobytes() for i in range(asm.n_contigs)]


def make_sweep_input(n_points: int = 200_000, n_clusters: int = 400, n_markers: int = 120,
                     seed: int = 20261019 + 4):
    """(table DataFrame [x_1, x_2, coverage, gc_content], markers DataFrame) per SURVEY.md §8d (C4)."""
    import pandas as pd

    rng = np.random.default_rng(seed)
    K = n_clusters
    centres = rng.uniform(-100, 100, size=(K, 2))
    w = rng.dirichlet(np.full(K, 2.0))
    member = rng.choice(K, size=n_points, p=w)
    xy = centres[member] + rng.normal(0, 2.0, size=(n_points, 2))
    ck = np.exp(rng.uniform(np.log(2.0), np.log(500.0), size=K))
    coverage = ck[member] * np.exp(rng.normal(0, 0.05, size=n_points))
    gk = rng.uniform(25, 75, size=K)
    gc = gk[member] + rng.normal(0, 1.0, size=n_points)
    ids = [f"NODE_{i + 1}" for i in range(n_points)]
    table = pd.DataFrame(
        {"x_1": xy[:, 0], "x_2": xy[:, 1], "coverage": coverage, "gc_content": gc},
        index=pd.Index(ids, name="contig"),
    )
    markers = np.zeros((n_points, n_markers), dtype=np.int64)
    order = np.argsort(member, kind="stable")
    bounds = np.searchsorted(member[order], np.arange(K + 1))
    for k in range(K):
        mem = order[bounds[k] : bounds[k + 1]]
        if len(mem) == 0:
            continue
        have = np.nonzero(rng.random(n_markers) < 0.9)[0]
        who = mem[rng.integers(0, len(mem), size=len(have))]
        np.add.at(markers, (who, have), 1)
    keep = markers.sum(axis=1) > 0
    markers_df = pd.DataFrame(
        markers[keep], index=pd.Index(np.asarray(ids, dtype=object)[keep], name="contig"),
        columns=[f"PF{m:05d}" for m in range(n_markers)],
    )
    return table, markers_df


def add_taxonomy(table, seed: int = 3):
    """Deterministic canonical-rank columns for ``taxon_guided_binning`` inputs: taxa are bands of the
    embedding plane (so that contigs of a taxon are spatially coherent), a few contigs "unclassified"."""
    rng = np.random.default_rng(seed)
    x, y = table["x_1"].to_numpy(), table["x_2"].to_numpy()
    out = table.copy()
    out["superkingdom"] = np.where(x < 0, "Bacteria", "Archaea")
    out["phylum"] = ["p_%d" % v for v in (np.floor(x / 50).astype(int) * 10 + np.floor(y / 100).astype(int))]
    out["class"] = ["c_%d" % v for v in (np.floor(x / 25).astype(int) * 100 + np.floor(y / 50).astype(int))]
    for rank in ("order", "family", "genus", "species"):
        out[rank] = "unclassified"
    drop = rng.random(len(out)) < 0.05
    for rank in ("superkingdom", "phylum", "class"):
        col = out[rank].to_numpy(dtype=object)
        col[drop] = "unclassified"
        out[rank] = col
    return out
