"""TEST INFRASTRUCTURE: seeded inputs of ``large_data_mode.cluster_by_taxon_partitioning``, shared by
tests/golden/make_golden.py (which runs the reference on them) and the GPU parity test.  Uses the oracle's
C count port and GC restatement to build the tables, so it lives under tests/ and not in the package."""
import numpy as np

from autometa_b200.synth import make_assembly, sequences


def make_ldm_input(n_contigs: int = 1700, k: int = 4, seed: int = 20261019 + 6, n_markers: int = 40):
    """Inputs of ``large_data_mode.cluster_by_taxon_partitioning`` from a seeded synthetic assembly:
    ``(main, counts, markers)`` — ``main``: coverage, gc_content and the seven canonical-rank columns
    (taxa = groups of source genomes, one taxon per genome from class down; ~4 % of the contigs are
    unclassified everywhere); ``counts``: the float/NaN k-mer frame ``kmers.load`` returns; ``markers``: wide
    marker table (each genome carries each marker with p = 0.9 on one of its contigs).
    Partition sizes are chosen so that the reference's ``PCA(svd_solver="auto")`` takes an EXACT solver
    everywhere at k = 4 (D = 136): <= 500 rows -> "full", >= 1360 rows -> "covariance_eigh"; never the
    randomized branch in between."""
    import pandas as pd

    from oracle import oracle_c, oracle_np

    rng = np.random.default_rng(seed)
    asm = make_assembly(n_contigs, n_contigs * 3300, seed=seed, noise=True)
    names, _ = oracle_np.kmer_columns(k)
    cnt = oracle_c.count(asm.seq, asm.offsets, k, 1).astype(np.float64)
    cnt[cnt == 0] = np.nan
    index = pd.Index(asm.ids, name="contig")
    counts = pd.DataFrame(cnt, index=index, columns=names)
    genome = asm.genome
    G = int(genome.max()) + 1
    sizes = np.bincount(genome, minlength=G)
    # superkingdom: trailing genomes -> Archaea until it holds 120 .. 300 contigs; the rest Bacteria
    arch, tot = [], 0
    for g in range(G - 1, -1, -1):
        if tot >= 120:
            break
        arch.append(g)
        tot += sizes[g]
    arch = set(arch)
    # phylum: consecutive genomes grouped up to 430 contigs; class: one per genome
    phylum, cur, acc = {}, 0, 0
    for g in range(G):
        if acc + sizes[g] > 430 and acc > 0:
            cur, acc = cur + 1, 0
        phylum[g] = f"p_{cur}"
        acc += sizes[g]
    cov_g = np.exp(rng.uniform(np.log(3.0), np.log(300.0), size=G))
    coverage = cov_g[genome] * np.exp(rng.normal(0, 0.04, size=n_contigs))
    gc = np.array([oracle_np.gc_fraction_remove(s) * 100.0 for s in sequences(asm)])
    main = pd.DataFrame({"coverage": coverage, "gc_content": gc}, index=index)
    main["superkingdom"] = np.where(np.isin(genome, list(arch)), "Archaea", "Bacteria")
    main["phylum"] = [phylum[g] for g in genome]
    main["class"] = [f"c_{g}" for g in genome]
    # the reference cannot step over two consecutive ranks without classified contigs (its fallback embeds the
    # previous rank's — empty — table, kmers.py:576-580), so every rank is populated: one taxon per genome below class
    for rank in ("order", "family", "genus", "species"):
        main[rank] = [f"{rank[0]}_{g}" for g in genome]
    drop = rng.random(n_contigs) < 0.04
    for rank in ("superkingdom", "phylum", "class", "order", "family", "genus", "species"):
        col = main[rank].to_numpy(dtype=object)
        col[drop] = "unclassified"
        main[rank] = col
    mk = np.zeros((n_contigs, n_markers), dtype=np.int64)
    for g in range(G):
        mem = np.nonzero(genome == g)[0]
        if len(mem) == 0:
            continue
        have = np.nonzero(rng.random(n_markers) < 0.9)[0]
        who = mem[rng.integers(0, len(mem), size=len(have))]
        np.add.at(mk, (who, have), 1)
    keep = mk.sum(axis=1) > 0
    markers = pd.DataFrame(mk[keep], index=pd.Index(np.asarray(asm.ids, dtype=object)[keep], name="contig"),
                           columns=[f"PF{m:05d}" for m in range(n_markers)])
    return main, counts, markers
