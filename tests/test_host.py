"""CPU-side tests: the C-ABI library loads and exports every declared symbol, host
logic (FASTA index, packers, work list, metrics, sharding) matches the oracle and
the reference's golden outputs.  No CUDA compute calls here."""
import os
import re

import numpy as np
import pandas as pd
import pytest

from oracle import oracle_np as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    from autometa_b200 import _lib as L

    hdr = open(os.path.join(ROOT, "include", "autometa_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(am_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    for name in sorted(declared):
        assert hasattr(L.lib, name), f"{name} declared in the header but not exported"
    assert declared == set(L.EXPORTS)
    assert L.lib.am_version() >= 100


def test_column_table_matches_oracle():
    from autometa_b200 import device as D

    for k in range(1, 8):
        names, lut = O.kmer_columns(k)
        assert D.num_columns(k) == len(names)
        assert D.column_names(k) == names
        assert np.array_equal(D.column_lut(k), lut)


def test_init_kmers_api(golden_kmers):
    from autometa_b200.common import kmers

    d = kmers.init_kmers(5)
    assert list(d) == golden_kmers["count_columns"].tolist()
    assert list(d.values()) == list(range(512))
    with pytest.raises(TypeError):
        kmers.init_kmers(5.5)
    assert kmers._revcomp("AACGT") == "ACGTT"


def test_fasta_index_matches_oracle(small_fna, tmp_path):
    from autometa_b200 import assembly as A

    ids, seqs = O.read_fasta(small_fna)
    h = A.read_fasta(small_fna)
    assert h.ids == ids
    assert [bytes(h.seq[h.offsets[i] : h.offsets[i + 1]]) for i in range(h.n_contigs)] == seqs
    # gz, CRLF, blank lines, header-only record, no trailing newline
    import gzip

    raw = b">a x y\r\nAC GT\r\n\r\nNN\r\n>b\n>c\tq\nacgt"
    p = tmp_path / "t.fa.gz"
    with gzip.open(p, "wb") as fh:
        fh.write(raw)
    h = A.read_fasta(str(p))
    assert h.ids == ["a", "b", "c"]
    assert [bytes(h.seq[h.offsets[i] : h.offsets[i + 1]]) for i in range(3)] == [b"ACGTNN", b"", b"acgt"]
    assert O.read_fasta(str(p))[1] == [b"ACGTNN", b"", b"acgt"]
    with pytest.raises(FileNotFoundError):
        A.read_fasta(str(tmp_path / "missing.fa"))


def _py_pack(seqs):
    enc = {65: 0, 84: 1, 67: 2, 71: 3}
    starts, pos = [], 0
    for s in seqs:
        starts.append(pos)
        pos += (len(s) + 63) // 64 * 64
    total = pos + 128
    codes = np.zeros(total // 16, dtype=np.uint32)
    masks = {}
    for st, s in zip(starts, seqs):
        for j, ch in enumerate(s):
            b = st + j
            if ch in enc:
                codes[b >> 4] |= np.uint32(enc[ch] << (2 * (b & 15)))
            else:
                masks[b >> 6] = masks.get(b >> 6, 0) | (1 << (b & 63))
    return np.array(starts), total, codes, masks


def test_pack_host_matches_python_restatement(small_fna):
    from autometa_b200 import assembly as A

    h = A.read_fasta(small_fna)
    _, seqs = O.read_fasta(small_fna)
    for threads in (1, 3):
        starts, lengths, total, codes, bitmap, rank, masks = A.pack_host_arrays(h, threads)
        pstarts, ptotal, pcodes, pmasks = _py_pack(seqs)
        assert np.array_equal(starts, pstarts) and total == ptotal
        assert np.array_equal(lengths, [len(s) for s in seqs])
        assert np.array_equal(codes, pcodes)
        groups = sorted(pmasks)
        flagged = [g for g in range(total // 64) if (bitmap[g >> 5] >> (g & 31)) & 1]
        assert flagged == groups
        assert [int(m) for m in masks] == [pmasks[g] for g in groups]
        for g in groups:
            r = int(rank[g >> 5]) + bin(int(bitmap[g >> 5]) & ((1 << (g & 31)) - 1)).count("1")
            assert int(masks[r]) == pmasks[g]


def test_count_plan():
    from autometa_b200 import assembly as A

    lengths = np.array([3, 5, 6, 2048 + 5, 40000, 100000, 7], dtype=np.int32)
    items, multi = A.count_plan(lengths, 5, 32768)
    assert multi.tolist() == [4, 5]
    per = {}
    for c, w0, nw, fl in items.tolist():
        assert w0 % 2048 == 0 and 0 <= nw <= 32768
        assert fl == (1 if c in (4, 5) else 0)
        per.setdefault(c, []).append((w0, nw))
    for c, L in enumerate(lengths.tolist()):
        spans = sorted(per[c])
        assert sum(n for _, n in spans) == max(L - 5, 0)
        pos = 0
        for w0, n in spans:
            assert w0 == pos
            pos += n


def _sweep_table():
    from autometa_b200 import synth

    return synth.make_sweep_input(n_points=3000, n_clusters=12, n_markers=40, seed=11)


def test_add_metrics_matches_reference_golden(golden_dbscan):
    from autometa_b200.binning import utilities as U

    table, markers = _sweep_table()
    df = table.copy()
    df["cluster"] = golden_dbscan["labels_2"]  # eps = 1.2
    cdf, mdf = U.add_metrics(df, markers)
    assert mdf.index.name == "cluster"
    assert np.array_equal(mdf.index.to_numpy(), golden_dbscan["metrics_cluster_index"])
    assert list(mdf.columns) == U.METRICS
    for name in U.METRICS:
        np.testing.assert_allclose(mdf[name].to_numpy(dtype=float), golden_dbscan[f"metrics_{name}"],
                                   rtol=1e-12, atol=1e-12, equal_nan=True)
        np.testing.assert_allclose(cdf[name].to_numpy(dtype=float), golden_dbscan[f"contig_{name}"],
                                   rtol=1e-12, atol=1e-12, equal_nan=True)
    assert list(cdf.columns) == ["x_1", "x_2", "coverage", "gc_content", "cluster"] + U.METRICS
    f = U.apply_binning_metrics_filter(mdf, 20.0, 90.0, 25.0, 5.0)
    ok = U.passing_mask({k: mdf[k].to_numpy(dtype=float) for k in U.METRICS}, 20.0, 90.0, 25.0, 5.0)
    assert np.array_equal(f.index.to_numpy(), mdf.index.to_numpy()[ok])
    with pytest.raises(KeyError):
        U.apply_binning_metrics_filter(table)
    # no 'cluster' column -> NA metrics (utilities.py:163-171)
    c2, m2 = U.add_metrics(table.copy(), markers)
    assert c2[U.METRICS].isna().all().all()


def test_api_errors_before_any_gpu_work(tmp_path):
    from autometa_b200.common import kmers
    from autometa_b200.common.exceptions import TableFormatError

    with pytest.raises(TypeError):
        kmers.count("whatever.fna", size=5.5)
    df = pd.DataFrame({"AAA": [1, 2]}, index=pd.Index(["a", "b"], name="contig"))
    with pytest.raises(ValueError):
        kmers.normalize(df, method="am_clrs")
    with pytest.raises(TypeError):
        kmers.embed(kmers=12345)
    with pytest.raises(FileNotFoundError):
        kmers.embed(kmers=pd.DataFrame())
    with pytest.raises(ValueError):
        kmers.embed(kmers=df, method="nope")
    with pytest.raises(FileNotFoundError):
        kmers.load(str(tmp_path / "none.tsv"))
    bad = tmp_path / "bad.tsv"
    bad.write_text("a\tb\n1\t2\n")
    with pytest.raises(TableFormatError):
        kmers.load(str(bad))
    # cached outputs are returned without recomputation (kmers.py:310-314, 406-412)
    cache = tmp_path / "counts.tsv"
    df.to_csv(cache, sep="\t")
    got = kmers.count("does-not-matter.fna", size=5, out=str(cache), force=False)
    assert got.index.name == "contig" and got.shape == (2, 1)
    got = kmers.normalize(df, out=str(cache), force=False)
    assert got.shape == (2, 1)


def test_no_cpu_fallback(monkeypatch):
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from autometa_b200.common import kmers

    df = pd.DataFrame({"AAA": [1, 2], "AAT": [3, 4]}, index=pd.Index(["a", "b"], name="contig"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        kmers.normalize(df, method="am_clr")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        kmers.pca(np.random.rand(20, 4), n_components=2)


def test_synth_is_seeded():
    from autometa_b200 import synth

    a = synth.make_assembly(50, 50 * 3600, seed=9, tiny=True)
    b = synth.make_assembly(50, 50 * 3600, seed=9, tiny=True)
    assert a.ids == b.ids and np.array_equal(a.offsets, b.offsets)
    assert np.array_equal(a.seq[: a.total_bp], b.seq[: b.total_bp])
    assert a.total_bp == 50 * 3600 + 15 and a.lengths[-3:].tolist() == [4, 5, 6]
    assert a.lengths[:-3].min() >= 3000


def test_shard_ranges_balance_and_cover():
    from autometa_b200 import dist as DD, synth

    rng = np.random.default_rng(0)
    lens = synth.contig_lengths(5000, 5000 * 5000, rng)
    for w in (1, 2, 4, 8):
        r = DD.shard_ranges(lens, w)
        assert r[0][0] == 0 and r[-1][1] == len(lens)
        assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
        bp = np.array([lens[a:b].sum() for a, b in r])
        assert bp.max() - bp.min() <= 2 * lens.max()


# ------------------------------------------------------------------ metagenome (8f-2)
def test_gc_fraction_oracle_known_answers():
    # the values Bio.SeqUtils.gc_fraction documents for ambiguous="remove"
    assert O.gc_fraction_remove(b"ACTG") == 0.5
    assert O.gc_fraction_remove(b"ACTGN") == 0.5
    assert O.gc_fraction_remove(b"ACTGSSSS") == 0.75
    assert O.gc_fraction_remove(b"GGAUCUUCGGAUCU") == 0.5
    assert O.gc_fraction_remove(b"NNNN") == 0 and O.gc_fraction_remove(b"") == 0
    assert O.gc_fraction_remove(b"acgtWwSsUu") == 0.4


def test_metagenome_host_members(small_fna, tmp_path):
    from autometa_b200.common.metagenome import Metagenome

    ids, seqs = O.read_fasta(small_fna)
    lengths = [len(s) for s in seqs]
    mg = Metagenome(small_fna)
    assert str(mg) == os.path.realpath(small_fna)
    assert mg.nseqs == len(seqs) and mg.size == sum(lengths)
    assert mg.sequences == [s.decode() for s in seqs]
    assert mg.largest_seq == ids[lengths.index(max(lengths))]
    for q in (0.1, 0.5, 0.9):
        m = mg.fragmentation_metric(q)
        assert isinstance(m, int) and m == O.fragmentation_metric(lengths, q)
    with pytest.raises(ValueError):
        mg.fragmentation_metric(1.4)
    # length_filter: argument checks, force semantics, 60-column FASTA out, original kept when nothing passes
    out = tmp_path / "f.fna"
    with pytest.raises(ValueError):
        mg.length_filter(out=str(out), cutoff=-6)
    with pytest.raises(TypeError):
        mg.length_filter(out=str(out), cutoff="12324")
    with pytest.raises(FileExistsError):
        mg.length_filter(out=small_fna, cutoff=10)
    cutoff = sorted(lengths)[len(lengths) // 2]
    f = mg.length_filter(out=str(out), cutoff=cutoff)
    want = [(i, s) for i, s in zip(ids, seqs) if len(s) >= cutoff]
    assert f.nseqs == len(want) and f.ids == [i for i, _ in want]
    assert f.sequences == [s.decode() for _, s in want]
    lines = open(out).read().split("\n")
    assert max(len(l) for l in lines if not l.startswith(">")) == 60
    assert mg.length_filter(out=str(tmp_path / "g.fna"), cutoff=10 ** 9) is mg
    # header description survives the rewrite the way SeqIO.write keeps it
    p = tmp_path / "d.fa"
    p.write_bytes(b">c1 some description \r\nACGT\nAC\n>c2\nA\n")
    g = Metagenome(str(p)).length_filter(out=str(tmp_path / "d2.fa"), cutoff=2)
    assert open(g.assembly, "rb").read() == b">c1 some description\nACGTAC\n"


def test_bin_name_helpers():
    from autometa_b200.binning.utilities import reindex_bin_names, zero_pad_bin_names

    df = pd.DataFrame({"cluster": [7, 3, 7, pd.NA, 11, 3]}, index=list("abcdef"))
    r = reindex_bin_names(df.copy(), initial_index=4)
    assert r["cluster"].tolist()[:3] == ["bin_5", "bin_6", "bin_5"] and pd.isna(r["cluster"].iloc[3])
    assert r["cluster"].tolist()[4:] == ["bin_7", "bin_6"]
    many = pd.DataFrame({"cluster": [f"bin_{i}" for i in range(1, 13)] + [pd.NA]})
    z = zero_pad_bin_names(many)
    assert z["cluster"].tolist()[:12] == [f"bin_{i:02d}" for i in range(1, 13)] and pd.isna(z["cluster"].iloc[12])


# ------------------------------------------------------------------ TSV I/O (8f-4)
def _bits_equal(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return bool(np.all((a.view(np.int64) == b.view(np.int64)) | (np.isnan(a) & np.isnan(b))))


def test_tsv_writer_is_byte_identical_to_pandas(tmp_path):
    from autometa_b200 import tsv

    rng = np.random.default_rng(0)
    n, d = 3000, 64
    X = rng.normal(size=(n, d))
    X[:200] = np.frombuffer(rng.bytes(200 * d * 8), dtype=np.float64).reshape(200, d)  # every exponent, subnormals
    X[200:400] *= 10.0 ** rng.integers(-25, 25, size=(200, d))
    X[400:600] = np.round(X[400:600] * 1000)  # integral values -> "123.0"
    X[600, :12] = [0.0, -0.0, np.inf, -np.inf, np.nan, 1e16, 1e-5, 123456.0, 1e15, 9999999999999998.0, 1e-4, 5e-324]
    ids = [f"NODE_{i}_length_{i * 7}_cov_{i / 3:.3f}" for i in range(n)]
    ids[7] = 'odd"name'
    for cols in ([f"k{j}" for j in range(d)], list(range(d))):
        df = pd.DataFrame(X, index=pd.Index(ids, name="contig"), columns=cols)
        tsv.write_frame(df, str(tmp_path / "fast.tsv"))
        df.to_csv(tmp_path / "pd.tsv", sep="\t", index=True, header=True)
        assert (tmp_path / "fast.tsv").read_bytes() == (tmp_path / "pd.tsv").read_bytes()
    # count tables: 0 == pd.NA == empty cell
    C = rng.integers(0, 5, size=(500, 32)).astype(np.int32)
    C[3] = 0
    names = [f"K{j}" for j in range(32)]
    tsv.write_counts(str(tmp_path / "c.tsv"), ids[:500], names, C)
    cols = {nm: pd.arrays.IntegerArray(np.ascontiguousarray(C[:, j]), np.ascontiguousarray(C[:, j] == 0))
            for j, nm in enumerate(names)}
    pd.DataFrame(cols, index=pd.Index(ids[:500], name="contig")).to_csv(tmp_path / "c_pd.tsv", sep="\t")
    assert (tmp_path / "c.tsv").read_bytes() == (tmp_path / "c_pd.tsv").read_bytes()
    # frames outside the fast path go through pandas and still come out right
    mixed = pd.DataFrame({"a": [1, 2], "b": [0.5, np.nan]}, index=pd.Index(["x", "y"], name="contig"))
    tsv.write_frame(mixed, str(tmp_path / "m.tsv"))
    assert (tmp_path / "m.tsv").read_text() == mixed.to_csv(sep="\t")


def test_tsv_reader_equals_pandas_bit_for_bit(tmp_path):
    from autometa_b200 import tsv

    rng = np.random.default_rng(1)
    n, d = 4000, 48
    X = rng.normal(size=(n, d))
    X[:300] = np.frombuffer(rng.bytes(300 * d * 8), dtype=np.float64).reshape(300, d)
    X[300:600] *= 10.0 ** rng.integers(-25, 25, size=(300, d))
    X[600:900] = np.round(X[600:900] * 50)
    X[rng.random(X.shape) < 0.05] = np.nan
    X[901, :4] = [np.inf, -np.inf, 0.0, -0.0]
    ids = [f"c{i}" for i in range(n)]
    df = pd.DataFrame(X, index=pd.Index(ids, name="contig"), columns=[f"k{j}" for j in range(d)])
    p = tmp_path / "t.tsv"
    df.to_csv(p, sep="\t")
    a, b = tsv.read_table(str(p)), pd.read_csv(p, sep="\t", index_col="contig")
    pd.testing.assert_frame_equal(a, b, check_exact=True)
    assert _bits_equal(a.to_numpy(), b.to_numpy())
    # digits beyond the 17 pandas keeps, exponents, signs, blanks, CRLF, blank lines, integer columns
    text = ("contig\ti\tf\tg\r\n"
            "a\t12\t0.000123456789012345678901\t1e-320\r\n"
            "\r\n"
            "b\t-7\t123456789012345678901234567890\t+1.5E+10\r\n"
            "c\t+3\t\t-0.1e-5\r\n"
            "d\t0\t9007199254740993\tnan\r\n")
    p2 = tmp_path / "u.tsv"
    p2.write_bytes(text.encode())
    a, b = tsv.read_table(str(p2)), pd.read_csv(p2, sep="\t", index_col="contig")
    pd.testing.assert_frame_equal(a, b, check_exact=True)
    assert a["i"].dtype == np.int64 and _bits_equal(a["f"], b["f"]) and _bits_equal(a["g"], b["g"])
    # tables the fast path declines: same answer through pandas (numeric ids, quoted ids, index not first, gz)
    for body in ("contig\tx\n1\t2.5\n2\t3.5\n", 'contig\tx\n"a\tb"\t1.0\n', "x\tcontig\n1.0\tq\n", "contig\tx\nNA\t1.0\n",
                 "contig\tx\nq\ttrue\n"):
        p3 = tmp_path / "v.tsv"
        p3.write_text(body)
        pd.testing.assert_frame_equal(tsv.read_table(str(p3)), pd.read_csv(p3, sep="\t", index_col="contig"))
    df.iloc[:50].to_csv(tmp_path / "z.tsv.gz", sep="\t")
    pd.testing.assert_frame_equal(tsv.read_table(str(tmp_path / "z.tsv.gz")),
                                  pd.read_csv(tmp_path / "z.tsv.gz", sep="\t", index_col="contig"), check_exact=True)
    # pandas' default converter is not correctly rounded (17 digits, one scaling): the fast reader
    # reproduces its values, which differ from the written doubles in the last bit for some cells
    (tmp_path / "w.tsv").write_text("id\tx\nq\t1.0\n")
    with pytest.raises(ValueError):
        tsv.read_table(str(tmp_path / "w.tsv"))


def test_numa_binding_helper_is_safe_without_a_gpu():
    from autometa_b200 import dist as D

    before = os.sched_getaffinity(0)
    ok = D.bind_to_gpu_numa_node(0)  # no NVML device here: must report False and leave the affinity alone
    assert ok in (False, True)
    if not ok:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)


# ------------------------------------------------------------------ large-data-mode caches and checkpoints (8f-4)
def test_gz_tsv_round_trip_matches_pandas(tmp_path):
    """``.tsv.gz`` (the large-data-mode embedding caches, large_data_mode.py:104,112): decompressed bytes equal
    pandas' ``to_csv``; pandas reads our multi-member gzip; our reader returns pandas' frame."""
    import gzip

    from autometa_b200 import tsv

    rng = np.random.default_rng(3)
    df = pd.DataFrame(rng.normal(size=(4000, 3)) * 10.0 ** rng.integers(-8, 8, size=(4000, 3)),
                      index=pd.Index([f"NODE_{i}_length_{3000 + i}_cov_1.5" for i in range(4000)], name="contig"),
                      columns=["x_1", "x_2", "x_3"])
    ours, theirs = str(tmp_path / "a.tsv.gz"), str(tmp_path / "b.tsv.gz")
    old = tsv._GZ_CHUNK
    tsv._GZ_CHUNK = 50_000  # several gzip members
    try:
        tsv.write_frame(df, ours)
    finally:
        tsv._GZ_CHUNK = old
    df.to_csv(theirs, sep="\t", index=True, header=True)
    with gzip.open(ours, "rb") as a, gzip.open(theirs, "rb") as b:
        assert a.read() == b.read()
    ref = pd.read_csv(theirs, sep="\t", index_col="contig")
    pd.testing.assert_frame_equal(pd.read_csv(ours, sep="\t", index_col="contig"), ref, check_exact=True)
    pd.testing.assert_frame_equal(tsv.read_table(ours), ref, check_exact=True)
    pd.testing.assert_frame_equal(tsv.read_table(theirs), ref, check_exact=True)


def test_ldm_checkpoint_round_trip(tmp_path):
    """checkpoint() / get_checkpoint_info() (large_data_mode.py:117-224): '#' header with the run parameters and
    the rank / taxon the run stopped at, then the table; plain and gzip; columns accumulate per taxon."""
    import gzip

    from autometa_b200.binning import large_data_mode as M

    idx = pd.Index([f"c{i}" for i in range(6)], name="contig")
    first = pd.DataFrame({"cluster": ["bin_0", "bin_0", "bin_1"], "purity": [99.0, 99.0, 97.0]}, index=idx[:3])
    second = pd.DataFrame({"cluster": ["bin_2", "bin_2"]}, index=idx[3:5])
    for name in ("binning_checkpoints.tsv", "binning_checkpoints.tsv.gz"):
        path = str(tmp_path / name)
        kw = dict(completeness=20.0, purity=95.0, coverage_stddev=25.0, gc_content_stddev=5.0, cluster_method="dbscan",
                  norm_method="am_clr", pca_dimensions=50, embed_dimensions=2, embed_method="bhsne", min_contigs=51,
                  max_partition_size=10000, binning_checkpoints_fpath=path)
        ck = M.checkpoint(pd.DataFrame(), first, rank="phylum", rank_name_txt="Proteobacteria", **kw)
        ck = M.checkpoint(ck, second, rank="class", rank_name_txt="Bacilli", **kw)
        assert ck.columns.tolist() == ["Proteobacteria", "Bacilli"]
        info = M.get_checkpoint_info(path)
        assert info["starting_rank"] == "class" and info["starting_rank_name_txt"] == "Bacilli"
        got = info["binning_checkpoints"]
        assert got.columns.tolist() == ["Proteobacteria", "Bacilli"] and got.index.tolist() == idx[:5].tolist()
        assert got.loc["c0", "Proteobacteria"] == "bin_0" and got.loc["c4", "Bacilli"] == "bin_2"
        assert pd.isna(got.loc["c4", "Proteobacteria"]) and pd.isna(got.loc["c0", "Bacilli"])
        opener = gzip.open if name.endswith(".gz") else open
        with opener(path, "rt") as fh:
            head = [next(fh) for _ in range(18)]
        assert head[0].strip() == "#-- Parameters --#" and head[12].strip() == "#-- Runtime Variables --#"
        assert head[10].startswith("# min-partition-size: 51 (max([pca_dimensions + 1, embed_dimensions + 1])")
        assert head[13].strip() == "# rank: class" and head[14].strip() == "# name: Bacilli"


# ----------------------------------------------------------------------------- device-resident sweep (host side)
def _reference_loop(outcomes):
    """The loop body of the reference after the metrics (recursive_dbscan.py:172-215), fed with the
    (n_clusters, median) a step produced; returns the per-step (eps, take, step, done)."""
    eps, step = 0.3, 0.1
    clustering_rounds = {0: 0}
    best_median = float("-inf")
    out = []
    for n_clusters, median_completeness in outcomes:
        this_eps = eps
        take = median_completeness >= best_median
        if take:
            best_median = median_completeness
        if n_clusters in clustering_rounds:
            clustering_rounds[n_clusters] += 1
        else:
            clustering_rounds[n_clusters] = 1
        if clustering_rounds[n_clusters] > 10:
            step *= 10
        if median_completeness == 0:
            clustering_rounds[0] += 1
        done = False
        if clustering_rounds[0] >= 10:
            done = True
        else:
            eps += step
            done = not (n_clusters > 1)
        out.append((this_eps, take, step, done, best_median, eps))
        if done:
            break
    return out


@pytest.mark.parametrize("case", ["plateaus", "zero_medians", "single_cluster_at_once", "random"])
def test_sweep_controller_advances_like_the_reference_loop(case):
    """am_sweep_advance_host runs the same code as the device controller kernel (csrc/am_sweep.cu)."""
    from autometa_b200 import _lib as L

    rng = np.random.default_rng(5)
    if case == "plateaus":      # the same cluster count > 10 times -> step x 10, twice; ends at one cluster
        outcomes = [(50, float("-inf"))] * 3 + [(40, 12.5)] * 14 + [(7, 30.0)] * 12 + [(3, 30.0), (2, 10.0), (1, 0.5)]
    elif case == "zero_medians":  # median == 0 ten times -> break (and counts towards clustering_rounds[0])
        outcomes = [(90 - i, 0.0) for i in range(30)]
    elif case == "single_cluster_at_once":
        outcomes = [(1, float("-inf"))]
    else:
        ncs = np.maximum(1, 200 - np.cumsum(rng.integers(0, 6, size=400)))
        meds = rng.choice([float("-inf"), 0.0, 10.0, 20.0, 35.5], size=400)
        outcomes = list(zip(ncs.tolist(), meds.tolist()))
    want = _reference_loop(outcomes)
    n_max = max(nc for nc, _ in outcomes)
    ctl = np.zeros(L.SWEEP_CTL_WORDS, dtype=np.float64)
    ctl[L.SWEEP_EPS], ctl[L.SWEEP_STEP], ctl[L.SWEEP_BEST], ctl[L.SWEEP_BEST_STEP] = 0.3, 0.1, -np.inf, -1.0
    span = np.array([200.0, 150.0, 500.0, 0.0])
    ctl[L.SWEEP_SPAN : L.SWEEP_SPAN + 4] = span
    rounds = np.zeros(n_max + 1, dtype=np.int32)
    mn, g = np.zeros(4), np.zeros(2)
    for k, (nc, med) in enumerate(outcomes):
        eps_before = ctl[L.SWEEP_EPS]
        take = L.lib.am_sweep_advance_host(L.np_ptr(ctl), L.np_ptr(rounds), 3, int(nc), float(med))
        w_eps, w_take, w_step, w_done, w_best, w_next = want[k]
        assert eps_before == w_eps                      # bit-identical float accumulation
        assert take == int(w_take) and ctl[L.SWEEP_TAKE] == float(w_take)
        assert ctl[L.SWEEP_STEP] == w_step and ctl[L.SWEEP_BEST] == w_best
        assert ctl[L.SWEEP_EPS] == w_next and bool(ctl[L.SWEEP_DONE]) == w_done
        assert ctl[L.SWEEP_NSTEPS] == k + 1
        if w_done:
            break
        # the grid the next step reads = am_eps_grid of the next eps
        L.check(L.lib.am_eps_grid(float(w_next), L.np_ptr(mn), L.np_ptr(span), 3, L.np_ptr(g)))
        assert ctl[L.SWEEP_INV_CELL] == g[0] and ctl[L.SWEEP_EPS2] == g[1]
    assert k + 1 == len(want) and want[-1][3]
    # frozen afterwards: further steps change nothing
    before = ctl.copy()
    assert L.lib.am_sweep_advance_host(L.np_ptr(ctl), L.np_ptr(rounds), 3, 5, 50.0) == 0
    before[L.SWEEP_TAKE] = 0.0
    assert np.array_equal(ctl, before)


def test_sweep_controller_first_48_eps_are_the_reference_schedule():
    from autometa_b200 import _lib as L

    ctl = np.zeros(L.SWEEP_CTL_WORDS, dtype=np.float64)
    ctl[L.SWEEP_EPS], ctl[L.SWEEP_STEP], ctl[L.SWEEP_BEST] = 0.3, 0.1, -np.inf
    rounds = np.zeros(1000, dtype=np.int32)
    got = []
    for k in range(48):
        got.append(ctl[L.SWEEP_EPS])
        L.lib.am_sweep_advance_host(L.np_ptr(ctl), L.np_ptr(rounds), 2, 900 - k, 10.0)
    assert got == O.eps_schedule(48)
    assert got[5] == 0.7999999999999999 and got[47] == 4.999999999999998


def test_eps_grid_matches_the_step_formula():
    from autometa_b200 import _lib as L

    mn = np.array([-3.0, 0.0, 1.0, 0.0])
    mx = np.array([5.0, 2.0e12, 1.0, 0.0])
    g = np.zeros(2)
    for eps in (0.3, 0.7999999999999999, 5.0, 1e3):
        L.check(L.lib.am_eps_grid(eps, L.np_ptr(mn), L.np_ptr(mx), 3, L.np_ptr(g)))
        cell = eps * (1.0 + 1.0 / 1048576.0)
        for d in range(3):
            if (mx[d] - mn[d]) / cell > 1.0e9:
                cell = (mx[d] - mn[d]) / 1.0e9
        assert g[0] == 1.0 / cell and g[1] == eps * eps
    assert L.lib.am_eps_grid(0.0, L.np_ptr(mn), L.np_ptr(mx), 3, L.np_ptr(g)) == L.AM_EINVAL
    assert L.lib.am_eps_grid(1.0, L.np_ptr(mx), L.np_ptr(mn), 3, L.np_ptr(g)) == L.AM_EINVAL
    assert L.lib.am_sweep_ctl_bytes(1000, 0) >= 32 * 8 + 1001 * 4
    assert L.lib.am_sweep_ctl_bytes(1000, 64) - L.lib.am_sweep_ctl_bytes(1000, 0) == 64 * 32


def test_sweep_driver_keeps_order_and_limit():
    """_drive: results in job order, never more than ``limit`` jobs in flight, slots reused."""
    from autometa_b200.binning import recursive_dbscan as R

    class Wait:
        def __init__(self, polls):
            self.polls = polls

        def ready(self):
            self.polls -= 1
            return self.polls <= 0

        def wait(self):
            self.polls = 0

    live, peak, slots_seen = set(), [0], []

    def job(i, slot, n_yields):
        live.add(i)
        slots_seen.append(slot)
        peak[0] = max(peak[0], len(live))
        for y in range(n_yields):
            yield Wait(polls=(i * 7 + y) % 4)
        live.discard(i)
        return ("result", i)

    starters = [(lambda slot, i=i: job(i, slot, n_yields=(i * 3) % 5)) for i in range(11)]
    assert R._drive(starters, limit=3) == [("result", i) for i in range(11)]
    assert peak[0] == 3 and set(slots_seen) == {0, 1, 2}
    assert R._drive(starters[:2], limit=1) == [("result", 0), ("result", 1)]
    assert R._drive([], limit=4) == []
    assert R._in_flight(10_000, 140) == 16 and R._in_flight(2_000_000, 140) >= 1


def test_markers_coo_cached_table_equals_direct_conversion():
    """markers_coo goes through a CSR table built once per marker frame; the entries must be those of the
    direct conversion (restrict the frame to the table's contigs, take the non-zeros), for any subset / order
    of contigs, contigs without markers, NaN cells and an empty selection."""
    from autometa_b200.binning import utilities as U

    rng = np.random.default_rng(3)
    ids = np.array([f"c{i}" for i in range(500)], dtype=object)
    vals = rng.integers(0, 3, size=(300, 17)).astype(float) * (rng.random((300, 17)) < 0.2)
    vals[5, 3] = np.nan
    markers = pd.DataFrame(vals, index=pd.Index(rng.permutation(ids)[:300], name="contig"),
                           columns=[f"PF{m}" for m in range(17)])

    def direct(index):
        pos = index.get_indexer(markers.index)
        ok = pos >= 0
        v = np.nan_to_num(markers.to_numpy(dtype=np.float64, na_value=0.0)[ok], nan=0.0)
        r, c = np.nonzero(v)
        return sorted(zip(pos[ok][r].tolist(), c.tolist(), v[r, c].tolist()))

    for sel in (ids, rng.permutation(ids)[:120], ids[:1], np.array(["zzz"], dtype=object), ids[:0]):
        index = pd.Index(sel, name="contig")
        r, c, v = U.markers_coo(markers, index)
        assert r.dtype == np.int64 and c.dtype == np.int64
        assert sorted(zip(r.tolist(), c.tolist(), v.tolist())) == direct(index)
    assert U._marker_table(markers) is U._marker_table(markers)          # built once per frame
    other = markers.copy()
    assert U._marker_table(other) is not U._marker_table(markers)
    dup = pd.concat([markers.iloc[:3], markers.iloc[:3]])                  # non-unique index: the direct path
    r, c, v = U.markers_coo(dup, pd.Index(ids, name="contig"))
    assert len(r) == 2 * int((np.nan_to_num(markers.iloc[:3].to_numpy()) != 0).sum())


def _check_binning_frame(df, g, prefix):
    assert df.index.tolist() == g[f"{prefix}_index"].tolist()
    assert df.columns.tolist() == g[f"{prefix}_columns"].tolist()
    got = np.array(["" if pd.isna(c) else str(c) for c in df["cluster"]])
    assert np.array_equal(got, g[f"{prefix}_cluster"])
    for m in ["completeness", "purity", "coverage_stddev", "gc_content_stddev"]:
        a = pd.to_numeric(df[m], errors="coerce").to_numpy(dtype=np.float64, na_value=np.nan)
        assert np.allclose(a, g[f"{prefix}_{m}"], rtol=1e-12, atol=0, equal_nan=True), m


def test_binning_drivers_host_logic_against_reference_golden(monkeypatch, golden_binning):
    """get_clusters / get_clusters_many / taxon_guided_binning with the device sweep replaced by the numpy oracle
    (tests/_cpu_sweep_standin.py): everything the host does around a sweep — frames, rounds on the leftovers, bin
    names, sweeps in flight together, the walk over ranks and taxa — against the frames the unmodified reference
    produced (tests/golden/binning_golden.npz).  The same goldens pin the real device path in the GPU suite."""
    import _cpu_sweep_standin as S
    from autometa_b200 import synth
    from autometa_b200.binning import recursive_dbscan as R

    S.install(monkeypatch)
    table, markers = synth.make_sweep_input(n_points=1500, n_clusters=9, n_markers=40, seed=21)
    df = R.get_clusters(table.copy(), markers, 20.0, 90.0, 25.0, 5.0, method="dbscan")
    _check_binning_frame(df, golden_binning, "get_clusters")
    df0 = R.get_clusters(table.iloc[:40].copy(), markers, 99.0, 99.0, 0.001, 0.001, method="dbscan")
    _check_binning_frame(df0, golden_binning, "get_clusters_none")
    both = R.get_clusters_many([table.copy(), table.iloc[:40].copy()], markers, 20.0, 90.0, 25.0, 5.0, method="dbscan")
    _check_binning_frame(both[0], golden_binning, "get_clusters")
    tax = synth.add_taxonomy(table, seed=3)
    dft = R.taxon_guided_binning(tax.copy(), markers, 20.0, 90.0, 25.0, 5.0, method="dbscan")
    _check_binning_frame(dft, golden_binning, "taxon")


def test_large_data_mode_host_logic_against_reference_golden(monkeypatch, tmp_path):
    """cluster_by_taxon_partitioning with the device work replaced by the numpy oracle (sweeps, normalise + PCA):
    the walk over ranks and taxa, partition rules, the batching of a rank's sweeps, bin numbering, embedding
    caches and checkpoints against the frames and file lists the unmodified reference produced
    (tests/golden/ldm_golden.npz); the GPU suite runs the same comparison on the device path."""
    import _cpu_sweep_standin as S
    import _ldm_input
    from autometa_b200.binning import large_data_mode as M

    S.install_ldm(monkeypatch)
    kw = dict(norm_method="am_clr", pca_dimensions=10, embed_dimensions=2, embed_method="bhsne", completeness=20.0,
              purity=90.0, coverage_stddev=25.0, gc_content_stddev=5.0, starting_rank="superkingdom", method="dbscan")
    g = np.load(os.path.join(ROOT, "tests", "golden", "ldm_golden.npz"))
    main, counts, markers = _ldm_input.make_ldm_input()
    df = M.cluster_by_taxon_partitioning(main.copy(), counts.copy(), markers, max_partition_size=10000, **kw)
    _check_binning_frame(df, g, "ldm")
    cache = str(tmp_path / "cache")
    df2 = M.cluster_by_taxon_partitioning(main.copy(), counts.copy(), markers, max_partition_size=400, cache=cache, **kw)
    _check_binning_frame(df2, g, "ldm400")
    files = sorted(os.path.relpath(os.path.join(d, f), cache) for d, _, fs in os.walk(cache) for f in fs)
    assert files == g["ldm400_cache_files"].tolist()
    info = M.get_checkpoint_info(os.path.join(cache, "binning_checkpoints.tsv.gz"))
    assert info["binning_checkpoints"].columns.tolist() == g["ldm400_checkpoint_columns"].tolist()
    assert [info["starting_rank"], info["starting_rank_name_txt"]] == g["ldm400_checkpoint_rank"].tolist()


def test_counts_frame_and_back_through_the_threaded_host_conversions():
    """count()'s frame (one nullable-integer array per k-mer column, zero / uncountable -> pd.NA, duplicate ids
    merged as the reference's dict.update does) is built by am_counts_to_columns and read back for normalize by
    am_columns_to_counts: both against their plain numpy definitions."""
    from autometa_b200.common import kmers
    from autometa_b200.common.exceptions import TableFormatError

    rng = np.random.default_rng(11)
    n, D = 1237, 136
    counts = (rng.integers(0, 40, size=(n, D)) * (rng.random((n, D)) < 0.7)).astype(np.int32)
    countable = rng.random(n) > 0.02
    ids = [f"c{i}" for i in range(n)]
    names = [f"k{j}" for j in range(D)]
    df = kmers._counts_frame(ids, counts, countable, names)
    want_na = (counts == 0) | ~countable[:, None]
    assert df.shape == (n, D) and df.index.name == "contig" and list(df.columns) == names
    assert all(str(t) == "Int32" for t in df.dtypes)
    assert np.array_equal(df.isna().to_numpy(), want_na)
    assert np.array_equal(df.fillna(0).to_numpy(dtype=np.int64), np.where(want_na, 0, counts))
    back, isna = kmers._counts_matrix(df)
    assert back.dtype == np.int32 and back.flags["C_CONTIGUOUS"] and isna.dtype == bool
    assert np.array_equal(back, np.where(want_na, 0, counts)) and np.array_equal(isna, want_na)
    # the same answers as the generic conversion (float / NaN frame, as kmers.load returns)
    fl = df.astype("float64")
    b2, i2 = kmers._counts_matrix(fl)
    assert np.array_equal(b2, back) and np.array_equal(i2, isna)
    # Int64 columns; a frame that is not all nullable integers takes the generic path
    b3, i3 = kmers._counts_matrix(df.astype("Int64"))
    assert np.array_equal(b3, back) and np.array_equal(i3, isna)
    mixed = df.copy()
    mixed[names[0]] = mixed[names[0]].astype("float64")
    b4, i4 = kmers._counts_matrix(mixed)
    assert np.array_equal(b4, back) and np.array_equal(i4, isna)
    bad = df.astype("Int64")
    bad.iloc[5, 7] = -3
    with pytest.raises(TableFormatError):
        kmers._counts_matrix(bad)
    # duplicate ids: the later record's values at the first record's position (kmers.py:155-157)
    ids2 = list(ids)
    ids2[10] = ids2[3]
    d2 = kmers._counts_frame(ids2, counts, countable, names)
    assert d2.shape[0] == n - 1 and d2.index[3] == ids[3] and ids[10] not in d2.index
    assert np.array_equal(d2.iloc[3].fillna(0).to_numpy(dtype=np.int64), np.where(want_na[10], 0, counts[10]))
    # empty table
    e = kmers._counts_frame([], np.zeros((0, D), dtype=np.int32), np.zeros(0, dtype=bool), names)
    assert e.shape == (0, D)


def test_threaded_fasta_index_equals_single_pass(tmp_path):
    """am_fasta_index_mt cuts the buffer at line starts and parses the ranges in parallel (two passes): same ids,
    offsets and sequence bytes as one pass on one thread, on inputs with CRLF, blanks inside and after lines,
    empty lines, text before the first header, '>' inside titles and no final newline; and the one-thread pass
    equals the oracle's FASTA restatement."""
    from autometa_b200 import _lib as L

    rng = np.random.default_rng(77)
    alphabet = np.frombuffer(b"ACGTacgtNRY", dtype=np.uint8)

    def make(n_records, lead=b"", final_newline=True):
        parts = [lead]
        for i in range(n_records):
            title = b">" + (b" " if i % 97 == 0 else b"") + b"rec%d" % i + (b" some > text\tmore" if i % 5 == 0 else b"")
            parts.append(title + (b"\r\n" if i % 3 == 0 else b"\n"))
            for _ in range(int(rng.integers(0, 6))):
                ln = int(rng.integers(0, 90))
                line = alphabet[rng.integers(0, len(alphabet), size=ln)].tobytes()
                if ln > 10 and rng.random() < 0.1:
                    line = line[:5] + b" " + line[5:]
                tail = [b"", b" ", b"\r", b" \t ", b"\r"][int(rng.integers(0, 5))]
                parts.append(line + tail + b"\n")
            if rng.random() < 0.05:
                parts.append(b"\n")
        raw = b"".join(parts)
        return raw if final_newline else raw.rstrip(b"\n")

    def index(raw, threads):
        buf = np.frombuffer(raw, dtype=np.uint8)
        cap = max(16, raw.count(b">"))
        seq = np.zeros(len(buf) + 64, dtype=np.uint8)
        offsets = np.zeros(cap + 1, dtype=np.int64)
        idb, ide = np.zeros(cap, dtype=np.int64), np.zeros(cap, dtype=np.int64)
        n = np.zeros(1, dtype=np.int64)
        L.check(L.lib.am_fasta_index_mt(L.np_ptr(buf), len(buf), L.np_ptr(seq), L.np_ptr(offsets), L.np_ptr(idb),
                                        L.np_ptr(ide), cap, L.np_ptr(n), threads))
        k = int(n[0])
        ids = [raw[int(b):int(e)] for b, e in zip(idb[:k], ide[:k])]
        return ids, offsets[: k + 1].copy(), seq[: int(offsets[k])].tobytes()

    cases = [make(40_000), make(30_000, lead=b"junk line\nACGT\n"), make(25_000, final_newline=False)]
    assert all(len(c) > 3 << 20 for c in cases)  # several 1 MB ranges each
    for raw in cases:
        one = index(raw, 1)
        for threads in (2, 5, 8):
            many = index(raw, threads)
            assert many[0] == one[0] and np.array_equal(many[1], one[1]) and many[2] == one[2]
    small = make(300, lead=b"not a record\n")
    path = tmp_path / "x.fna"
    path.write_bytes(small)
    ids, seqs = O.read_fasta(str(path))
    got = index(small, 1)
    assert [i.decode() for i in got[0]] == ids
    assert [got[2][a:b] for a, b in zip(got[1][:-1], got[1][1:])] == seqs
