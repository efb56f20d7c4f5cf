"""The numpy oracle against the golden vectors produced by the reference itself
(tests/golden/make_golden.py), plus the known answers of SURVEY.md Appendix A."""
import numpy as np
import pytest

from oracle import oracle_np as O


def _seqs(path):
    return O.read_fasta(path)


def test_column_order_known_answers():
    names, lut = O.kmer_columns(5)
    assert len(names) == 512
    assert names[:8] == ["AAAAA", "AAAAT", "AAAAC", "AAAAG", "AAATA", "AAATT", "AAATC", "AAATG"]
    assert names[-3:] == ["GCGCC", "GGACC", "GGCCC"]
    assert [len(O.kmer_columns(k)[0]) for k in range(1, 8)] == [2, 10, 32, 136, 512, 2080, 8192]


def test_columns_match_reference_golden(golden_kmers):
    names, _ = O.kmer_columns(5)
    assert names == golden_kmers["count_columns"].tolist()


@pytest.mark.parametrize("k", [3, 4, 5, 6])
def test_counts_match_reference(golden_kmers, small_fna, k):
    ids, seqs = _seqs(small_fna)
    counts, ok = O.count_sequences(seqs, k)
    # duplicate id: the later record overwrites (kmers.py:155-157)
    gold_ids = golden_kmers["count_ids"].tolist()
    last = {cid: i for i, cid in enumerate(ids)}
    rows = [last[c] for c in gold_ids]
    assert np.array_equal(counts[rows], golden_kmers[f"count_k{k}"])


def test_count_known_answers():
    rng = np.random.default_rng(0)
    names, lut = O.kmer_columns(5)

    def rnd(L):
        return np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=L)].tobytes()

    for L, expect in [(5, 0), (6, 1), (7, 2), (40, 35), (2000, 1995)]:
        c, ok = O.count_sequence(rnd(L), 5, lut, 512)
        assert c.sum() == expect and ok == (L > 5)
    s = bytearray(rnd(300))
    s[100] = ord("N")
    s[150:162] = bytes(s[150:162]).lower()
    c, _ = O.count_sequence(bytes(s), 5, lut, 512)
    assert c.sum() == 274


def test_am_clr_matches_reference(golden_kmers):
    X = O.counts_to_float_frame(golden_kmers["count_k5"].astype(np.int64))
    got, row_keep, col_keep = O.am_clr(X)
    ref = golden_kmers["am_clr_values"]
    assert got.shape == ref.shape
    assert np.array(golden_kmers["count_ids"])[row_keep].tolist() == golden_kmers["am_clr_index"].tolist()
    assert np.max(np.abs(got - ref)) <= 1e-13 * np.max(np.abs(ref))
    # log1p form used by the kernel is the same function
    Xr = np.nan_to_num(X[row_keep][:, col_keep])
    alt = np.log1p(Xr) - np.log1p(Xr).mean(axis=1, keepdims=True)
    assert np.max(np.abs(alt - ref)) <= 1e-12 * np.max(np.abs(ref))


def test_am_clr_dropped_columns(golden_kmers):
    ids = golden_kmers["count_ids"].tolist()
    rows = [ids.index(r) for r in golden_kmers["am_clr6_rows"].tolist()]
    all_rows = list(range(5, 12))
    X = O.counts_to_float_frame(golden_kmers["count_k6"][all_rows].astype(np.int64))
    got, row_keep, col_keep = O.am_clr(X)
    names6, _ = O.kmer_columns(6)
    assert np.array(names6)[col_keep].tolist() == golden_kmers["am_clr6_columns"].tolist()
    assert [all_rows[i] for i in np.nonzero(row_keep)[0]] == rows
    assert col_keep.sum() < 2080
    ref = golden_kmers["am_clr6_values"]
    assert np.max(np.abs(got - ref)) <= 1e-13 * np.max(np.abs(ref))


@pytest.mark.parametrize("method", ["clr", "ilr"])
def test_clr_ilr_match_reference_call_site(golden_kmers, method):
    ids = golden_kmers["count_ids"].tolist()
    rows = [ids.index(r) for r in golden_kmers["clr_rows"].tolist()]
    X = golden_kmers["count_k5"][rows].astype(np.float64)
    got = getattr(O, method)(X)
    ref = golden_kmers[f"{method}_values"]
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref))


def test_pca_matches_sklearn_golden(golden_pca):
    X = golden_pca["X"]
    P = golden_pca["scores"].shape[1]
    r = O.pca_cov_eigh(X, P)
    ev = golden_pca["explained_variance"]
    assert np.max(np.abs(r["explained_variance"] - ev) / ev) < 1e-10
    assert np.max(np.abs(r["explained_variance_ratio"] - golden_pca["explained_variance_ratio"])) < 1e-12
    cos = np.sum(r["components"] * golden_pca["components"], axis=1)
    assert np.all(cos > 1 - 1e-9)  # same sign rule
    assert np.max(np.abs(r["scores"] - golden_pca["scores"])) < 1e-9 * np.max(np.abs(golden_pca["scores"]))


def _sweep_table():
    from autometa_b200 import synth

    return synth.make_sweep_input(n_points=3000, n_clusters=12, n_markers=40, seed=11)


def test_dbscan_labels_match_reference(golden_dbscan):
    table, _ = _sweep_table()
    X = np.ascontiguousarray(table[["x_1", "x_2", "coverage"]].to_numpy())
    for i, eps in enumerate(golden_dbscan["eps_list"]):
        lab = O.dbscan_labels(X, float(eps))
        assert np.array_equal(lab, golden_dbscan[f"labels_{i}"]), eps
    sub = X[:400]
    for eps in (0.5, 2.0):
        assert np.array_equal(O.dbscan_labels(sub, eps), O.dbscan_labels_bruteforce(sub, eps))


def test_cluster_metrics_match_reference(golden_dbscan):
    table, markers = _sweep_table()
    X = np.ascontiguousarray(table[["x_1", "x_2", "coverage"]].to_numpy())
    lab = O.dbscan_labels(X, 1.2)
    mk = markers.reindex(table.index).fillna(0).to_numpy().astype(np.int64)
    m = O.cluster_metrics(lab, mk, table["coverage"].to_numpy(), table["gc_content"].to_numpy(), markers.shape[1])
    for name in ["completeness", "purity", "coverage_stddev", "gc_content_stddev"]:
        np.testing.assert_allclose(m[name], golden_dbscan[f"metrics_{name}"], rtol=1e-12, atol=1e-12, equal_nan=True)
    for name in ["present_marker_count", "single_copy_marker_count"]:
        assert np.array_equal(m[name], golden_dbscan[f"metrics_{name}"].astype(np.int64))


def test_eps_schedule_known_answers():
    s = O.eps_schedule(12)
    assert s[5] == 0.7999999999999999 and s[8] == 1.0999999999999999 and s[11] == 1.4000000000000001


def test_c_port_matches_numpy_oracle_and_golden(golden_kmers, small_fna):
    from autometa_b200 import synth
    from oracle import oracle_c

    ids, seqs = O.read_fasta(small_fna)
    offsets = np.concatenate([[0], np.cumsum([len(s) for s in seqs])]).astype(np.int64)
    seq = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    for k in (3, 5, 6):
        ref, _ = O.count_sequences(seqs, k)
        for threads in (1, 3):
            assert np.array_equal(oracle_c.count(seq, offsets, k, threads), ref)
    asm = synth.make_assembly(80, 80 * 3500, seed=4, tiny=True)
    ref, ok = O.count_sequences(synth.sequences(asm), 5)
    got = oracle_c.count(asm.seq, asm.offsets, 5, 4)
    assert np.array_equal(got, ref)
    keep = ref.sum(axis=1) > 0
    a = oracle_c.am_clr(got[keep], 2)
    b, _, _ = O.am_clr(O.counts_to_float_frame(ref))
    assert np.max(np.abs(a - b)) <= 1e-13 * np.max(np.abs(b))


# ----------------------------------------------------------------------------- scikit-bio published vectors
# scikit-bio is not installable here (no wheel, no network), so clr / ilr / multiplicative_replacement are
# restated from its documentation (oracle_np.py).  These are the known answers scikit-bio itself publishes in
# the doctests of skbio.stats.composition (0.5.x / 0.6.x: closure, multiplicative_replacement, clr, ilr) — the
# same functions the reference calls at kmers.py:23,418-421 — so the restatement is pinned to the library's
# own published vectors, not only to itself.
SKBIO_CLOSURE_IN = np.array([[2, 2, 6], [4, 4, 2]], dtype=np.float64)
SKBIO_CLOSURE_OUT = np.array([[0.2, 0.2, 0.6], [0.4, 0.4, 0.2]])
SKBIO_MR_IN = np.array([[0.2, 0.4, 0.4, 0.0], [0.0, 0.5, 0.5, 0.0]])
SKBIO_MR_OUT = np.array([[0.1875, 0.375, 0.375, 0.0625], [0.0625, 0.4375, 0.4375, 0.0625]])
SKBIO_X = np.array([0.1, 0.3, 0.4, 0.2])
SKBIO_CLR = np.array([-0.79451346, 0.30409883, 0.5917809, -0.10136628])
SKBIO_ILR = np.array([-0.7768362, -0.68339802, 0.11704769])


def test_skbio_published_vectors():
    # closure (the first step of multiplicative_replacement / clr / ilr)
    mr = O.multiplicative_replacement(SKBIO_CLOSURE_IN)  # no zeros: reduces to closure
    np.testing.assert_allclose(mr, SKBIO_CLOSURE_OUT, rtol=0, atol=1e-15)
    # multiplicative_replacement, default delta = (1/D)^2
    np.testing.assert_allclose(O.multiplicative_replacement(SKBIO_MR_IN), SKBIO_MR_OUT, rtol=0, atol=1e-15)
    np.testing.assert_allclose(O.multiplicative_replacement(SKBIO_MR_IN).sum(axis=1), 1.0, atol=1e-15)
    # clr / ilr doctest vectors (8 printed digits)
    np.testing.assert_allclose(O.clr(SKBIO_X[None, :])[0], SKBIO_CLR, rtol=0, atol=5e-9)
    np.testing.assert_allclose(O.ilr(SKBIO_X[None, :])[0], SKBIO_ILR, rtol=0, atol=5e-9)
    # the call site hands over COUNTS (kmers.py:418-421): the same composition as integers
    np.testing.assert_allclose(O.clr(np.array([[1.0, 3.0, 4.0, 2.0]]))[0], SKBIO_CLR, rtol=0, atol=5e-9)
    np.testing.assert_allclose(O.ilr(np.array([[10.0, 30.0, 40.0, 20.0]]))[0], SKBIO_ILR, rtol=0, atol=5e-9)
    # ilr basis: orthonormal rows, orthogonal to the ones vector (the documented Gram-Schmidt basis)
    B = O.ilr_basis(4)
    np.testing.assert_allclose(B @ B.T, np.eye(3), atol=1e-15)
    np.testing.assert_allclose(B.sum(axis=1), 0.0, atol=1e-15)
    with pytest.raises(ValueError):
        O.multiplicative_replacement(np.array([[0.0, 0.0, 0.0]]))
