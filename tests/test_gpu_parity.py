"""Parity tests proper (-m gpu): the CUDA path, called through the C ABI and the
reference-shaped Python functions, against the oracle, the reference's golden
outputs and size-independent properties at BASELINE.json's full sizes.

Tolerances (BASELINE.json north_star / SURVEY.md §8d):
  counts            bit-exact
  normalised        max|d| <= 1e-12 * max|row|   (fp64)
  PCA               explained variance within 1e-6 relative; components equal up to the
                    sklearn sign rule (|cos| >= 1 - 1e-9); scores within 1e-8 * max|score|
  DBSCAN            label vectors identical (min_samples=1: no border points exist)
"""
import os

import numpy as np
import pandas as pd
import pytest

from oracle import oracle_np as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    return torch


# ----------------------------------------------------------------------------- count
def test_count_api_matches_reference_golden(torch_cuda, golden_kmers, small_fna, tmp_path):
    from autometa_b200.common import kmers

    out = tmp_path / "kmers.tsv"
    df = kmers.count(small_fna, size=5, out=str(out), force=True, cpus=2)
    assert df.index.name == "contig"
    assert df.index.tolist() == golden_kmers["count_ids"].tolist()
    assert df.columns.tolist() == golden_kmers["count_columns"].tolist()
    got = df.fillna(0).to_numpy().astype(np.int64)
    ref = golden_kmers["count_k5"].astype(np.int64)
    assert np.array_equal(got, ref)
    assert np.array_equal(df.isna().to_numpy(), ref == 0)  # zero counts are NA (kmers.py:205)
    # TSV round trip = what kmers.load gives the reference's next stage
    back = kmers.load(str(out))
    assert np.array_equal(np.nan_to_num(back.to_numpy(dtype=float)), ref)
    assert np.array_equal(np.isnan(back.to_numpy(dtype=float)), ref == 0)
    # out exists and not force -> read back (kmers.py:310-314)
    again = kmers.count("nonexistent.fna", size=5, out=str(out), force=False)
    assert again.shape == df.shape


@pytest.mark.parametrize("k", [1, 2, 3, 4, 6, 7])
def test_count_other_k(torch_cuda, golden_kmers, small_fna, k):
    from autometa_b200 import assembly as A, device as D

    h = A.read_fasta(small_fna)
    packed = A.pack_to_device(h, device="cuda:0")
    counts, present, row_sum = D.count_packed(packed, k)
    got = counts.cpu().numpy()
    _, seqs = O.read_fasta(small_fna)
    ref, ok = O.count_sequences(seqs, k)
    assert np.array_equal(got, ref)
    assert np.array_equal(row_sum.cpu().numpy(), ref.sum(axis=1))
    assert np.array_equal(present.cpu().numpy() > 0, ref.sum(axis=0) > 0)
    if f"count_k{k}" in golden_kmers:
        ids = golden_kmers["count_ids"].tolist()
        last = {c: i for i, c in enumerate(h.ids)}
        assert np.array_equal(got[[last[c] for c in ids]], golden_kmers[f"count_k{k}"])


def test_device_packer_equals_host_packer(torch_cuda, small_fna):
    from autometa_b200 import assembly as A, synth

    for h in (A.read_fasta(small_fna), synth.make_assembly(300, 300 * 3400, seed=2, tiny=True)):
        starts, lengths, total, codes, bitmap, rank, masks = A.pack_host_arrays(h, 2)
        p = A.pack_to_device(h, device="cuda:0", on_device=True)
        assert p.total_packed_bases == total and p.n_exc == len(masks)
        assert np.array_equal(p.codes.cpu().numpy().view(np.uint32), codes)
        assert np.array_equal(p.exc_bitmap.cpu().numpy().view(np.uint32), bitmap)
        assert np.array_equal(p.exc_rank.cpu().numpy().view(np.uint32), rank)
        assert np.array_equal(p.exc_mask.cpu().numpy().view(np.uint64)[: p.n_exc], masks)


def test_count_synthetic_vs_oracle_and_sharding_invariance(torch_cuda):
    from autometa_b200 import assembly as A, device as D, dist as DD, synth

    asm = synth.make_assembly(1500, 1500 * 4000, seed=20261019, tiny=True)
    ref, ok = O.count_sequences(synth.sequences(asm), 5)
    for on_device in (False, True):
        packed = A.pack_to_device(asm, device="cuda:0", on_device=on_device)
        counts, present, row_sum = D.count_packed(packed, 5)
        assert np.array_equal(counts.cpu().numpy(), ref)
    # sharding by bp never changes a row (the multi-GPU path counts shards independently)
    parts = []
    for lo, hi in DD.shard_ranges(asm.lengths, 3):
        p = A.pack_to_device(asm.slice(lo, hi), device="cuda:0", on_device=True)
        parts.append(D.count_packed(p, 5)[0].cpu().numpy())
    assert np.array_equal(np.concatenate(parts), ref)


def test_count_empty_and_degenerate(torch_cuda):
    from autometa_b200 import assembly as A, device as D

    h = A.from_sequences(["e", "t", "x"], [b"", b"ACGTA", b"ACGTAC"])
    packed = A.pack_to_device(h, device="cuda:0", on_device=True)
    counts, present, row_sum = D.count_packed(packed, 5)
    c = counts.cpu().numpy()
    assert c[:2].sum() == 0 and c[2].sum() == 1
    h0 = A.from_sequences([], [])
    p0 = A.pack_to_device(h0, device="cuda:0")
    c0, _, _ = D.count_packed(p0, 5)
    assert c0.shape == (0, 512)


def _valid_window_counts(asm, k):
    """rows sums expected from the raw bytes, vectorised (independent of the oracle's histogramming)."""
    seq = asm.seq[: asm.total_bp]
    good = (seq == 65) | (seq == 67) | (seq == 71) | (seq == 84)
    bad_prefix = np.concatenate([[0], np.cumsum(~good, dtype=np.int64)])
    out = np.zeros(asm.n_contigs, dtype=np.int64)
    off = asm.offsets
    for i in range(asm.n_contigs):
        a, b = int(off[i]), int(off[i + 1])
        nwin = b - a - k
        if nwin <= 0:
            continue
        starts = np.arange(a, a + nwin)
        out[i] = np.count_nonzero(bad_prefix[starts + k] - bad_prefix[starts] == 0)
    return out


def test_count_full_size_properties(torch_cuda):
    """C2 (200k contigs / 1 Gbp): row sums equal the number of all-ACGT windows among the
    first L-k (checked exactly on a contig sample and in aggregate), linearity across shards."""
    import torch

    from autometa_b200 import assembly as A, device as D, synth

    asm = synth.make_assembly(200_000, 1_000_000_000, seed=20261019 + 2)
    packed = A.pack_to_device(asm, device="cuda:0", on_device=True)
    counts, present, row_sum = D.count_packed(packed, 5)
    rs = row_sum.cpu().numpy().astype(np.int64)
    assert np.array_equal(counts.sum(dim=1, dtype=torch.int64).cpu().numpy(), rs)
    assert int(present.min().item()) == 1
    lens = asm.lengths
    assert np.all(rs <= lens - 5) and rs.sum() > 0.99 * (lens - 5).sum()
    sample = np.random.default_rng(0).choice(asm.n_contigs, size=300, replace=False)
    sub = A.HostAssembly([asm.ids[i] for i in sample], asm.seq, asm.offsets)  # reuse bytes
    exp = np.zeros(len(sample), dtype=np.int64)
    for j, i in enumerate(sample):
        one = A.HostAssembly([asm.ids[i]], asm.seq[int(asm.offsets[i]) : int(asm.offsets[i + 1])],
                             np.array([0, lens[i]], dtype=np.int64))
        exp[j] = _valid_window_counts(one, 5)[0]
    assert np.array_equal(rs[sample], exp)
    # oracle on a slice of whole contigs
    seqs = [asm.seq[int(asm.offsets[i]) : int(asm.offsets[i + 1])].tobytes() for i in sample[:60]]
    ref, _ = O.count_sequences(seqs, 5)
    assert np.array_equal(counts[torch.from_numpy(sample[:60]).cuda()].cpu().numpy(), ref)


# ------------------------------------------------------------------------- normalise
def _golden_count_frame(golden_kmers, k=5):
    X = golden_kmers[f"count_k{k}"].astype(np.float64)
    X[X == 0] = np.nan
    names, _ = O.kmer_columns(k)
    return pd.DataFrame(X, index=pd.Index(golden_kmers["count_ids"].tolist(), name="contig"), columns=names)


def test_normalize_am_clr_matches_reference_golden(torch_cuda, golden_kmers, tmp_path):
    from autometa_b200.common import kmers

    df = _golden_count_frame(golden_kmers)
    out = tmp_path / "norm.tsv"
    norm = kmers.normalize(df, method="am_clr", out=str(out), force=True)
    ref = golden_kmers["am_clr_values"]
    assert norm.index.tolist() == golden_kmers["am_clr_index"].tolist()
    assert norm.columns.tolist() == golden_kmers["am_clr_columns"].tolist()
    assert norm.index.name == "contig"
    got = norm.to_numpy()
    assert np.max(np.abs(got - ref) / np.max(np.abs(ref), axis=1, keepdims=True)) <= 1e-12
    assert out.exists()
    # dropped columns (k=6 sub-table where many 6-mers never occur)
    d6 = _golden_count_frame(golden_kmers, 6).iloc[5:12]
    n6 = kmers.normalize(d6, method="am_clr")
    assert n6.columns.tolist() == golden_kmers["am_clr6_columns"].tolist()
    assert n6.index.tolist() == golden_kmers["am_clr6_rows"].tolist()
    r6 = golden_kmers["am_clr6_values"]
    assert np.max(np.abs(n6.to_numpy() - r6) / np.max(np.abs(r6), axis=1, keepdims=True)) <= 1e-12


@pytest.mark.parametrize("method", ["clr", "ilr"])
def test_normalize_clr_ilr_match_reference_call_site(torch_cuda, golden_kmers, method):
    from autometa_b200.common import kmers

    df = _golden_count_frame(golden_kmers).loc[golden_kmers["clr_rows"].tolist()]
    norm = kmers.normalize(df, method=method)
    ref = golden_kmers[f"{method}_values"]
    assert norm.shape == ref.shape
    assert list(norm.columns) == list(range(ref.shape[1]))  # integer labels (kmers.py:422)
    assert np.max(np.abs(norm.to_numpy() - ref) / np.max(np.abs(ref), axis=1, keepdims=True)) <= 1e-12
    with pytest.raises(ValueError):  # skbio closure(): all-zero row
        kmers.normalize(_golden_count_frame(golden_kmers), method=method)


def test_normalize_synthetic_all_methods(torch_cuda):
    import torch

    from autometa_b200 import device as D

    rng = np.random.default_rng(5)
    counts = rng.poisson(9.0, size=(3000, 512)).astype(np.int32)
    counts[rng.random(counts.shape) < 0.05] = 0
    counts[7, :40] = rng.integers(2049, 90000, size=40)  # beyond the log table
    d = torch.from_numpy(counts).cuda()
    cs = torch.zeros(512, dtype=torch.float64, device="cuda")
    got = D.normalize_counts(d, "am_clr", colsum=cs).cpu().numpy()
    ref, _, _ = O.am_clr(O.counts_to_float_frame(counts.astype(np.int64)))
    assert np.max(np.abs(got - ref) / np.max(np.abs(ref), axis=1, keepdims=True)) <= 1e-12
    np.testing.assert_allclose(cs.cpu().numpy(), ref.sum(axis=0), rtol=0, atol=1e-9)
    for method, fn in (("clr", O.clr), ("ilr", O.ilr)):
        got = D.normalize_counts(d, method).cpu().numpy()
        ref = fn(counts.astype(np.float64))
        assert got.shape == ref.shape
        assert np.max(np.abs(got - ref) / np.max(np.abs(ref), axis=1, keepdims=True)) <= 1e-12
    # gathers
    ridx = torch.tensor([5, 1, 2999, 7], dtype=torch.int64, device="cuda")
    cidx = torch.arange(0, 512, 3, dtype=torch.int32, device="cuda")
    got = D.normalize_counts(d, "am_clr", ridx, cidx).cpu().numpy()
    sub = counts[[5, 1, 2999, 7]][:, ::3].astype(np.float64)
    ref = np.log1p(sub) - np.log1p(sub).mean(axis=1, keepdims=True)
    assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref))


# ------------------------------------------------------------------------------- PCA
def _check_pca(res, ref, score_tol=1e-8):
    ev, rev = res["explained_variance"], ref["explained_variance"]
    assert np.max(np.abs(ev - rev) / rev) <= 1e-6
    rr = ref["explained_variance_ratio"]
    assert np.max(np.abs(res["explained_variance_ratio"] - rr) / rr) <= 1e-6
    cos = np.sum(res["components"] * ref["components"], axis=1)
    assert np.all(cos >= 1 - 1e-9), cos.min()
    assert np.max(np.abs(res["scores"] - ref["scores"])) <= score_tol * np.max(np.abs(ref["scores"]))


def test_pca_matches_sklearn_golden(torch_cuda, golden_pca):
    from autometa_b200.common import kmers

    P = golden_pca["scores"].shape[1]
    res = kmers.pca(golden_pca["X"], n_components=P)
    _check_pca(res, golden_pca)
    assert np.allclose(res["scores"][:, :3], golden_pca["embed_first3"], rtol=0,
                       atol=1e-8 * np.abs(golden_pca["scores"]).max())
    with pytest.raises(ValueError):
        kmers.pca(golden_pca["X"][:5], n_components=10)


def test_pca_on_kmer_features_vs_oracle(torch_cuda):
    """512-column am_clr features of a synthetic assembly (N >= 10*D: sklearn's covariance_eigh branch)."""
    from sklearn.decomposition import PCA

    from autometa_b200 import assembly as A, pipeline, synth

    asm = synth.make_assembly(6000, 6000 * 3600, seed=77, tiny=True)
    packed = A.pack_to_device(asm, device="cuda:0", on_device=True)
    res = pipeline.featurise(packed, 5, "am_clr", 50)
    X = res["normalized"].cpu().numpy()
    # am_clr drops the tiny contigs of length 4 and 5 (no window: all-NA rows, kmers.py:187-192,369);
    # the length-6 one has exactly one window (L-k = 1) and stays
    assert X.shape == (6001, 512)
    ref_counts, _ = O.count_sequences(synth.sequences(asm), 5)
    assert np.array_equal(res["counts"].cpu().numpy(), ref_counts)
    refX, _, _ = O.am_clr(O.counts_to_float_frame(ref_counts))
    assert np.max(np.abs(X - refX) / np.max(np.abs(refX), axis=1, keepdims=True)) <= 1e-12
    p = PCA(n_components=50, random_state=np.random.RandomState(42))
    sc = p.fit_transform(refX)
    assert p._fit_svd_solver == "covariance_eigh"
    ref = dict(explained_variance=p.explained_variance_, explained_variance_ratio=p.explained_variance_ratio_,
               components=p.components_, scores=sc)
    got = {k: res[k].cpu().numpy() for k in ref}
    _check_pca(got, ref)


def test_eigh_against_numpy(torch_cuda):
    import torch

    from autometa_b200 import device as D

    rng = np.random.default_rng(1)
    for n in (2, 5, 64, 511, 512):
        A_ = rng.normal(size=(n, n + 3))
        S = A_ @ A_.T / n
        w, V = D.eigh_desc(torch.from_numpy(S).cuda())
        w = w.cpu().numpy()
        V = V.cpu().numpy()
        wr = np.linalg.eigvalsh(S)[::-1]
        assert np.max(np.abs(w - wr)) <= 1e-12 * wr[0]
        assert np.max(np.abs(V @ V.T - np.eye(n))) < 1e-12
        assert np.max(np.abs(S @ V.T - V.T * w)) <= 1e-11 * wr[0]
        idx = np.argmax(np.abs(V), axis=1)
        assert np.all(V[np.arange(n), idx] > 0)  # svd_flip(u_based_decision=False)


def test_clr_ilr_kernels_match_skbio_published_vectors(torch_cuda):
    """The normalise kernels on the integer compositions of scikit-bio's own doctest vectors
    (tests/test_oracle.py: SKBIO_*): clr / ilr of [1, 3, 4, 2] are the published clr / ilr of
    [.1, .3, .4, .2]; rows with zeros go through multiplicative_replacement's published example."""
    import torch

    from test_oracle import SKBIO_CLR, SKBIO_ILR, SKBIO_MR_OUT

    from autometa_b200 import device as D

    counts = torch.tensor([[1, 3, 4, 2], [10, 30, 40, 20], [2, 4, 4, 0], [0, 5, 5, 0]], dtype=torch.int32).cuda()
    clr = D.normalize_counts(counts, "clr").cpu().numpy()
    ilr = D.normalize_counts(counts, "ilr").cpu().numpy()
    for r in (0, 1):
        np.testing.assert_allclose(clr[r], SKBIO_CLR, rtol=0, atol=5e-9)
        np.testing.assert_allclose(ilr[r], SKBIO_ILR, rtol=0, atol=5e-9)
    for r, q in ((2, SKBIO_MR_OUT[0]), (3, SKBIO_MR_OUT[1])):
        lq = np.log(q)
        np.testing.assert_allclose(clr[r], lq - lq.mean(), rtol=0, atol=1e-14)
        np.testing.assert_allclose(ilr[r], (lq - lq.mean()) @ O.ilr_basis(4).T, rtol=0, atol=1e-14)


def test_eigh_topk_direct(torch_cuda):
    """am_eigh_topk itself (the production solver: cluster tridiagonalisation, bisection, inverse iteration,
    re-orthonormalisation, WY back-transformation) against numpy.linalg.eigh: every kernel branch by shape
    (D < 19: reflector-pair back-transformation; D <= 512: register tridiagonalisation; 512 < D <= 576:
    shared-memory tridiagonalisation), P up to 128, repeated and tightly clustered eigenvalues (the CholeskyQR2
    branch), rank-deficient input, and no silent downgrade to the Jacobi solver."""
    import torch

    from autometa_b200 import device as D

    rng = np.random.default_rng(7)

    def spectrum_matrix(n, lam):
        Q, _ = np.linalg.qr(rng.normal(size=(n, n)))
        A = (Q * lam) @ Q.T
        return 0.5 * (A + A.T)

    def check(A, P, tag, sep_tol=1e-6):
        n = A.shape[0]
        w, V, tr = D.eigh_topk(torch.from_numpy(A).cuda(), P)
        w, V, tr = w.cpu().numpy(), V.cpu().numpy(), float(tr.item())
        wr, Vr = np.linalg.eigh(A)
        wr, Vr = wr[::-1], Vr[:, ::-1].T
        scale = max(abs(wr[0]), abs(wr[-1]), 1e-300)
        assert np.max(np.abs(w - wr[:P])) <= 1e-12 * scale, tag
        assert abs(tr - np.trace(A)) <= 1e-12 * max(abs(np.trace(A)), scale), tag
        assert np.max(np.abs(V @ V.T - np.eye(P))) < 1e-11, tag
        assert np.max(np.abs(A @ V.T - V.T * w)) <= 1e-11 * scale, tag
        idx = np.argmax(np.abs(V), axis=1)
        assert np.all(V[np.arange(P), idx] > 0), tag  # svd_flip(u_based_decision=False)
        # isolated eigenvalues: the vector itself must match (up to nothing: the sign rule is applied)
        gaps = np.minimum(np.abs(np.diff(wr, prepend=np.inf)), np.abs(np.diff(wr, append=-np.inf)))[:P]
        iso = gaps > sep_tol * scale
        if iso.any():
            cos = np.abs(np.sum(V[iso] * Vr[:P][iso], axis=1))
            assert np.all(cos >= 1 - 1e-9), (tag, cos.min())

    for n in (3, 18, 19, 64, 511, 512, 576):
        B = rng.normal(size=(n, n + 5))
        S = B @ B.T / n
        for P in (1, 50, 128):
            if P <= n:
                check(S, P, f"wishart n={n} P={P}")
    for n in (64, 512):
        # a triple eigenvalue at the top, a cluster 1e-11 apart around rank 10, then a smooth decay
        lam = np.sort(np.concatenate([[9.0, 9.0, 9.0], 5.0 + 1e-11 * np.arange(4), rng.uniform(0.1, 4.0, n - 7)]))[::-1]
        check(spectrum_matrix(n, lam.copy()), min(50, n), f"clustered n={n}")
        # rank-deficient: 20 non-zero eigenvalues, P beyond the rank (null-space vectors are arbitrary but
        # must stay orthonormal with zero residual)
        lam = np.concatenate([np.linspace(3.0, 1.0, 20), np.zeros(n - 20)])
        check(spectrum_matrix(n, lam), min(50, n), f"rank20 n={n}")
    # the covariance of clr rows is exactly singular (rows sum to zero): D = 512 with one zero eigenvalue
    Xr = rng.normal(size=(6000, 512))
    Xr -= Xr.mean(axis=1, keepdims=True)
    check(np.cov(Xr, rowvar=False), 50, "clr-like singular covariance")
    assert not D._topk_state.get("disabled"), "am_eigh_topk fell back to the Jacobi solver on this GPU"


def test_small_entry_points_against_numpy(torch_cuda):
    """The small C entries the sharded / streamed / large-data-mode passes are built from, each against numpy:
    am_table_stats, am_stats_pack, am_gram_recenter (+ am_cov_finish), am_axpb_f64, am_pca_finish."""
    import torch

    from autometa_b200 import _lib as L
    from autometa_b200 import device as D

    rng = np.random.default_rng(4)
    dev = torch.device("cuda", 0)
    st = L.stream_ptr(dev)
    # am_table_stats: presence and row sums of a row subset, D not a multiple of 32 and > 1024
    for Dc in (136, 512, 2080):
        cnt = rng.poisson(0.6, size=(300, Dc)).astype(np.int32)
        cnt[:, 5] = 0
        cnt[17] = 0
        rows = np.sort(rng.choice(300, size=120, replace=False)).astype(np.int64)
        d_cnt, d_rows = torch.from_numpy(cnt).to(dev), torch.from_numpy(rows).to(dev)
        present = torch.empty(Dc, dtype=torch.int32, device=dev)
        row_sum = torch.empty(len(rows), dtype=torch.int32, device=dev)
        L.check(L.lib.am_table_stats(L.t_ptr(d_cnt), L.t_ptr(d_rows), len(rows), Dc, L.t_ptr(present), L.t_ptr(row_sum), st))
        assert np.array_equal(present.cpu().numpy(), (cnt[rows].sum(axis=0) > 0).astype(np.int32))
        assert np.array_equal(row_sum.cpu().numpy(), cnt[rows].sum(axis=1))
        rs_full = torch.empty(300, dtype=torch.int32, device=dev)
        L.check(L.lib.am_table_stats(L.t_ptr(d_cnt), None, 300, Dc, L.t_ptr(present), L.t_ptr(rs_full), st))  # all rows
        assert np.array_equal(rs_full.cpu().numpy(), cnt.sum(axis=1))
        assert np.array_equal(present.cpu().numpy(), (cnt.sum(axis=0) > 0).astype(np.int32))
        # am_stats_pack: the SUM-reducible form
        out = torch.full((Dc + 1,), -1.0, dtype=torch.float64, device=dev)
        rs_all = torch.from_numpy(cnt.sum(axis=1).astype(np.int32)).to(dev)
        pr_all = torch.from_numpy((cnt.sum(axis=0) > 0).astype(np.int32)).to(dev)
        D.stats_pack(pr_all, rs_all, 300, out)
        got = out.cpu().numpy()
        assert np.array_equal(got[:Dc], (cnt.sum(axis=0) > 0).astype(np.float64))
        assert got[Dc] == float((cnt.sum(axis=1) == 0).sum())
    # am_gram_recenter + am_cov_finish: Gram about c_old -> Gram about c_new -> covariance
    n, Dg = 500, 40
    X = rng.normal(size=(n, Dg)) * rng.uniform(0.5, 3.0, size=Dg) + rng.normal(size=Dg) * 5
    c_old, c_new = X[:50].mean(axis=0), rng.normal(size=Dg)
    G = torch.from_numpy((X - c_old).T @ (X - c_old)).to(dev)
    cs = torch.from_numpy(X.sum(axis=0)).to(dev)
    n_t = torch.tensor([float(n)], dtype=torch.float64, device=dev)
    D.gram_recenter(G, cs, torch.from_numpy(c_old).to(dev), torch.from_numpy(c_new).to(dev), n_t)
    ref = (X - c_new).T @ (X - c_new)
    assert np.max(np.abs(G.cpu().numpy() - ref)) <= 1e-12 * np.max(np.abs(ref))
    D.gram_recenter(G, cs, torch.from_numpy(c_new).to(dev), None, n_t)  # ... about the origin
    assert np.max(np.abs(G.cpu().numpy() - X.T @ X)) <= 1e-12 * np.max(np.abs(X.T @ X))
    mu = D.cov_finish(G, cs, torch.zeros(Dg, dtype=torch.float64, device=dev), n_t)
    np.testing.assert_allclose(mu.cpu().numpy(), X.mean(axis=0), rtol=1e-14)
    C = np.cov(X, rowvar=False)
    assert np.max(np.abs(G.cpu().numpy() - C)) <= 1e-10 * np.max(np.abs(C))  # X^T X - n mu mu^T: sklearn's own form
    # am_axpb_f64, am_pca_finish
    v = torch.from_numpy(rng.normal(size=1000)).to(dev)
    v0 = v.cpu().numpy().copy()
    L.check(L.lib.am_axpb_f64(L.t_ptr(v), 1000, 0.25, -3.0, st))
    np.testing.assert_array_equal(v.cpu().numpy(), np.array([np.float64(x) * 0.25 - 3.0 for x in v0]))  # fma: exact here
    ev = torch.tensor([5.0, 2.0, -1e-18, 0.5], dtype=torch.float64, device=dev)
    tot = torch.tensor([10.0], dtype=torch.float64, device=dev)
    ratio = torch.empty(4, dtype=torch.float64, device=dev)
    L.check(L.lib.am_pca_finish(L.t_ptr(ev), L.t_ptr(tot), 4, L.t_ptr(ratio), st))
    assert ev.cpu().tolist() == [5.0, 2.0, 0.0, 0.5] and ratio.cpu().tolist() == [0.5, 0.2, 0.0, 0.05]
    # reserved SMs round trip
    L.check(L.lib.am_set_reserved_sms(16))
    assert L.lib.am_get_reserved_sms() == 16
    L.check(L.lib.am_set_reserved_sms(0))
    assert L.lib.am_set_reserved_sms(-1) != 0


# ---------------------------------------------------------------------------- DBSCAN
def _sweep_table():
    from autometa_b200 import synth

    return synth.make_sweep_input(n_points=3000, n_clusters=12, n_markers=40, seed=11)


def test_run_dbscan_matches_reference_golden(torch_cuda, golden_dbscan):
    from autometa_b200.binning import recursive_dbscan as R
    from autometa_b200.common.exceptions import TableFormatError

    table, _ = _sweep_table()
    for i, eps in enumerate(golden_dbscan["eps_list"]):
        df = table.copy()
        df["purity"] = 1.0  # must be dropped in place (recursive_dbscan.py:95)
        out = R.run_dbscan(df, eps=float(eps))
        assert "purity" not in df.columns and "purity" not in out.columns
        assert out["cluster"].dtype == np.int64
        assert np.array_equal(out["cluster"].to_numpy(), golden_dbscan[f"labels_{i}"]), eps
        assert out.index.equals(table.index)
    one = R.run_dbscan(table.iloc[:1].copy(), eps=0.3)
    assert one["cluster"].isna().all()
    bad = table.copy()
    bad.iloc[3, 0] = np.nan
    with pytest.raises(TableFormatError):
        R.run_dbscan(bad, eps=0.3)


def test_recursive_dbscan_matches_reference_golden(torch_cuda, golden_dbscan):
    from autometa_b200.binning import recursive_dbscan as R

    table, markers = _sweep_table()
    clustered, unclustered = R.recursive_dbscan(table.copy(), markers, 20.0, 90.0, 25.0, 5.0)
    assert clustered.columns.tolist() == golden_dbscan["clustered_columns"].tolist()
    assert clustered.index.tolist() == golden_dbscan["clustered_index"].tolist()
    assert np.array_equal(clustered["cluster"].to_numpy(), golden_dbscan["clustered_cluster"])
    for m in ["completeness", "purity", "coverage_stddev", "gc_content_stddev"]:
        np.testing.assert_allclose(clustered[m].to_numpy(dtype=float), golden_dbscan[f"clustered_{m}"],
                                   rtol=1e-12, atol=1e-12)
    assert unclustered.index.tolist() == golden_dbscan["unclustered_index"].tolist()
    assert np.array_equal(unclustered["cluster"].to_numpy(), golden_dbscan["unclustered_cluster"])


def _frames_equal(a, b):
    assert a.index.tolist() == b.index.tolist() and a.columns.tolist() == b.columns.tolist()
    for c in a.columns:
        x, y = a[c].to_numpy(), b[c].to_numpy()
        if x.dtype.kind == "f" or y.dtype.kind == "f":
            # the sigma columns are sums of fp64 atomics: the last ulp depends on the order the adds retire in
            assert np.allclose(x.astype(float), y.astype(float), rtol=1e-12, atol=0, equal_nan=True), c
        else:
            assert [None if pd.isna(v) else v for v in x] == [None if pd.isna(v) else v for v in y], c


def test_device_resident_sweep_equals_host_driven_loop(torch_cuda, monkeypatch):
    """The sweep with its loop state on the device (controller kernel + CUDA-graph replays, csrc/am_sweep.cu)
    against the same sweep with one read-back per eps and the schedule in Python: same best eps, same step
    count, identical labels, counts and completeness / purity (sigma to 1e-12); graph replays == plain launches; extra steps past the end change
    nothing."""
    import torch

    from autometa_b200 import _lib as L
    from autometa_b200 import device as D
    from autometa_b200 import synth
    from autometa_b200.binning import recursive_dbscan as R
    from autometa_b200.binning.utilities import markers_coo

    dev = torch.device("cuda", 0)
    for n, k, seed in ((3000, 12, 11), (700, 5, 4), (2, 1, 1), (60, 3, 9)):
        table, markers = synth.make_sweep_input(n_points=n, n_clusters=k, n_markers=40, seed=seed)
        cut = (20.0, 90.0, 25.0, 5.0)
        X = R._feature_matrix(table)
        cov, gc = table["coverage"].to_numpy(), table["gc_content"].to_numpy()
        rows, cols, vals = markers_coo(markers, table.index)
        want_labels, want_m, want_best = R._sweep_hostloop(X, cov, gc, rows, cols, vals, markers.shape[1], cut, dev, False)
        states = []
        for graph in (True, False):
            run = D.DeviceSweep(X, cov, gc, rows, cols, vals, markers.shape[1], cut, dev, trace_cap=512, graph=graph)
            batches = 0
            while True:
                run.launch(7)
                run.wait()
                batches += 1
                if run.poll():
                    break
            run.launch(5)  # past the end: frozen
            run.wait()
            labels, m, best, nc, trace = run.result()
            st = run.state()
            run.close()
            assert np.array_equal(labels, want_labels), (n, graph)
            assert best == want_best and nc == len(want_m["completeness"])
            for key in want_m:  # sigma: sums of fp64 atomics, last-ulp order dependence; the rest is exact
                if key.endswith("stddev"):
                    assert np.allclose(m[key], want_m[key], rtol=1e-12, atol=0, equal_nan=True), (n, graph, key)
                else:
                    assert np.array_equal(m[key], want_m[key], equal_nan=True), (n, graph, key)
            n_steps = int(st[L.SWEEP_NSTEPS])
            assert trace.shape == (n_steps, 4) and n_steps <= 7 * batches
            # the eps of every step = the reference's schedule replayed on the outcomes the steps produced
            eps, step, rounds = 0.3, 0.1, {0: 0}
            for t_eps, t_nc, t_med, _ in trace:
                assert t_eps == eps
                rounds[int(t_nc)] = rounds.get(int(t_nc), 0) + 1
                if rounds[int(t_nc)] > 10:
                    step *= 10
                if t_med == 0:
                    rounds[0] += 1
                eps += step
            assert trace[int(st[L.SWEEP_BEST_STEP]), 0] == st[L.SWEEP_BEST_EPS]
            assert trace[-1, 3] == best
            states.append((n_steps, st[L.SWEEP_BEST_EPS], st[L.SWEEP_EPS], st[L.SWEEP_STEP]))
        assert states[0] == states[1]
    # the API in both modes (frames as the reference builds them)
    table, markers = _sweep_table()
    a = R.recursive_dbscan(table.copy(), markers, 20.0, 90.0, 25.0, 5.0, verbose=True)
    monkeypatch.setenv("AUTOMETA_B200_SWEEP", "host")
    b = R.recursive_dbscan(table.copy(), markers, 20.0, 90.0, 25.0, 5.0)
    monkeypatch.delenv("AUTOMETA_B200_SWEEP")
    _frames_equal(a[0], b[0])
    _frames_equal(a[1], b[1])
    launches = L.launch_count()
    R.recursive_dbscan(table.copy(), markers, 20.0, 90.0, 25.0, 5.0)
    assert L.launch_count() - launches >= 16 * 15  # replays are counted as the kernels they launch


def test_get_clusters_many_equals_one_at_a_time(torch_cuda):
    """Sweeps in flight together (own streams, own scratch) give the frames of one-at-a-time calls."""
    from autometa_b200 import synth
    from autometa_b200.binning import recursive_dbscan as R

    tables = []
    markers = None
    for i, (n, k) in enumerate(((1500, 9), (400, 4), (1, 1), (2500, 15), (90, 3), (800, 6), (1200, 7))):
        t, m = synth.make_sweep_input(n_points=n, n_clusters=k, n_markers=40, seed=21 + i)
        t.index = pd.Index([f"T{i}_{c}" for c in t.index], name="contig")
        m.index = pd.Index([f"T{i}_{c}" for c in m.index], name="contig")
        tables.append(t)
        markers = m if markers is None else pd.concat([markers, m])
    one = [R.get_clusters(t.copy(), markers, 20.0, 90.0, 25.0, 5.0, method="dbscan") for t in tables]
    many = R.get_clusters_many([t.copy() for t in tables], markers, 20.0, 90.0, 25.0, 5.0, method="dbscan")
    assert len(many) == len(one)
    for a, b in zip(one, many):
        _frames_equal(a, b)
    old = R._IN_FLIGHT_MAX
    try:
        R._IN_FLIGHT_MAX = 2  # slots (streams) reused by later jobs
        few = R.get_clusters_many([t.copy() for t in tables], markers, 20.0, 90.0, 25.0, 5.0, method="dbscan")
    finally:
        R._IN_FLIGHT_MAX = old
    for a, b in zip(one, few):
        _frames_equal(a, b)
    assert R.get_clusters_many([], markers, 20.0, 90.0, 25.0, 5.0, method="dbscan") == []


def test_sweep_incremental_equals_fresh_and_oracle(torch_cuda):
    import torch

    from autometa_b200 import device as D, synth

    table, _ = synth.make_sweep_input(20000, 60, 10, seed=5)
    X = np.ascontiguousarray(table[["x_1", "x_2", "coverage"]].to_numpy())
    Xd = torch.from_numpy(X).cuda()
    sweep = D.EpsSweep(Xd)
    prev_n = None
    for eps in O.eps_schedule(48)[::6] + [14.6, 40.0, 400.0]:
        inc = sweep.step(eps).cpu().numpy()
        fresh = D.EpsSweep(Xd).step(eps).cpu().numpy()
        assert np.array_equal(inc, fresh)
        n = int(sweep.n_clusters.item())
        assert n == inc.max() + 1
        # labels are numbered by first occurrence (sklearn's visiting order)
        first = np.unique(inc, return_index=True)[1]
        assert np.all(np.diff(first) > 0)
        if prev_n is not None:
            assert n <= prev_n
        prev_n = n
        if eps <= 3.5:
            assert np.array_equal(inc, O.dbscan_labels(X, eps)), eps
    assert prev_n == 1


@pytest.mark.parametrize("F", [1, 2, 4])
def test_eps_step_other_feature_counts(torch_cuda, F):
    """The neighbour kernel is instantiated per feature count (3^F cells per point: 3 / 9 / 81, the last in three
    rounds of 32 lanes); F = 3 is what every other test runs.  Labels against the oracle and against sklearn,
    through EpsSweep and through run_dbscan (x_1 .. x_{F-1} + coverage)."""
    import torch
    from sklearn.cluster import DBSCAN

    from autometa_b200 import device as D
    from autometa_b200.binning import recursive_dbscan as R

    rng = np.random.default_rng(40 + F)
    n, k = 4000, 25
    centres = rng.uniform(-30, 30, size=(k, F))
    X = np.ascontiguousarray(centres[rng.integers(0, k, size=n)] + rng.normal(0, 0.8, size=(n, F)))
    X[:50] = X[50:100]  # exact duplicates
    sweep = D.EpsSweep(torch.from_numpy(X).cuda())
    for eps in (0.05, 0.3, 0.7999999999999999, 1.5, 4.0, 200.0):
        lab = sweep.step(eps).cpu().numpy()
        assert np.array_equal(lab, O.dbscan_labels(X, eps)), (F, eps)
        ref = DBSCAN(eps=eps, min_samples=1).fit(X).labels_
        assert np.array_equal(lab, ref.astype(np.int64)), (F, eps)
    cols = [f"x_{i + 1}" for i in range(F - 1)] + ["coverage"]
    df = pd.DataFrame(X, columns=cols, index=pd.Index([f"c{i}" for i in range(n)], name="contig"))
    out = R.run_dbscan(df.copy(), eps=0.7)
    assert np.array_equal(out["cluster"].to_numpy(), DBSCAN(eps=0.7, min_samples=1).fit(X).labels_)


def test_sweep_full_size_c4(torch_cuda):
    """C4: 200k points, the reference's 48-value eps grid; the label vector is identical to sklearn's
    (what recursive_dbscan.py:109 executes) at EVERY eps of the grid."""
    import torch
    from sklearn.cluster import DBSCAN

    from autometa_b200 import device as D, synth

    table, _ = synth.make_sweep_input()
    X = np.ascontiguousarray(table[["x_1", "x_2", "coverage"]].to_numpy())
    sweep = D.EpsSweep(torch.from_numpy(X).cuda())
    grid = O.eps_schedule(48)
    prev = None
    for i, eps in enumerate(grid):
        lab = sweep.step(eps).cpu().numpy()
        n = int(sweep.n_clusters.item())
        if prev is not None:
            assert n <= prev
        prev = n
        ref = DBSCAN(eps=eps, min_samples=1, n_jobs=-1).fit(X).labels_
        assert np.array_equal(lab, ref.astype(np.int64)), (i, eps)
    assert i == 47


def test_device_cluster_metrics_match_host_mirror(torch_cuda):
    """am_cluster_metrics (device add_metrics + filter + median) vs the numpy mirror of
    binning/utilities.py:125-262 at several eps of a synthetic sweep."""
    import torch

    from autometa_b200 import device as D, synth
    from autometa_b200.binning.utilities import cluster_metrics_arrays, markers_coo, passing_mask

    table, markers = synth.make_sweep_input(n_points=20000, n_clusters=40, n_markers=60, seed=9)
    X = np.ascontiguousarray(table[["x_1", "x_2", "coverage"]].to_numpy())
    cov = table["coverage"].to_numpy(dtype=np.float64)
    gc = table["gc_content"].to_numpy(dtype=np.float64)
    rows, cols, vals = markers_coo(markers, table.index)
    cut = (20.0, 90.0, 25.0, 5.0)
    sweep = D.EpsSweep(torch.from_numpy(X).cuda())
    met = D.ClusterMetrics(len(table), rows, cols, vals, markers.shape[1], cov, gc, cut, sweep.dev)
    for eps in (0.3, 0.8999999999999999, 2.0, 3.5000000000000004, 6.0, 30.0):
        lab_d = sweep.step(eps)
        summ = met.compute(lab_d, sweep.n_clusters).cpu().numpy()
        lab = lab_d.cpu().numpy()
        k = int(sweep.n_clusters.item())
        ref = cluster_metrics_arrays(lab, k, rows, cols, vals, markers.shape[1], cov, gc)
        got = met.arrays(k)
        assert int(summ[0]) == k
        assert np.array_equal(got["present_marker_count"], ref["present_marker_count"])
        assert np.array_equal(got["single_copy_marker_count"], ref["single_copy_marker_count"])
        for name in ("completeness", "purity", "coverage_stddev", "gc_content_stddev"):
            np.testing.assert_allclose(got[name], ref[name], rtol=1e-11, atol=1e-11, equal_nan=True)
        ok = passing_mask(ref, *cut)
        med = float(np.median(ref["completeness"][ok])) if ok.any() else float("-inf")
        assert summ[1] == med or (np.isinf(med) and np.isinf(summ[1])), (eps, summ[1], med)
        assert int(summ[2]) == int(ok.sum())


@pytest.mark.parametrize("method", ["clr", "ilr"])
def test_featurise_clr_ilr_vs_oracle_and_sklearn(torch_cuda, method):
    """Whole device pipeline for the skbio-style transforms: clr takes the fused normalise + int8
    digit-plane path (tcgen05 Gram and projection), ilr the generic path (511 columns)."""
    from sklearn.decomposition import PCA

    from autometa_b200 import assembly as A, pipeline, synth

    asm = synth.make_assembly(5600, 5600 * 3400, seed=123)
    packed = A.pack_to_device(asm, device="cuda:0", on_device=True)
    res = pipeline.featurise(packed, 5, method, 50)
    ref_counts, _ = O.count_sequences(synth.sequences(asm), 5)
    refX = (O.clr if method == "clr" else O.ilr)(ref_counts.astype(np.float64))
    X = res["normalized"].cpu().numpy()
    assert X.shape == refX.shape
    assert np.max(np.abs(X - refX) / np.max(np.abs(refX), axis=1, keepdims=True)) <= 1e-12
    p = PCA(n_components=50, random_state=np.random.RandomState(42))
    sc = p.fit_transform(refX)
    assert p._fit_svd_solver == "covariance_eigh"
    ref = dict(explained_variance=p.explained_variance_, explained_variance_ratio=p.explained_variance_ratio_,
               components=p.components_, scores=sc)
    got = {k: res[k].cpu().numpy() for k in ref}
    _check_pca(got, ref)


def test_digit_planes_fused_equals_generic_quantiser(torch_cuda):
    """The planes written by the fused normalise kernel are byte-identical to quantising its fp64
    output with the generic kernel (same centre and scale), and the tcgen05 Gram of the planes
    agrees with the fp64 CUDA-core Gram."""
    import torch

    from autometa_b200 import device as D

    rng = np.random.default_rng(11)
    counts = rng.poisson(6.0, size=(4100, 512)).astype(np.int32)
    counts[3, :17] = rng.integers(3000, 200000, size=17)
    d = torch.from_numpy(counts).cuda()
    center = torch.from_numpy(rng.normal(0, 0.3, size=512)).cuda()
    e, sc = D.plane_scales(center, None, float(np.log1p(200000.0)))
    cs = torch.zeros(512, dtype=torch.float64, device="cuda")
    X, planes = D.normalize_planes(d, "am_clr", None, center, sc, cs)
    planes2 = D.quantize_planes(X, center, sc)
    n = X.shape[0]
    assert torch.equal(planes[: 6 * n * 512], planes2[: 6 * n * 512])
    np.testing.assert_allclose(cs.cpu().numpy(), X.sum(dim=0).cpu().numpy(), rtol=0, atol=1e-9)
    G8 = D.gram_planes(planes, n, 512, e).cpu().numpy()
    G64 = D.gram_centered(X, center).cpu().numpy()
    scale = np.sqrt(np.outer(np.diag(G64), np.diag(G64)))
    assert np.max(np.abs(G8 - G64) / scale) <= 5e-12
    assert np.array_equal(G8, G8.T)


# ----------------------------------------------------------------------------- FASTA-ingest by-products (8f-2)
def test_gc_content_matches_oracle(torch_cuda, small_fna, tmp_path):
    from autometa_b200.common.metagenome import Metagenome

    ids, seqs = O.read_fasta(small_fna)
    gc, length = O.gc_content_table(seqs)
    mg = Metagenome(small_fna)
    df = mg.gc_content()
    assert df.index.name == "contig" and list(df.columns) == ["gc_content", "length"]
    assert df.index.tolist() == ids
    assert np.array_equal(df["length"].to_numpy(), length)
    assert np.array_equal(df["gc_content"].to_numpy(), gc)  # bit-exact: integer counts, one division
    assert mg.length_weighted_gc == O.length_weighted_gc(seqs)
    d = mg.describe()
    assert d.index.name == "assembly" and d["nseqs"].iloc[0] == len(seqs)
    assert d["length_weighted_gc_content (%)"].iloc[0] == O.length_weighted_gc(seqs)
    # ragged lengths around the 64-base group size, every byte class, empty and header-only records
    rng = np.random.default_rng(5)
    alphabet = np.frombuffer(b"ACGTacgtNnSsWwUuRYKMBDHV-*.", dtype=np.uint8)
    recs = []
    for L in [0, 1, 3, 4, 5, 63, 64, 65, 127, 128, 129, 255, 4096, 70001] + list(rng.integers(1, 3000, 200)):
        p = np.full(len(alphabet), 0.2 / (len(alphabet) - 8))
        p[:8] = 0.1
        recs.append(alphabet[rng.choice(len(alphabet), size=int(L), p=p / p.sum())].tobytes())
    fa = tmp_path / "r.fa"
    with open(fa, "wb") as fh:
        for i, s in enumerate(recs):
            fh.write(b">r%d d\n" % i + b"\n".join(s[j : j + 70] for j in range(0, len(s), 70)) + b"\n")
    _, seqs2 = O.read_fasta(str(fa))
    assert seqs2 == recs
    gc2, len2 = O.gc_content_table(recs)
    df2 = Metagenome(str(fa)).gc_content()
    assert np.array_equal(df2["length"].to_numpy(), len2)
    assert np.array_equal(df2["gc_content"].to_numpy(), gc2)
    # the packed codes of the GC pass are the ones the plain packer writes
    from autometa_b200 import assembly as A

    h = A.read_fasta(str(fa))
    a, b = A.pack_to_device(h, on_device=True), A.pack_to_device(h, gc=True)
    assert torch_cuda.equal(a.codes, b.codes) and torch_cuda.equal(a.exc_bitmap, b.exc_bitmap)
    assert a.gc_content is None and b.gc_content.shape[0] == len(recs)


# ----------------------------------------------------------------------------- callers of the sweep (8f-3)
def _check_binning_frame(df, g, prefix):
    assert df.index.tolist() == g[f"{prefix}_index"].tolist()
    assert df.columns.tolist() == g[f"{prefix}_columns"].tolist()
    got = np.array(["" if pd.isna(c) else str(c) for c in df["cluster"]])
    assert np.array_equal(got, g[f"{prefix}_cluster"])
    for m in ["completeness", "purity", "coverage_stddev", "gc_content_stddev"]:
        a = pd.to_numeric(df[m], errors="coerce").to_numpy(dtype=np.float64, na_value=np.nan)
        assert np.allclose(a, g[f"{prefix}_{m}"], rtol=1e-12, atol=0, equal_nan=True), m


def test_get_clusters_matches_reference_golden(torch_cuda, golden_binning):
    from autometa_b200 import synth
    from autometa_b200.binning import recursive_dbscan as R

    table, markers = synth.make_sweep_input(n_points=1500, n_clusters=9, n_markers=40, seed=21)
    df = R.get_clusters(table.copy(), markers, 20.0, 90.0, 25.0, 5.0, method="dbscan")
    _check_binning_frame(df, golden_binning, "get_clusters")
    df0 = R.get_clusters(table.iloc[:40].copy(), markers, 99.0, 99.0, 0.001, 0.001, method="dbscan")
    _check_binning_frame(df0, golden_binning, "get_clusters_none")
    with pytest.raises(ValueError):
        R.get_clusters(table.copy(), markers, 20.0, 90.0, 25.0, 5.0, method="kmeans")


def test_taxon_guided_binning_matches_reference_golden(torch_cuda, golden_binning):
    from autometa_b200 import synth
    from autometa_b200.binning import recursive_dbscan as R
    from autometa_b200.common.exceptions import BinningError, TableFormatError

    table, markers = synth.make_sweep_input(n_points=1500, n_clusters=9, n_markers=40, seed=21)
    tax = synth.add_taxonomy(table, seed=3)
    df = R.taxon_guided_binning(tax.copy(), markers, 20.0, 90.0, 25.0, 5.0, method="dbscan")
    _check_binning_frame(df, golden_binning, "taxon")
    dfr = R.taxon_guided_binning(tax.copy(), markers, 20.0, 90.0, 25.0, 5.0, starting_rank="phylum",
                                 method="dbscan", reverse_ranks=True)
    _check_binning_frame(dfr, golden_binning, "taxon_rev")
    with pytest.raises(TableFormatError):
        R.taxon_guided_binning(tax.copy(), markers.iloc[:0], method="dbscan")
    with pytest.raises(BinningError):
        R.taxon_guided_binning(tax.loc[[markers.index[0]]].copy(), markers, method="dbscan")


# ----------------------------------------------------------------------------- large-data mode (8f-3 / 8f-4)
LDM_KW = dict(norm_method="am_clr", pca_dimensions=10, embed_dimensions=2, embed_method="bhsne", completeness=20.0,
              purity=90.0, coverage_stddev=25.0, gc_content_stddev=5.0, starting_rank="superkingdom", method="dbscan")


@pytest.fixture()
def stub_bhsne(monkeypatch):
    """The golden runs of the reference used a ``tsne.bh_sne`` that returns the first d columns it is handed
    (the PCA scores): the same stub here, so that the frames are comparable end to end."""
    import sys
    import types

    stub = types.ModuleType("tsne")
    stub.bh_sne = lambda data, d, **kw: np.ascontiguousarray(data[:, :d])
    monkeypatch.setitem(sys.modules, "tsne", stub)


def test_partition_pca_matches_sequential_path(torch_cuda):
    """PartitionPCA (all partitions of a rank queued together over CUDA streams) == kmers.normalize + kmers.pca
    partition by partition, for the fused k = 5 path with row gathers (incl. a partition with an all-NA contig
    and one with an absent k-mer column -> the general path) and for a k = 4 table."""
    from autometa_b200 import assembly as A, synth
    from autometa_b200.binning.large_data_mode import PartitionPCA
    from autometa_b200.common import kmers

    asm = synth.make_assembly(2600, 2600 * 3300, seed=41, tiny=True)
    from oracle import oracle_c

    for k in (5, 4):
        cnt = oracle_c.count(asm.seq, asm.offsets, k, 1).astype(np.float64)
        if k == 5:
            cnt[100:160, 7] = 0  # a k-mer absent from one partition only
        cnt[cnt == 0] = np.nan
        names, _ = O.kmer_columns(k)
        counts = pd.DataFrame(cnt, index=pd.Index(asm.ids, name="contig"), columns=names)
        rng = np.random.default_rng(5)
        perm = rng.permutation(2600)
        parts = {"a": counts.index[np.sort(perm[:900])], "b": counts.index[np.sort(perm[900:1500])],
                 "gap": counts.index[100:160], "tail": counts.index[2400:]}  # "tail" holds the length-4/5/6 contigs
        for method in ("am_clr", "ilr"):
            use = {n: i for n, i in parts.items() if not (method == "ilr" and n == "tail")}
            got = PartitionPCA(counts, method, 20).run(use)
            for name, idx in use.items():
                norm = kmers.normalize(counts.loc[idx], method=method)
                ref = kmers.pca(norm.to_numpy(), n_components=20)["scores"]
                assert got[name].index.tolist() == norm.index.tolist(), (k, method, name)
                assert np.max(np.abs(got[name].to_numpy() - ref)) <= 1e-8 * np.max(np.abs(ref)), (k, method, name)


def test_large_data_mode_matches_reference_golden(torch_cuda, stub_bhsne, tmp_path):
    """cluster_by_taxon_partitioning on the seeded inputs of tests/_ldm_input.py against the frames the
    unmodified reference produced (tests/golden/make_ldm_golden.py): bins, metrics, column order; with a cache
    directory also the embedding-cache file names, the checkpoint columns and the checkpoint header; and a
    second call on the same cache resumes from the checkpoint instead of starting over."""
    import _ldm_input

    from autometa_b200.binning import large_data_mode as M

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ldm_golden.npz"))
    main, counts, markers = _ldm_input.make_ldm_input()
    df = M.cluster_by_taxon_partitioning(main.copy(), counts.copy(), markers, max_partition_size=10000, **LDM_KW)
    _check_binning_frame(df, g, "ldm")
    cache = str(tmp_path / "cache")
    df2 = M.cluster_by_taxon_partitioning(main.copy(), counts.copy(), markers, max_partition_size=400, cache=cache,
                                          **LDM_KW)
    _check_binning_frame(df2, g, "ldm400")
    files = sorted(os.path.relpath(os.path.join(d, f), cache) for d, _, fs in os.walk(cache) for f in fs)
    assert files == g["ldm400_cache_files"].tolist()
    info = M.get_checkpoint_info(os.path.join(cache, "binning_checkpoints.tsv.gz"))
    assert info["binning_checkpoints"].columns.tolist() == g["ldm400_checkpoint_columns"].tolist()
    assert [info["starting_rank"], info["starting_rank_name_txt"]] == g["ldm400_checkpoint_rank"].tolist()
    # every cached embedding is a gzip TSV that pandas reads back as the frame we read back
    for f in files:
        if f.endswith(".tsv.gz") and "binning_checkpoints" not in f:
            a = pd.read_csv(os.path.join(cache, f), sep="\t", index_col="contig")
            b = M._tsv.read_table(os.path.join(cache, f), index_col="contig")
            pd.testing.assert_frame_equal(a, b, check_exact=True)
    # resume: the checkpointed bins are kept (the same contig groups; zero_pad_bin_names renumbers bins in order
    # of first appearance, utilities.py:328-363, so the names of a resumed run may be a permutation)
    df3 = M.cluster_by_taxon_partitioning(main.copy(), counts.copy(), markers, max_partition_size=400, cache=cache,
                                          **LDM_KW)
    binned = df2["cluster"].notna()
    before, after = df2.loc[binned, "cluster"].tolist(), df3.loc[df2.index[binned], "cluster"].tolist()
    assert not pd.isna(after).any()
    assert len(set(zip(before, after))) == len(set(before)) == len(set(after))
    with pytest.raises(FileNotFoundError):
        M.cluster_by_taxon_partitioning(main.copy(), counts.copy(), markers, binning_checkpoints_fpath=str(tmp_path / "no.tsv"),
                                        **LDM_KW)


# ----------------------------------------------------------------------------- full-size featurise properties
@pytest.mark.parametrize("case", ["C2_am_clr", "C5_shard_ilr"])
def test_featurise_full_size_properties(torch_cuda, case):
    """BASELINE.json sizes through size-independent properties: C2 (200k contigs / 1 Gbp, am_clr, fused
    tensor-core path) and one GPU's share of C5 (125k contigs / 625 Mbp, ilr, generic path).  Checked:
    the normalised rows of a contig sample against the oracle, ilr's isometry with clr, orthonormal
    components with sklearn's sign rule, column variances of the scores == explained_variance_ (the
    defining property of PCA scores), explained_variance_ratio_ == EV / trace, and the scores of a row
    sample recomputed in float64 from the returned mean-free rows and components."""
    import torch

    from autometa_b200 import assembly as A, pipeline, synth

    n, bp, method = (200_000, 1_000_000_000, "am_clr") if case == "C2_am_clr" else (125_000, 625_000_000, "ilr")
    asm = synth.make_assembly(n, bp, seed=20261019 + (2 if method == "am_clr" else 5))
    packed = A.pack_to_device(asm, device="cuda:0", on_device=True)
    res = pipeline.featurise(packed, 5, method, 50)
    X, S = res["normalized"], res["scores"]
    D = 512 if method == "am_clr" else 511
    assert tuple(X.shape) == (n, D) and tuple(S.shape) == (n, 50)
    sample = np.random.default_rng(1).choice(n, size=200, replace=False)
    seqs = [asm.seq[int(asm.offsets[i]): int(asm.offsets[i + 1])].tobytes() for i in sample]
    ref_counts, _ = O.count_sequences(seqs, 5)
    assert np.array_equal(res["counts"][torch.from_numpy(sample).cuda()].cpu().numpy(), ref_counts)
    Xs = X[torch.from_numpy(sample).cuda()].cpu().numpy()
    if method == "am_clr":
        refX = O.am_clr(ref_counts.astype(np.float64))[0]  # no column is absent at this size
    else:
        refX = O.ilr(ref_counts.astype(np.float64))
        clr_rows = O.clr(ref_counts.astype(np.float64))
        assert np.allclose((Xs ** 2).sum(1), (clr_rows ** 2).sum(1), rtol=1e-11, atol=0)  # ilr is an isometry
    assert np.max(np.abs(Xs - refX) / np.max(np.abs(refX), axis=1, keepdims=True)) <= 1e-12
    V = res["components"].cpu().numpy()
    ev, evr = res["explained_variance"].cpu().numpy(), res["explained_variance_ratio"].cpu().numpy()
    assert np.max(np.abs(V @ V.T - np.eye(50))) < 1e-11
    j = np.argmax(np.abs(V), axis=1)
    assert np.all(V[np.arange(50), j] > 0)  # svd_flip(u_based_decision=False)
    assert np.all(np.diff(ev) <= 0) and ev[-1] > 0
    mean = X.mean(dim=0, dtype=torch.float64)
    var_scores = ((S - S.mean(dim=0)) ** 2).sum(dim=0) / (n - 1)
    assert np.allclose(var_scores.cpu().numpy(), ev, rtol=1e-9, atol=0)
    assert float(S.mean(dim=0).abs().max()) < 1e-9 * float(S.abs().max())
    trace = float((((X - mean) ** 2).sum(dim=0) / (n - 1)).sum())
    assert np.allclose(evr, ev / trace, rtol=1e-9, atol=0)
    s_ref = (Xs - mean.cpu().numpy()) @ V.T
    s_got = S[torch.from_numpy(sample).cuda()].cpu().numpy()
    assert np.max(np.abs(s_got - s_ref)) <= 1e-8 * float(S.abs().max())
    if case == "C2_am_clr":
        # directly against sklearn on the full 200k x 512 matrix (what kmers.py:605 executes)
        from sklearn.decomposition import PCA

        Xh = X.cpu().numpy()
        p = PCA(n_components=50, random_state=np.random.RandomState(42))
        sc = p.fit_transform(Xh)
        assert p._fit_svd_solver == "covariance_eigh"
        _check_pca(dict(explained_variance=ev, explained_variance_ratio=evr, components=V, scores=S.cpu().numpy()),
                   dict(explained_variance=p.explained_variance_, explained_variance_ratio=p.explained_variance_ratio_,
                        components=p.components_, scores=sc))
        del Xh, sc
    # C V^T = V^T diag(ev): the components are eigenvectors of the sample covariance
    Xc = X - mean
    CV = (Xc.T @ (Xc @ res["components"].T)) / (n - 1)
    resid = (CV - res["components"].T * res["explained_variance"]).abs().max()
    assert float(resid) <= 1e-9 * float(ev[0])


def test_feature_stream_equals_serial_featurise(torch_cuda):
    """Throughput mode (two passes in flight on two streams) returns exactly what the serial pass returns,
    for every submitted assembly, including one whose optimistic am_clr assumption fails (tiny contigs ->
    all-NA rows -> that pass is redone on the general path)."""
    from autometa_b200 import assembly as A, pipeline, synth

    asms = [synth.make_assembly(5200, 5200 * 3400, seed=31), synth.make_assembly(5300, 5300 * 3300, seed=32, tiny=True),
            synth.make_assembly(5400, 5400 * 3200, seed=33), synth.make_assembly(5200, 5200 * 3400, seed=31)]
    packed = [A.pack_to_device(a, device="cuda:0", on_device=True) for a in asms]
    fs = pipeline.FeatureStream("cuda:0", 5, "am_clr", 50, depth=2, keep=True)
    for p in packed:
        fs.submit(p)
    got = fs.drain()
    assert len(got) == len(packed)
    for p, g in zip(packed, got):
        ref = pipeline.featurise(p, 5, "am_clr", 50)
        for key in ("counts", "normalized", "scores", "explained_variance", "explained_variance_ratio", "components",
                    "mean"):
            assert torch_cuda.equal(g[key], ref[key]), key


def test_host_featuriser_paths_agree(torch_cuda):
    """End-to-end entries from pinned host buffers: packed and ASCII host formats, one call at a time (run),
    double-buffered (run_many) and through FeatureStream (run_stream) all return the scores of the
    device-resident pass."""
    from autometa_b200 import assembly as A, pipeline, synth

    asm = synth.make_assembly(5600, 5600 * 3300, seed=71)
    ref = pipeline.featurise(A.pack_to_device(asm, device="cuda:0", on_device=True), 5, "am_clr", 50)
    ref_s, ref_ev = ref["scores"].cpu().numpy(), ref["explained_variance"].cpu().numpy()
    for fmt in ("packed", "ascii"):
        hf = pipeline.HostFeaturiser(asm, device="cuda:0", host_format=fmt)
        for call in (lambda: hf.run(), lambda: hf.run_many(3), lambda: hf.run_stream(5)):
            s, ev = call()
            assert np.array_equal(s.numpy(), ref_s) and np.array_equal(ev.numpy(), ref_ev), fmt
        assert hf.h2d_bytes > 0 and hf.d2h_bytes == ref_s.nbytes + ref_ev.nbytes


# ----------------------------------------------------------------------------- embed() through its own call site
def test_embed_pca_stage_matches_reference_embed_call(torch_cuda, golden_pca, tmp_path, monkeypatch):
    """The reference's embed() was run (tests/golden/make_golden.py) with a stub ``tsne.bh_sne`` that returns
    the first d columns it is handed, i.e. the PCA scores.  The same stub here: embed() must return the
    reference's frame (PCA stage on the GPU), write/read it back unchanged and keep the error behaviour."""
    import sys
    import types

    from autometa_b200.common import kmers
    from autometa_b200.common.exceptions import TableFormatError

    stub = types.ModuleType("tsne")
    stub.bh_sne = lambda data, d, **kw: np.ascontiguousarray(data[:, :d])
    monkeypatch.setitem(sys.modules, "tsne", stub)
    X = golden_pca["X"]
    df = pd.DataFrame(X, index=pd.Index([f"c{i}" for i in range(X.shape[0])], name="contig"))
    out = tmp_path / "kmers.embedded.tsv"
    emb = kmers.embed(df, out=str(out), method="bhsne", embed_dimensions=3, pca_dimensions=10, seed=42)
    assert list(emb.columns) == ["x_1", "x_2", "x_3"] and emb.index.equals(df.index)
    ref = golden_pca["embed_first3"]
    assert np.max(np.abs(emb.to_numpy() - ref)) <= 1e-8 * np.max(np.abs(ref))
    # cached file is returned as is; the TSV round trip is pandas-identical
    again = kmers.embed(df, out=str(out), method="bhsne", embed_dimensions=3, pca_dimensions=10)
    pd.testing.assert_frame_equal(again, pd.read_csv(out, sep="\t", index_col="contig"), check_exact=True)
    # pca_dimensions = 0 or >= the number of columns: no PCA, the manifold step sees the input
    raw = kmers.embed(df, method="bhsne", embed_dimensions=2, pca_dimensions=0)
    assert np.array_equal(raw.to_numpy(), X[:, :2])
    # sklearn's TSNE is installed: the sksne branch runs end to end on the GPU scores
    small = kmers.embed(df.iloc[:120], method="sksne", embed_dimensions=2, pca_dimensions=10, seed=1)
    assert small.shape == (120, 2) and np.isfinite(small.to_numpy()).all()
    (tmp_path / "bad.tsv").write_text("id\tx\nq\t1.0\n")
    with pytest.raises(TableFormatError):
        kmers.embed(str(tmp_path / "bad.tsv"), method="bhsne")
