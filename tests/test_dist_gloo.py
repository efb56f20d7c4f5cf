"""world_size-2 gloo tests of the N>1 host logic: bp sharding + the three exchange
steps (column presence OR, column sums / row count, Gram) reproduce the single-rank
statistics, and ragged row shards gather back in rank order."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    from autometa_b200 import dist as DD
    from autometa_b200 import synth
    from oracle import oracle_np as O

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        asm = synth.make_assembly(60, 60 * 3300, seed=3, tiny=True)
        seqs = synth.sequences(asm)
        lo, hi = DD.shard_ranges(asm.lengths, world)[rank]
        counts, ok = O.count_sequences(seqs[lo:hi], 4)
        present = torch.from_numpy((counts.sum(axis=0) > 0).astype(np.int32))
        DD.allreduce_max_(present)
        X = np.log1p(counts[counts.sum(axis=1) > 0].astype(np.float64))
        X = X - X.mean(axis=1, keepdims=True)
        cs = torch.from_numpy(X.sum(axis=0))
        n = torch.tensor([float(X.shape[0])], dtype=torch.float64)
        DD.allreduce_sum_([cs, n])
        mu = (cs / n).numpy()
        G = torch.from_numpy((X - mu).T @ (X - mu))
        DD.allreduce_sum_([G])
        rows = DD.allgather_rows(torch.from_numpy(X))
        if rank == 0:
            q.put(dict(present=present.numpy(), mu=mu, G=G.numpy(), n=float(n.item()), rows=rows.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_statistics_match_single_rank():
    import torch.multiprocessing as mp

    from autometa_b200 import synth
    from oracle import oracle_np as O

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    asm = synth.make_assembly(60, 60 * 3300, seed=3, tiny=True)
    counts, ok = O.count_sequences(synth.sequences(asm), 4)
    assert np.array_equal(got["present"], (counts.sum(axis=0) > 0).astype(np.int32))
    X = np.log1p(counts[counts.sum(axis=1) > 0].astype(np.float64))
    X = X - X.mean(axis=1, keepdims=True)
    assert got["n"] == X.shape[0]
    np.testing.assert_allclose(got["mu"], X.mean(axis=0), rtol=0, atol=1e-13)
    mu = X.mean(axis=0)
    np.testing.assert_allclose(got["G"], (X - mu).T @ (X - mu), rtol=1e-12, atol=1e-10)
    assert np.array_equal(got["rows"], X)


# ----------------------------------------------------------------------------- pipeline.featurise host logic
# The device calls are replaced by tests/_cpu_device_standin.py (oracle arithmetic on CPU tensors); what is
# under test is the host logic of the sharded pass: ONE all-reduce on the k = 5 path, and the same collective
# sequence on every rank whatever its shard holds (an all-NA row or an empty shard on one
# rank must not make that rank post collectives the others never post).
def _featurise_worker(rank, world, port, q, case, method):
    import torch
    import torch.distributed as dist

    import _cpu_device_standin as S
    from autometa_b200 import dist as DD
    from autometa_b200 import pipeline, synth

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    calls = []
    real = dist.all_reduce

    def counting_all_reduce(t, *a, **kw):
        calls.append(int(t.numel()))
        return real(t, *a, **kw)

    dist.all_reduce = counting_all_reduce
    pipeline._dev = S
    try:
        asm = synth.make_assembly(40, 40 * 3300, seed=11, tiny=(case == "tiny"))
        if case == "empty":
            lo, hi = (0, asm.n_contigs) if rank == 0 else (asm.n_contigs, asm.n_contigs)
        else:
            lo, hi = DD.shard_ranges(asm.lengths, world)[rank]
        err = None
        try:
            res = pipeline.featurise(S.fake_packed(asm.slice(lo, hi)), 5, method, 10, distributed=True)
            out = dict(ev=res["explained_variance"].numpy(), comps=res["components"].numpy(),
                       scores=res["scores"].numpy(), n_norm=int(res["normalized"].shape[0]))
        except ValueError as exc:
            err, out = str(exc), {}
        q.put((rank, calls, err, out))
    finally:
        dist.destroy_process_group()


def _run_featurise_case(case, method):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_featurise_worker, args=(r, 2, port, q, case, method)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=240) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return got


def _single_rank_reference(case, method, P=10):
    import _cpu_device_standin as S
    from autometa_b200 import synth
    from oracle import oracle_c
    from oracle import oracle_np as O

    asm = synth.make_assembly(40, 40 * 3300, seed=11, tiny=(case == "tiny"))
    counts = oracle_c.count(asm.seq, asm.offsets, 5, 1).astype(np.int64)
    counts = counts[counts.sum(axis=1) > 0]
    X = np.log(counts + 1.0) - np.log(counts + 1.0).mean(axis=1, keepdims=True) if method == "am_clr" else (
        O.clr(counts.astype(float)) if method == "clr" else O.ilr(counts.astype(float)))
    return O.pca_cov_eigh(X, P), X.shape[0]


@pytest.mark.timeout(300)
@pytest.mark.parametrize("case", ["plain", "tiny", "empty"])
def test_featurise_collectives_identical_on_every_rank(case):
    """plain: one all-reduce; tiny: rank 1's shard ends with contigs of length 4, 5, 6 (all-NA rows, kmers.py:187-192)
    -> BOTH ranks re-run with row gathers; empty: rank 1 holds no contig at all."""
    got = _run_featurise_case(case, "am_clr")
    (r0, calls0, err0, out0), (r1, calls1, err1, out1) = got
    assert err0 is None and err1 is None
    assert calls0 == calls1, (calls0, calls1)
    D = 512
    one = D * D + 2 * D + 2
    if case == "tiny":
        assert calls0 == [one, one]  # optimistic pass + the re-run without the all-NA rows
    else:
        assert calls0 == [one]
    ref, n_ref = _single_rank_reference(case, "am_clr")
    assert out0["n_norm"] + out1["n_norm"] == n_ref
    for out in (out0, out1):
        np.testing.assert_allclose(out["ev"], ref["explained_variance"], rtol=1e-9)
        cos = np.abs(np.sum(out["comps"] * ref["components"], axis=1))
        assert np.all(cos > 1 - 1e-9)
    scores = np.concatenate([out0["scores"], out1["scores"]], axis=0)
    np.testing.assert_allclose(scores, ref["scores"], atol=1e-8 * np.max(np.abs(ref["scores"])))


@pytest.mark.timeout(300)
def test_featurise_clr_all_zero_row_raises_on_every_rank():
    """clr / ilr: a contig without counts anywhere makes EVERY rank raise skbio's closure error
    (kmers.py:418-421), after the same single all-reduce."""
    got = _run_featurise_case("tiny", "clr")
    (r0, calls0, err0, _), (r1, calls1, err1, _) = got
    assert calls0 == calls1 and len(calls0) == 1
    assert err0 and err1 and "rows with all zeros" in err0 and "rows with all zeros" in err1
