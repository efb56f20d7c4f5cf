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
