"""world_size-2 gloo test of the multi-GPU host logic on CPU: documents are sharded by nnz, every
rank runs the E-step on its shard (the oracle stands in for the device here), the K x V statistics,
alpha statistics and int64 counters are summed across ranks — exactly the exchange the library does
with one NCCL all-reduce — and the result must equal the single-rank run."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _shard_stats(c, lb, alpha):
    from oracle import oracle as O
    K = c.n_topics
    r = O.estep(K, c.row_ptr, c.term_id, c.count, lb, alpha, want_log_phi=True)
    stats = np.zeros((c.n_terms, K))
    if c.nnz:
        np.add.at(stats, c.term_id - 1, np.exp(r["log_phi"]))
    return stats, r["alpha_ss"], r["counter"], r["gamma"]


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    from mrlda_b200 import corpus
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = corpus.generate(50, 150, 5, 25, seed=9)
    lb = corpus.init_logbeta(150, 5, seed=1)
    alpha = np.full(5, 0.01)
    mine, (lo, hi) = corpus.shard(full, rank, world)
    stats, ss, counter, gamma = _shard_stats(mine, lb, alpha)
    t_stats, t_ss = torch.from_numpy(stats), torch.from_numpy(ss)
    t_cnt = torch.tensor([counter, mine.n_docs, mine.n_tokens], dtype=torch.int64)
    for t in (t_stats, t_ss, t_cnt):
        dist.all_reduce(t)
    if rank == 0:
        np.savez(out_path, stats=t_stats.numpy(), ss=t_ss.numpy(), cnt=t_cnt.numpy(), lo=lo, hi=hi,
                 gamma=gamma)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_exchange_equals_single_rank(tmp_path):
    sys.path.insert(0, ROOT)
    from mrlda_b200 import corpus
    out = str(tmp_path / "r0.npz")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    full = corpus.generate(50, 150, 5, 25, seed=9)
    lb = corpus.init_logbeta(150, 5, seed=1)
    alpha = np.full(5, 0.01)
    stats, ss, counter, gamma = _shard_stats(full, lb, alpha)
    assert np.allclose(got["stats"], stats, rtol=1e-13, atol=1e-300)
    assert np.allclose(got["ss"], ss, rtol=1e-13)
    assert list(got["cnt"]) == [counter, full.n_docs, full.n_tokens]  # integers: bit-exact
    assert np.array_equal(got["gamma"], gamma[int(got["lo"]):int(got["hi"])])


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_shard_bounds_partition_and_balance(world):
    sys.path.insert(0, ROOT)
    from mrlda_b200 import corpus
    c = corpus.generate(500, 300, 4, 40, seed=5)
    b = corpus.shard_bounds(c.row_ptr, world)
    assert b[0] == 0 and b[-1] == c.n_docs and len(b) == world + 1
    assert all(x <= y for x, y in zip(b, b[1:]))
    nnz = [int(c.row_ptr[y] - c.row_ptr[x]) for x, y in zip(b, b[1:])]
    assert sum(nnz) == c.nnz
    assert max(nnz) - min(nnz) <= 2 * int(np.diff(c.row_ptr).max())
    parts = [corpus.shard(c, r, world)[0] for r in range(world)]
    assert np.array_equal(np.concatenate([p.term_id for p in parts]), c.term_id)
    assert np.array_equal(np.concatenate([p.count for p in parts]), c.count)


def test_shard_bounds_edge_cases():
    sys.path.insert(0, ROOT)
    from mrlda_b200 import corpus
    assert corpus.shard_bounds(np.array([0]), 4) == [0, 0, 0, 0, 0]  # empty corpus
    assert corpus.shard_bounds(np.array([0, 5]), 2)[-1] == 1         # one document, two ranks
    c = corpus.from_docs([{1: 1}, None, {}, {2: 2}], 5, 2)
    assert c.n_docs == 4 and c.nnz == 2 and list(np.diff(c.row_ptr)) == [1, 0, 0, 1]
