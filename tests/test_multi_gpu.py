"""Two ranks on two GPUs: documents sharded by nnz, the K x V statistics / alpha statistics /
counters combined by the library's NCCL all-reduce; both ranks must end the iteration with the
oracle's single-process model.  Skipped on a box with fewer than two GPUs (the CPU coverage of the
exchange logic is tests/test_sharding_gloo.py)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rank(rank, world, id_path, out_dir, deterministic=0):
    sys.path.insert(0, ROOT)
    import time

    from mrlda_b200 import capi, corpus
    K, V = 20, 600
    full = corpus.generate(240, V, K, 50, seed=17)
    lb = corpus.init_logbeta(V, K, seed=2)
    alpha = np.full(K, 1e-3)
    mine, (lo, hi) = corpus.shard(full, rank, world)
    if rank == 0:
        with open(id_path + ".tmp", "wb") as f:
            f.write(capi.nccl_unique_id())
        os.rename(id_path + ".tmp", id_path)
    while not os.path.exists(id_path):
        time.sleep(0.05)
    uid = open(id_path, "rb").read()
    with capi.Context(K, V, device=rank, rank=rank, world_size=world, deterministic=deterministic) as ctx:
        ctx.comm_init(uid)
        ctx.set_corpus_csr(mine.row_ptr, mine.term_id, mine.count)
        ctx.set_alpha(alpha)
        ctx.set_logbeta(lb)
        r1 = ctx.iterate()
        r2 = ctx.iterate()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), ll=[r1["log_likelihood"], r2["log_likelihood"]],
                 n_docs=r1["n_docs"], n_tokens=r1["n_tokens"], seen=r1["n_terms_seen"],
                 alpha=ctx.get_alpha(), logbeta=ctx.get_logbeta(), gamma=ctx.get_gamma(), lo=lo, hi=hi,
                 allreduce_ms=r1["allreduce_ms"])


def test_two_gpu_iteration_matches_oracle(tmp_path, oracle):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from conftest import rel_err
    from mrlda_b200 import corpus
    mp.spawn(_rank, args=(2, str(tmp_path / "nccl.id"), str(tmp_path)), nprocs=2, join=True)
    K, V = 20, 600
    full = corpus.generate(240, V, K, 50, seed=17)
    lb = corpus.init_logbeta(V, K, seed=2)
    alpha = np.full(K, 1e-3)
    ref1 = oracle.iterate(K, V, full.row_ptr, full.term_id, full.count, lb, alpha)
    seen = ref1["present"].astype(bool)
    lb2 = np.where(seen[:, None], ref1["logbeta"], lb)
    ref2 = oracle.iterate(K, V, full.row_ptr, full.term_id, full.count, lb2, ref1["alpha"],
                          gamma=ref1["gamma"])
    for rank in range(2):
        got = np.load(tmp_path / f"rank{rank}.npz")
        assert int(got["n_docs"]) == full.n_docs and int(got["n_tokens"]) == full.n_tokens
        assert int(got["seen"]) == ref1["n_terms_seen"]
        for it, ref in enumerate((ref1, ref2)):
            assert abs(got["ll"][it] - ref["log_likelihood"]) <= 1e-9 * abs(ref["log_likelihood"])
        assert rel_err(got["alpha"], ref2["alpha"]) < 1e-8
        assert rel_err(got["logbeta"][seen], ref2["logbeta"][seen]) < 1e-8
        lo, hi = int(got["lo"]), int(got["hi"])
        assert rel_err(got["gamma"], ref2["gamma"][lo:hi]) < 1e-8
        assert float(got["allreduce_ms"]) > 0
    a, b = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert np.array_equal(a["logbeta"], b["logbeta"]) and np.array_equal(a["alpha"], b["alpha"])


def test_deterministic_mode_is_independent_of_the_rank_count(tmp_path):
    """cfg.deterministic: fixed-point statistics all-reduced as integers.  Two ranks on two GPUs and one
    rank on one GPU must produce bit-identical likelihood, alpha and logbeta over two EM iterations
    (the reference's shuffle sums each key's values in one reducer, whatever the mapper count)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from mrlda_b200 import capi, corpus
    mp.spawn(_rank, args=(2, str(tmp_path / "nccl.id"), str(tmp_path), 1), nprocs=2, join=True)
    K, V = 20, 600
    full = corpus.generate(240, V, K, 50, seed=17)
    with capi.Context(K, V, deterministic=1) as ctx:
        ctx.set_corpus_csr(full.row_ptr, full.term_id, full.count)
        ctx.set_alpha(np.full(K, 1e-3))
        ctx.set_logbeta(corpus.init_logbeta(V, K, seed=2))
        ll = [ctx.iterate()["log_likelihood"], ctx.iterate()["log_likelihood"]]
        alpha, logbeta, gamma = ctx.get_alpha(), ctx.get_logbeta(), ctx.get_gamma()
    for rank in range(2):
        got = np.load(tmp_path / f"rank{rank}.npz")
        assert list(got["ll"]) == ll
        assert np.array_equal(got["alpha"], alpha) and np.array_equal(got["logbeta"], logbeta)
        assert np.array_equal(got["gamma"], gamma[int(got["lo"]):int(got["hi"])])
