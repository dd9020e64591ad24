"""One EM iteration against the committed fixture tests/golden/iteration_vectors.json (four small
cases: vector alpha from a cold start, symmetric alpha, stored gamma, held-out mode).

The fixture holds what both CPU restatements of the reference's algorithm agreed on when it was
generated (tests/golden/make_iteration_vectors.py; the Java reference cannot run in this image, so
these are not reference outputs).  Not gpu: the C oracle still reproduces them.  gpu: the CUDA path
through the C ABI reproduces them within north_star's 1e-9, integer bookkeeping bit-exact."""
import json
import os
from types import SimpleNamespace

import numpy as np
import pytest

from conftest import rel_err

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "iteration_vectors.json")) as f:
    CASES = json.load(f)["cases"]


def _load(case):
    K, V = case["K"], case["V"]
    rp = np.array(case["row_ptr"], dtype=np.int64)
    D = len(rp) - 1
    inp = SimpleNamespace(K=K, V=V, D=D, row_ptr=rp, term_id=np.array(case["term_id"], dtype=np.int32),
                          count=np.array(case["count"], dtype=np.int32),
                          logbeta=np.array(case["logbeta"]).reshape(V, K), alpha=np.array(case["alpha"]),
                          gamma0=None if case["gamma0"] is None else np.array(case["gamma0"]).reshape(D, K),
                          symmetric=case["symmetric"], learning=case["learning"])
    e = case["expect"]
    exp = dict(counter=e["counter"], log_likelihood=-e["counter"] / 1e6, gamma=np.array(e["gamma"]).reshape(D, K),
               alpha_ss=np.array(e["alpha_ss"]))
    if case["learning"]:
        exp.update(n_terms_seen=e["n_terms_seen"], present=np.array(e["present"], dtype=np.uint8),
                   normaliser=np.array(e["normaliser"], dtype=np.float32),
                   digamma_lambda=np.array(e["digamma_lambda"]).reshape(V, K),
                   logbeta=np.array(e["logbeta"]).reshape(V, K), alpha=np.array(e["alpha"]))
    return inp, exp


def _compare(got, exp, inp, tol, counter_slack):
    assert abs(got["counter"] - exp["counter"]) <= counter_slack
    assert abs(got["log_likelihood"] - exp["log_likelihood"]) <= tol * abs(exp["log_likelihood"])
    nz = np.diff(inp.row_ptr) > 0
    assert rel_err(got["gamma"][nz], exp["gamma"][nz]) < tol
    assert rel_err(got["alpha_ss"], exp["alpha_ss"]) < tol
    if inp.learning:
        assert got["n_terms_seen"] == exp["n_terms_seen"]
        assert np.array_equal(got["present"], exp["present"])
        assert np.array_equal(got["normaliser"].view(np.uint32), exp["normaliser"].view(np.uint32))
        seen = exp["present"].astype(bool)
        # cells that received statistics (lambda >= 1e-3 <=> digamma(lambda) > -1001): the literal bound;
        # the dead cells (lambda ~ 1e-12) sit on the (x + 6) - 6 cancellation of the reference's own
        # digamma series and are compared in units of |ref|
        live = seen[:, None] & (exp["digamma_lambda"] > -1001.0)
        for key in ("digamma_lambda", "logbeta"):
            a, b = got[key], exp[key]
            assert np.max(np.abs(a[live] - b[live]) / np.maximum(np.abs(b[live]), 1.0)) < tol, key
            dead = seen[:, None] & ~live
            if dead.any():
                mag = np.abs(b[dead])
                assert np.max(np.abs(a[dead] - b[dead]) / (mag * (1.0 + 2e-6 * mag))) < tol, key
        assert rel_err(got["alpha"], exp["alpha"]) < tol


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_reproduces_the_fixture(oracle, case):
    inp, exp = _load(case)
    got = oracle.iterate(inp.K, inp.V, inp.row_ptr, inp.term_id, inp.count, inp.logbeta, inp.alpha,
                         gamma=inp.gamma0, learning=inp.learning, symmetric=inp.symmetric)
    _compare(got, exp, inp, 1e-12, 0)


def test_fixture_is_what_the_generator_writes():
    """Shapes and bookkeeping of the committed file (a stale or hand-edited fixture fails here)."""
    assert [c["name"] for c in CASES] == ["vector_alpha_cold_start", "symmetric_alpha", "stored_gamma", "heldout_mode"]
    for case in CASES:
        inp, exp = _load(case)
        assert inp.row_ptr[0] == 0 and inp.row_ptr[-1] == len(inp.term_id) == len(inp.count)
        assert inp.term_id.min() >= 1 and inp.term_id.max() <= inp.V and inp.count.min() >= 1
        if inp.learning:
            present = np.zeros(inp.V, np.uint8)
            present[inp.term_id - 1] = 1
            assert np.array_equal(present, exp["present"]) and exp["n_terms_seen"] == int(present.sum())
        # gamma sums: sum_k gamma_dk = sum_k alpha_k + N_d (phi rows sum to one)
        n_d = np.add.reduceat(inp.count, inp.row_ptr[:-1]) if inp.D else np.zeros(0)
        assert np.allclose(exp["gamma"].sum(1), inp.alpha.sum() + n_d, rtol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_cuda_path_reproduces_the_fixture(capi, case):
    inp, exp = _load(case)
    with capi.Context(inp.K, inp.V, learning=int(inp.learning), symmetric_alpha=int(inp.symmetric)) as ctx:
        ctx.set_corpus_csr(inp.row_ptr, inp.term_id, inp.count, gamma0=inp.gamma0)
        ctx.set_alpha(inp.alpha)
        ctx.set_logbeta(inp.logbeta)
        r = ctx.iterate()
        got = dict(counter=r["ll_counter"], log_likelihood=r["log_likelihood"], gamma=ctx.get_gamma(),
                   alpha_ss=ctx.get_alpha_ss())
        assert r["n_docs"] == inp.D and r["n_tokens"] == int(inp.count.sum())
        if inp.learning:
            dl, nm, pr = ctx.get_beta()
            got.update(n_terms_seen=r["n_terms_seen"], present=pr, normaliser=nm, digamma_lambda=dl,
                       logbeta=ctx.get_logbeta(), alpha=ctx.get_alpha())
    # the int64 counter sums per-document truncations: one unit per document of slack (test_gpu_parity.py)
    _compare(got, exp, inp, 1e-9, inp.D)
