"""The oracle pinned against the reference's own known-answer tests and checked for the invariants
the algorithm guarantees.  Runs without a GPU.

Golden vectors: src/test/java/cc/mrlda/VariationalInferenceTest.java:27-62 (verbatim below)."""
import math

import numpy as np
import pytest

VECTOR_KAT = dict(
    K=3, D=112,
    alpha=[0.4736839726180464, 9.928726975283879, 8.319361678447014],
    ss=[-23792.9569126969113, -22519.9434073184025, -23973.2360888324797],
    want=[0.4736839726180464, 9.92872697528388, 8.319361678447015])

SCALAR_KATS = [
    (5, 2246, 100, -40100.9192398908126052, 0.2958548131184747),
    (5, 2246, 100, -34828.2371112336259102, 0.3731832583179411),
    (5, 2246, 100, -37309.1699276268700487, 0.3319329678764105),
    (5, 2246, 100, -44085.8660385293114814, 0.2568195157403902),
    (10, 2246, 100, -155990.5727383689954877, 0.1531475153565107),
    (10, 2246, 100, -196359.2521305996051524, 0.1150183709445565),
    (10, 2246, 100, -226577.3570433593704365, 0.0972395316113154),
    (10, 2246, 100, -256318.9209672076685820, 0.0845206104885002),
]
PRECISION_10 = 1e-10  # VariationalInferenceTest.java:25


def test_update_alpha_vector_kat(oracle):
    k = VECTOR_KAT
    out = oracle.update_vector_alpha(k["K"], k["D"], k["alpha"], k["ss"])
    assert len(out) == len(k["want"])
    assert np.max(np.abs(out - np.array(k["want"]))) < PRECISION_10


@pytest.mark.parametrize("K,D,init,ss,want", SCALAR_KATS)
def test_update_alpha_scalar_kat(oracle, K, D, init, ss, want):
    assert abs(oracle.update_scalar_alpha(K, D, init, ss) - want) < PRECISION_10


def test_digamma_trigamma_against_scipy(oracle):
    """The lda-c series is an approximation: close to, not equal to, the true functions."""
    sp = pytest.importorskip("scipy.special")
    for x in [1e-3, 0.1, 0.5, 1.0, 2.5, 10.0, 123.4, 1e4]:
        assert abs(oracle.digamma(x) - sp.digamma(x)) < 2e-9 * max(1.0, abs(sp.digamma(x)))
        assert abs(oracle.trigamma(x) - sp.polygamma(1, x)) < 1e-7 * sp.polygamma(1, x)
        assert abs(oracle.lngamma(x) - sp.gammaln(x)) < 1e-12 * max(1.0, abs(sp.gammaln(x)))
    # the (x+6)-6 cancellation the finalize kernel must reproduce: deterministic, ~1e-3 off 1/x
    v = oracle.digamma(1e-12)
    assert -1.2e12 < v < -0.8e12


def test_logadd(oracle):
    assert oracle.logadd(0.0, 0.0) == pytest.approx(math.log(2.0), rel=1e-15)
    assert oracle.logadd(-1e12, 3.0) == 3.0
    assert oracle.logadd(float("-inf"), float("-inf")) == float("-inf")
    assert oracle.logadd(1.0, 2.0) == oracle.logadd(2.0, 1.0)
    oracle.set_logadd_cutoff(20.0)
    try:
        assert oracle.logadd(0.0, -21.0) == 0.0
        assert oracle.logadd(0.0, -19.0) > 0.0
    finally:
        oracle.set_logadd_cutoff(-1.0)


def _toy():
    from mrlda_b200 import corpus
    c = corpus.generate(40, 200, 6, 30, seed=2)
    alpha = np.full(6, 0.05)
    lb = corpus.init_logbeta(200, 6, seed=3)
    return c, alpha, lb


def test_estep_invariants(oracle):
    c, alpha, lb = _toy()
    K = c.n_topics
    r = oracle.estep(K, c.row_ptr, c.term_id, c.count, lb, alpha, want_log_phi=True)
    ntok = np.add.reduceat(c.count, c.row_ptr[:-1])
    # gamma = alpha + sum_w n_w phi_w  =>  sum_k gamma_k = sum_k alpha_k + N_d
    assert np.allclose(r["gamma"].sum(1), alpha.sum() + ntok, rtol=1e-12)
    # sum_k phi_wk = 1 for every (doc, term): exp(log_phi) sums to the count
    assert np.allclose(np.exp(r["log_phi"]).sum(1), c.count, rtol=1e-12)
    # the counter is the sum of per-document truncations (DocumentMapper.java:253-254)
    assert r["counter"] == int(r["doc_counter"].sum())
    assert np.all(r["doc_counter"] > 0)  # -LL * 1e6 is positive: LL is a log probability bound


def test_estep_linear_space_identity(oracle):
    """The identity the CUDA kernel relies on: the log-space sweep equals
    gamma_k = alpha_k + theta_k sum_w n_w E_wk / (sum_k theta_k E_wk)."""
    c, alpha, lb = _toy()
    K = c.n_topics
    one = oracle.estep(K, c.row_ptr, c.term_id, c.count, lb, alpha, sweeps=1)
    ntok = np.add.reduceat(c.count, c.row_ptr[:-1])
    for d in range(c.n_docs):
        b, e = c.row_ptr[d], c.row_ptr[d + 1]
        g0 = alpha + np.float64(np.float32(ntok[d]) / np.float32(K))
        theta = np.exp([oracle.digamma(x) for x in g0])
        E = np.exp(lb[c.term_id[b:e] - 1])
        norm = E @ theta
        g1 = alpha + theta * ((c.count[b:e] / norm) @ E)
        assert np.allclose(g1, one["gamma"][d], rtol=1e-12)


def test_iterate_elbo_improves_and_is_permutation_invariant(oracle):
    c, alpha, lb = _toy()
    K, V = c.n_topics, c.n_terms
    r1 = oracle.iterate(K, V, c.row_ptr, c.term_id, c.count, lb, alpha)
    lb2 = np.where(r1["present"].astype(bool)[:, None], r1["logbeta"], lb)
    r2 = oracle.iterate(K, V, c.row_ptr, c.term_id, c.count, lb2, alpha, gamma=r1["gamma"])
    lb3 = np.where(r2["present"].astype(bool)[:, None], r2["logbeta"], lb2)
    r3 = oracle.iterate(K, V, c.row_ptr, c.term_id, c.count, lb3, alpha, gamma=r2["gamma"])
    # iteration 1 scores the un-normalised random beta (rows of exp(logbeta) sum to ~V/2), so its
    # "likelihood" is not a bound; from iteration 2 on the ELBO rises with alpha held fixed
    assert r3["log_likelihood"] > r2["log_likelihood"]
    assert r1["n_terms_seen"] == len(np.unique(c.term_id))
    # documents in reverse order: same model up to summation-order rounding
    order = np.arange(c.n_docs)[::-1]
    n = np.diff(c.row_ptr)
    rp = np.zeros(c.n_docs + 1, dtype=np.int64)
    rp[1:] = np.cumsum(n[order])
    idx = np.concatenate([np.arange(c.row_ptr[d], c.row_ptr[d + 1]) for d in order])
    rr = oracle.iterate(K, V, rp, c.term_id[idx], c.count[idx], lb, alpha)
    assert abs(rr["log_likelihood"] - r1["log_likelihood"]) < 1e-5
    seen = r1["present"].astype(bool)
    assert np.allclose(rr["logbeta"][seen], r1["logbeta"][seen], rtol=1e-12)
    assert np.allclose(rr["gamma"][::-1], r1["gamma"], rtol=1e-12)


def test_import_beta_float_normaliser(oracle):
    """logbeta = value - (double)(float)normaliser; rows absent from the file stay 0
    (DocumentMapper.java:511,514-523)."""
    K, V = 3, 4
    dl = np.arange(12, dtype=np.float64).reshape(V, K) - 20.0
    nm = np.array([0.1, 1.0 / 3.0, 2.5], dtype=np.float32)
    pr = np.array([1, 0, 1, 1], dtype=np.uint8)
    out = oracle.import_beta(K, V, dl, nm, pr)
    assert np.array_equal(out[1], np.zeros(K))
    assert out[0, 1] == dl[0, 1] - float(np.float32(1.0 / 3.0))
    assert out[3, 0] == dl[3, 0] - float(np.float32(0.1))


def test_cutoff_mode_gap_is_reported_not_hidden(oracle):
    """The unpinned LogMath.add drop threshold: with a 20-nat cutoff the oracle moves by
    ~1e-8 relative, i.e. more than the 1e-9 parity tolerance — which is why the exact mode is the
    declared target and this gap is documented in DESIGN.md."""
    c, alpha, lb = _toy()
    K, V = c.n_topics, c.n_terms
    exact = oracle.iterate(K, V, c.row_ptr, c.term_id, c.count, lb, alpha)
    oracle.set_logadd_cutoff(20.0)
    try:
        cut = oracle.iterate(K, V, c.row_ptr, c.term_id, c.count, lb, alpha)
    finally:
        oracle.set_logadd_cutoff(-1.0)
    gap = np.max(np.abs(cut["gamma"] - exact["gamma"]) / exact["gamma"])
    assert 0 < gap < 1e-6
