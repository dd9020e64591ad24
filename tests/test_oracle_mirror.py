"""Pins oracle/mrlda_oracle.c (the checker of every GPU parity test) with oracle/mirror.py, a second
restatement of DocumentMapper.map / updatePhi (DocumentMapper.java:175-259,402-429) and
TermCombiner / TermReducer / importBeta (TermReducer.java:134-238, DocumentMapper.java:475-536)
written independently from the Java, in NumPy with the reference's fold order and in 40-digit
mpmath.  Runs without a GPU.

Tolerance: 1e-13 relative (the two float64 restatements differ only in libm vs NumPy exp/log
rounding).  The beta-file values digamma(lambda) get 1e-13 where lambda >= 1e-3; below that the
reference's own series evaluates 1/((x+6)-6), so a last-bit difference in lambda can move the result
by ulp(6)/lambda and the bound is 1e-13 + 1.8e-15/lambda."""
import numpy as np
import pytest

from oracle import mirror as M

TOL = 1e-13


def _rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def _case(D, V, K, mean_len, seed, alpha0=1e-3):
    from mrlda_b200 import corpus
    c = corpus.generate(D, V, K, mean_len, seed=seed)
    return c, np.full(K, alpha0), corpus.init_logbeta(V, K, seed=seed + 100)


def test_digamma_matches_c_oracle_bitwise_or_ulp(oracle):
    xs = np.concatenate([np.logspace(-12, 5, 400), [1e-3, 0.5, 1.0, 2.0, 100.0]])
    want = np.array([oracle.digamma(x) for x in xs])
    got = M.digamma(xs)
    assert _rel(got, want) <= 4e-16 * 8  # log() may differ in the last bit between libm and NumPy


@pytest.mark.parametrize("D,V,K,mean_len,seed", [(12, 120, 5, 25, 1), (8, 300, 20, 60, 2), (5, 80, 3, 10, 3)])
def test_map_phase_matches_c_oracle(oracle, D, V, K, mean_len, seed):
    c, alpha, lb = _case(D, V, K, mean_len, seed)
    ref = oracle.estep(K, c.row_ptr, c.term_id, c.count, lb, alpha, want_log_phi=True)
    got = M.map_phase(K, c.row_ptr, c.term_id, c.count, lb, alpha)
    assert _rel(got["gamma"], ref["gamma"]) <= TOL
    assert _rel(got["alpha_ss"], ref["alpha_ss"]) <= TOL
    # log(n phi): cells far below the row maximum are differences of large numbers; compare the
    # probabilities n*phi they encode and the log values where phi is not negligible
    big = ref["log_phi"] > -30
    assert _rel(got["log_phi"][big], ref["log_phi"][big]) <= TOL * 30
    assert _rel(np.exp(got["log_phi"]), np.exp(ref["log_phi"])) <= 1e-12
    # the int64 LOG_LIKELIHOOD counter: (long)(-LL * 1e6) per document, so +-1 unit per document
    assert np.max(np.abs(got["doc_counter"] - ref["doc_counter"])) <= 1
    assert abs(got["counter"] - ref["counter"]) <= D


def test_stored_gamma_start_matches_c_oracle(oracle):
    c, alpha, lb = _case(10, 150, 6, 30, 7)
    first = oracle.estep(6, c.row_ptr, c.term_id, c.count, lb, alpha)
    ref = oracle.estep(6, c.row_ptr, c.term_id, c.count, lb, alpha, gamma=first["gamma"])
    got = M.map_phase(6, c.row_ptr, c.term_id, c.count, lb, alpha, gamma=first["gamma"])
    assert _rel(got["gamma"], ref["gamma"]) <= TOL


@pytest.mark.parametrize("D,V,K,mean_len,seed", [(30, 200, 6, 30, 4), (20, 500, 16, 50, 5)])
def test_reduce_terms_matches_c_oracle(oracle, D, V, K, mean_len, seed):
    c, alpha, lb = _case(D, V, K, mean_len, seed)
    ref = oracle.iterate(K, V, c.row_ptr, c.term_id, c.count, lb, alpha)
    e = oracle.estep(K, c.row_ptr, c.term_id, c.count, lb, alpha, want_log_phi=True)
    dl, nrm, present, lam = M.reduce_terms(K, V, c.row_ptr, c.term_id, e["log_phi"])
    assert np.array_equal(present, ref["present"])
    assert int(present.sum()) == ref["n_terms_seen"]
    # the (float) cast of the topic normaliser is a discontinuity: identical bits are required
    assert np.array_equal(nrm.view(np.uint32), ref["normaliser"].view(np.uint32))
    seen = present.astype(bool)
    want = ref["digamma_lambda"][seen]
    got = dl[seen]
    tol = TOL + 1.8e-15 / np.maximum(lam[seen], 1e-300) * (lam[seen] < 1e-3)
    assert np.all(np.abs(got - want) <= tol * np.abs(want))
    lbn = M.import_beta(dl, nrm, present)
    assert np.all(np.abs(lbn[seen] - ref["logbeta"][seen]) <= tol * np.maximum(np.abs(want), np.abs(ref["logbeta"][seen])))
    assert np.all(lbn[~seen] == 0.0) and np.all(ref["logbeta"][~seen] == 0.0)


def test_whole_iteration_mirror_end_to_end(oracle):
    """map phase and reduce both from the mirror (nothing shared with the C file)."""
    c, alpha, lb = _case(25, 180, 5, 30, 9)
    K, V = 5, 180
    ref = oracle.iterate(K, V, c.row_ptr, c.term_id, c.count, lb, alpha)
    m = M.map_phase(K, c.row_ptr, c.term_id, c.count, lb, alpha)
    dl, nrm, present, lam = M.reduce_terms(K, V, c.row_ptr, c.term_id, m["log_phi"])
    seen = present.astype(bool)
    assert abs(m["counter"] - ref["counter"]) <= c.n_docs
    assert np.array_equal(nrm.view(np.uint32), ref["normaliser"].view(np.uint32))
    ok = lam[seen] >= 1e-3
    assert _rel(dl[seen][ok], ref["digamma_lambda"][seen][ok]) <= 1e-12


def test_mpmath_agrees_with_both(oracle):
    """40-digit evaluation of the same formulas: float64 rounding is the only difference, so this
    bounds the rounding error of the oracle itself (and of the fold order it uses)."""
    pytest.importorskip("mpmath")
    c, alpha, lb = _case(3, 40, 4, 8, 11, alpha0=0.05)
    for d in range(c.n_docs):
        b, e = int(c.row_ptr[d]), int(c.row_ptr[d + 1])
        g_mp, lp_mp, ll_mp = M.map_document_mp(c.term_id[b:e], c.count[b:e], lb, alpha)
        g_np, lp_np, ll_np, _ = M.map_document(c.term_id[b:e], c.count[b:e], lb, alpha)
        one = oracle.estep(4, np.array([0, e - b]), c.term_id[b:e], c.count[b:e], lb, alpha, want_log_phi=True)
        assert _rel(g_np, g_mp) <= 1e-11
        assert _rel(one["gamma"][0], g_mp) <= 1e-11
        assert abs(ll_np - ll_mp) <= 1e-11 * abs(ll_mp)
        assert abs(-one["counter"] / 1e6 - ll_mp) <= 1e-6 + 1e-11 * abs(ll_mp)
        assert _rel(np.exp(lp_np), np.exp(np.array(lp_mp))) <= 1e-10


def test_alpha_updates_match_reference_kats_and_c_oracle(oracle):
    """updateVectorAlpha (Java array aliasing included) and updateScalarAlpha restated a second time
    from VariationalInference.java:409-511,573-625: the reference's own known answers
    (VariationalInferenceTest.java:29-61) and 200 random cases against the C oracle, among them
    off-model statistics that force the step-halving ("singular Hessian") branch."""
    got = M.update_vector_alpha(3, 112, [0.4736839726180464, 9.928726975283879, 8.319361678447014],
                                [-23792.9569126969113, -22519.9434073184025, -23973.2360888324797])
    assert np.max(np.abs(got - np.array([0.4736839726180464, 9.92872697528388, 8.319361678447015]))) <= 1e-10
    for K, ss, want in ((5, -40100.9192398908126052, 0.2958548131184747), (5, -34828.2371112336259102, 0.3731832583179411),
                        (5, -37309.1699276268700487, 0.3319329678764105), (5, -44085.8660385293114814, 0.2568195157403902),
                        (10, -155990.5727383689954877, 0.1531475153565107), (10, -196359.2521305996051524, 0.1150183709445565),
                        (10, -226577.3570433593704365, 0.0972395316113154), (10, -256318.9209672076685820, 0.0845206104885002)):
        assert abs(M.update_scalar_alpha(K, 2246, 100, ss) - want) <= 1e-10
    rng = np.random.default_rng(0)
    shrunk = 0
    for trial in range(200):
        K = int(rng.integers(2, 30))
        D = int(rng.integers(5, 3000))
        alpha = rng.uniform(0.001, 3.0, K) * rng.choice([1, 1, 1, 0.01, 10])
        gam = rng.gamma(0.5, 2.0, (D, K)) + 1e-3
        ss = (M.digamma(gam) - M.digamma(gam.sum(1))[:, None]).sum(0)
        if trial % 5 == 0:
            ss = ss * rng.uniform(0.2, 3)
        a, b = M.update_vector_alpha(K, D, alpha, ss), oracle.update_vector_alpha(K, D, alpha, ss)
        assert _rel(a, b) <= 1e-12, (trial, K, D)
        shrunk += int(np.any(a < 0.5 * alpha))
        s = M.update_scalar_alpha(K, D, float(alpha.mean()), float(ss.sum()))
        t = oracle.update_scalar_alpha(K, D, float(alpha.mean()), float(ss.sum()))
        assert (np.isnan(s) and np.isnan(t)) or abs(s - t) <= 1e-12 * abs(t), (trial, K, D)
    assert shrunk > 0  # some cases really took large steps
