"""GPU parity tests: the CUDA path through the C ABI against the CPU oracle on the same inputs.

Tolerances (BASELINE.json north_star): per-iteration ELBO and logbeta within 1e-9 relative in fp64;
integer bookkeeping (doc/term counts, presence) bit-exact.  The ELBO counter is the int64 sum of
per-document truncations (DocumentMapper.java:253-254); a 1-ulp difference in a document's
likelihood can move one truncation by one unit (1e-6), so counters are compared within D units and
the decoded likelihood within 1e-9 relative.
"""
import numpy as np
import pytest

from conftest import rel_err, small_corpus

pytestmark = pytest.mark.gpu

TOL = 1e-9


def _ctx(capi, c, **kw):
    return capi.Context(c.n_topics, c.n_terms, **kw)


def _run_pair(capi, oracle, c, alpha, logbeta, gamma0=None, **cfg):
    K, V = c.n_topics, c.n_terms
    with _ctx(capi, c, **cfg) as ctx:
        ctx.set_corpus_csr(c.row_ptr, c.term_id, c.count, gamma0=gamma0)
        ctx.set_alpha(alpha)
        ctx.set_logbeta(logbeta)
        r = ctx.iterate()
        out = dict(r=r, gamma=ctx.get_gamma(), alpha_ss=ctx.get_alpha_ss())
        if cfg.get("learning", 1):
            out.update(alpha=ctx.get_alpha(), logbeta=ctx.get_logbeta(), beta=ctx.get_beta(),
                       stats=ctx.get_stats())
    ref = oracle.iterate(K, V, c.row_ptr, c.term_id, c.count, logbeta, alpha, gamma=gamma0,
                         random_start=bool(cfg.get("random_start", 0)),
                         learning=bool(cfg.get("learning", 1)),
                         symmetric=bool(cfg.get("symmetric_alpha", 0)),
                         sweeps=cfg.get("sweep_cap", 100) - 1)
    return out, ref


def _lb_err(a, b):
    """Error measure for the beta file values digamma(lambda) and for logbeta = digamma(lambda) -
    normaliser, in units of the tolerance below (so `< TOL` reads like the other checks).

    Two properties of the reference's own arithmetic shape it.  (1) The values pass through zero
    (digamma crosses 0 at 1.46), where a purely relative error is ill-conditioned: |ref| is floored
    at 1.  (2) The lda-c digamma series computes 1/((x+6)-6): a last-bit difference in lambda moves
    x+6 to the neighbouring double, i.e. moves the result by up to ulp(6)/lambda = 8.9e-16/lambda
    relative (1e-9 at lambda = 1e-6, the known ~1e-3 at the dead cells' lambda = 1e-12).  Since
    |digamma(lambda)| ~ 1/lambda there, the admissible relative error is 1e-9 + 2e-15 * |ref|."""
    a, b = np.asarray(a), np.asarray(b)
    mag = np.maximum(np.abs(b), 1.0)
    return float(np.max(np.abs(a - b) / (mag * (1.0 + 2e-6 * mag))))


def _check(out, ref, c, learning=True):
    D = c.n_docs
    assert out["r"]["n_docs"] == D
    assert out["r"]["n_tokens"] == c.n_tokens
    assert abs(out["r"]["ll_counter"] - ref["counter"]) <= D
    assert abs(out["r"]["log_likelihood"] - ref["log_likelihood"]) <= TOL * abs(ref["log_likelihood"])
    nz = np.diff(c.row_ptr) > 0
    assert rel_err(out["gamma"][nz], ref["gamma"][nz]) < TOL
    assert rel_err(out["alpha_ss"], ref["alpha_ss"]) < TOL
    if learning:
        assert out["r"]["n_terms_seen"] == ref["n_terms_seen"]
        dl, nm, pr = out["beta"]
        assert np.array_equal(pr, ref["present"])
        seen = pr.astype(bool)
        assert np.array_equal(nm, ref["normaliser"])  # (float) cast, TermReducer.java:173
        assert _lb_err(dl[seen], ref["digamma_lambda"][seen]) < TOL  # digamma crosses 0 at 1.46
        assert _lb_err(out["logbeta"][seen], ref["logbeta"][seen]) < TOL
        assert rel_err(out["alpha"], ref["alpha"]) < TOL


@pytest.fixture(params=["auto", "general"])
def kernel_mode(request, monkeypatch):
    """auto: register-stationary kernel for documents that fit one SM, general kernel for the rest;
    general: force the general (shared-memory tile + L2 streaming) kernel for every document."""
    if request.param == "general":
        monkeypatch.setenv("MRLDA_ESTEP_KERNEL", "1")
    else:
        monkeypatch.delenv("MRLDA_ESTEP_KERNEL", raising=False)
    return request.param


@pytest.mark.parametrize("K", [3, 8, 20, 33, 50, 100, 130, 200, 256])
def test_one_iteration_matches_oracle(capi, oracle, K, kernel_mode):
    from mrlda_b200 import corpus
    c = small_corpus(D=40, V=400, K=K, mean_len=50, seed=K)
    alpha = np.full(K, 1e-3)
    lb = corpus.init_logbeta(c.n_terms, K, seed=K + 1)
    out, ref = _run_pair(capi, oracle, c, alpha, lb)
    _check(out, ref, c)
    assert out["r"]["total_launches"] >= 5


def test_large_k_variants(capi, oracle):
    from mrlda_b200 import corpus
    for K in (300, 600, 1000):
        c = small_corpus(D=6, V=150, K=K, mean_len=30, seed=K)
        alpha = np.full(K, 1e-3)
        lb = corpus.init_logbeta(c.n_terms, K, seed=5)
        out, ref = _run_pair(capi, oracle, c, alpha, lb, sweep_cap=12)
        _check(out, ref, c)


def test_stats_are_phi_weighted_counts(capi, oracle):
    """sum_k stats[v][k] = total count of term v (sum_k phi = 1), and stats match exp(log_phi)."""
    from mrlda_b200 import corpus
    K = 8
    c = small_corpus(D=50, V=200, K=K, seed=3)
    alpha = np.full(K, 0.1)
    lb = corpus.init_logbeta(c.n_terms, K, seed=9)
    out, _ = _run_pair(capi, oracle, c, alpha, lb)
    tot = np.bincount(c.term_id - 1, weights=c.count, minlength=c.n_terms)
    assert np.allclose(out["stats"].sum(1), tot, rtol=1e-12, atol=1e-12)
    e = oracle.estep(K, c.row_ptr, c.term_id, c.count, lb, alpha, want_log_phi=True)
    ref_stats = np.zeros((c.n_terms, K))
    np.add.at(ref_stats, c.term_id - 1, np.exp(e["log_phi"]))
    big = ref_stats > 1e-200
    assert rel_err(out["stats"][big], ref_stats[big]) < TOL
    # gamma conservation: sum_k gamma = sum_k alpha + N_d
    ntok = np.add.reduceat(c.count, c.row_ptr[:-1])
    assert np.allclose(out["gamma"].sum(1), alpha.sum() + ntok, rtol=1e-12)


def test_stored_gamma_and_random_start(capi, oracle, kernel_mode):
    from mrlda_b200 import corpus
    K = 20
    c = small_corpus(D=30, V=300, K=K, seed=4)
    alpha = np.full(K, 0.05)
    lb = corpus.init_logbeta(c.n_terms, K, seed=2)
    rng = np.random.default_rng(0)
    g0 = rng.gamma(1.0, 2.0, size=(c.n_docs, K)) + 0.01
    out, ref = _run_pair(capi, oracle, c, alpha, lb, gamma0=g0, sweep_cap=20)
    _check(out, ref, c)
    out2, ref2 = _run_pair(capi, oracle, c, alpha, lb, gamma0=g0, sweep_cap=20, random_start=1)
    _check(out2, ref2, c)
    assert rel_err(out["gamma"], out2["gamma"]) > 1e-6  # the stored gamma really was used


def test_test_mode_runs_e_step_only(capi, oracle):
    from mrlda_b200 import corpus
    K = 8
    c = small_corpus(D=25, V=200, K=K, seed=6)
    alpha = np.full(K, 0.2)
    lb = corpus.init_logbeta(c.n_terms, K, seed=3)
    out, ref = _run_pair(capi, oracle, c, alpha, lb, learning=0)
    _check(out, ref, c, learning=False)
    assert out["r"]["n_terms_seen"] == 0


def test_symmetric_alpha_iteration(capi, oracle):
    from mrlda_b200 import corpus
    K = 20
    c = small_corpus(D=80, V=300, K=K, seed=8)
    alpha = np.full(K, 1e-3)
    lb = corpus.init_logbeta(c.n_terms, K, seed=4)
    out, ref = _run_pair(capi, oracle, c, alpha, lb, symmetric_alpha=1)
    _check(out, ref, c)
    assert np.all(out["alpha"] == out["alpha"][0])


def test_empty_and_single_token_documents(capi, oracle, kernel_mode):
    """null content is counted and skipped (DocumentMapper.java:178,197-201); 1-token documents."""
    from mrlda_b200 import corpus
    K, V = 8, 50
    docs = [{1: 3, 7: 1}, None, {5: 1}, {}, {50: 2, 49: 1, 1: 1}, {2: 1}]
    c = corpus.from_docs(docs, V, K)
    alpha = np.full(K, 1e-3)
    lb = corpus.init_logbeta(V, K, seed=1)
    out, ref = _run_pair(capi, oracle, c, alpha, lb)
    _check(out, ref, c)
    assert out["r"]["n_docs"] == 6


def test_long_documents_exceed_the_tile(capi, oracle):
    """Documents with more distinct terms than one SM holds go to the general kernel and stream
    rows from L2; shorter ones in the same corpus take the register-stationary kernel."""
    from mrlda_b200 import corpus
    K = 200
    long_docs = corpus.generate(6, 3000, K, 1500, seed=11)
    short_docs = corpus.generate(10, 3000, K, 150, seed=12)
    n = np.concatenate([np.diff(long_docs.row_ptr), np.diff(short_docs.row_ptr)])
    rp = np.zeros(len(n) + 1, dtype=np.int64)
    rp[1:] = np.cumsum(n)
    c = corpus.Corpus(rp, np.concatenate([long_docs.term_id, short_docs.term_id]),
                      np.concatenate([long_docs.count, short_docs.count]), 3000, K)
    assert np.diff(c.row_ptr).max() > 300 and np.diff(c.row_ptr).min() < 190
    alpha = np.full(K, 1e-3)
    lb = corpus.init_logbeta(c.n_terms, K, seed=7)
    out, ref = _run_pair(capi, oracle, c, alpha, lb, sweep_cap=8)
    _check(out, ref, c)


def test_sparse_beta_takes_exact_slow_path(capi, oracle, kernel_mode):
    """A model where most (topic,term) cells are dead (lambda = eta => logbeta ~ -1e12) and the
    stored gamma points at the wrong topics: the linear-space norm underflows and the kernel must
    redo those rows in log space without NaNs."""
    K, V, D = 8, 40, 12
    rng = np.random.default_rng(5)
    from mrlda_b200 import corpus
    lb = np.full((V, K), -1.0e12)
    owner = rng.integers(0, K, size=V)
    lb[np.arange(V), owner] = np.log(rng.random(V) + 0.1)
    docs = []
    for d in range(D):
        terms = rng.choice(V, size=6, replace=False) + 1
        docs.append({int(t): int(rng.integers(1, 4)) for t in terms})
    c = corpus.from_docs(docs, V, K)
    alpha = np.full(K, 1e-3)
    g0 = np.full((D, K), 1e-3)
    g0[:, 0] = 30.0  # all mass on topic 0; theta_k underflows for k != 0
    out, ref = _run_pair(capi, oracle, c, alpha, lb, gamma0=g0, sweep_cap=6)
    assert np.all(np.isfinite(out["gamma"]))
    assert out["r"]["slow_path_rows"] > 0
    _check(out, ref, c)


@pytest.mark.parametrize("K", [300, 1000])
def test_sparse_beta_slow_path_in_cluster_kernels(capi, oracle, K):
    """The same situation with the topics split across a thread-block cluster (K > 208): every CTA
    of the cluster must take the same underflow decisions, read the other CTAs' gamma for the exact
    redo and add the result to its own topics only.  Documents of 3 to 120 distinct terms."""
    from mrlda_b200 import corpus
    V, D = 400, 24
    rng = np.random.default_rng(5)
    lb = np.full((V, K), -1.0e12)
    owner = rng.integers(0, K, size=V)
    lb[np.arange(V), owner] = np.log(rng.random(V) + 0.1)
    docs = []
    for d in range(D):
        terms = rng.choice(V, size=int(rng.integers(3, 120)), replace=False) + 1
        docs.append({int(t): int(rng.integers(1, 4)) for t in terms})
    c = corpus.from_docs(docs, V, K)
    alpha = np.full(K, 1e-3)
    g0 = np.full((D, K), 1e-3)
    g0[:, 0] = 30.0
    out, ref = _run_pair(capi, oracle, c, alpha, lb, gamma0=g0, sweep_cap=6)
    assert np.all(np.isfinite(out["gamma"]))
    assert out["r"]["slow_path_rows"] > 0
    _check(out, ref, c)


def test_three_em_iterations_track_oracle(capi, oracle, kernel_mode):
    """Per-iteration ELBO / logbeta / alpha parity over consecutive iterations (gamma carried)."""
    from mrlda_b200 import corpus
    K = 20
    c = small_corpus(D=120, V=500, K=K, mean_len=60, seed=12)
    alpha = np.full(K, 1e-3)
    lb = corpus.init_logbeta(c.n_terms, K, seed=13)
    gamma = None
    a_ref, lb_ref = alpha.copy(), lb.copy()
    with capi.Context(K, c.n_terms) as ctx:
        ctx.set_corpus_csr(c.row_ptr, c.term_id, c.count)
        ctx.set_alpha(alpha)
        ctx.set_logbeta(lb)
        for it in range(3):
            r = ctx.iterate()
            ref = oracle.iterate(K, c.n_terms, c.row_ptr, c.term_id, c.count, lb_ref, a_ref,
                                 gamma=gamma)
            assert abs(r["log_likelihood"] - ref["log_likelihood"]) <= TOL * abs(ref["log_likelihood"]), it
            seen = ref["present"].astype(bool)
            assert rel_err(ctx.get_logbeta()[seen], ref["logbeta"][seen]) < 10 * TOL, it
            assert rel_err(ctx.get_alpha(), ref["alpha"]) < 10 * TOL, it
            assert rel_err(ctx.get_gamma(), ref["gamma"]) < 10 * TOL, it
            gamma, a_ref = ref["gamma"], ref["alpha"]
            lb_ref = np.where(seen[:, None], ref["logbeta"], lb_ref)


def test_informed_prior_changes_eta_per_cell(capi, oracle):
    """-informedprior: log eta = (float)log 1000 for seed (topic, term) cells, (float)log 0.001 for
    the others (InformedPrior.java:43-44,172-177; TermReducer.java:162-164)."""
    from mrlda_b200 import corpus
    K = 8
    c = small_corpus(D=60, V=120, K=K, seed=31)
    alpha = np.full(K, 0.05)
    lb = corpus.init_logbeta(c.n_terms, K, seed=6)
    rng = np.random.default_rng(2)
    topics = rng.integers(1, K + 1, size=40)
    terms = rng.integers(1, c.n_terms + 1, size=40)
    seed = np.zeros((c.n_terms, K), dtype=np.uint8)
    seed[terms - 1, topics - 1] = 1
    ref = oracle.iterate(K, c.n_terms, c.row_ptr, c.term_id, c.count, lb, alpha, seed=seed)
    plain = oracle.iterate(K, c.n_terms, c.row_ptr, c.term_id, c.count, lb, alpha)
    with capi.Context(K, c.n_terms) as ctx:
        ctx.set_corpus_csr(c.row_ptr, c.term_id, c.count)
        ctx.set_alpha(alpha)
        ctx.set_logbeta(lb)
        ctx.set_informed_prior(topics, terms)
        ctx.iterate()
        dl, nm, pr = ctx.get_beta()
        got_lb = ctx.get_logbeta()
        with pytest.raises(capi.MrldaError):
            ctx.set_informed_prior([K + 1], [1])
        ctx.set_informed_prior([], [])  # removes the prior again
        ctx.set_logbeta(lb)
        ctx.set_alpha(alpha)
        ctx.set_corpus_csr(c.row_ptr, c.term_id, c.count)
        ctx.iterate()
        dl0, nm0, _ = ctx.get_beta()
    seen = ref["present"].astype(bool)
    assert np.array_equal(nm, ref["normaliser"]) and np.array_equal(nm0, plain["normaliser"])
    assert rel_err(dl[seen], ref["digamma_lambda"][seen]) < TOL
    assert rel_err(got_lb[seen], ref["logbeta"][seen]) < TOL
    assert rel_err(dl0[seen], plain["digamma_lambda"][seen]) < TOL
    assert rel_err(dl[seen], plain["digamma_lambda"][seen]) > 1e-3  # the prior really acted


@pytest.mark.parametrize("K", [8, 20, 100, 200])
def test_fp32_mode(capi, oracle, K):
    """Optional fp32 mode (single-precision tile / dot products / S, everything else fp64).

    Stated tolerance: per-iteration ELBO within 1e-6 relative always; logbeta (max over
    all cells) within 1e-4 once the topics have separated (from the third iteration of a fit), and
    within 5e-3 in the first two iterations from the reference's random initial beta, where the
    topics are near-identical and the split of a document's mass between them is ill-conditioned
    (the same amplification turns 1e-16 into ~1e-12 in fp64).  Each GPU iteration is compared with
    the oracle run on the same inputs."""
    from mrlda_b200 import corpus
    c = small_corpus(D=60, V=500, K=K, mean_len=80, seed=100 + K)
    alpha = np.full(K, 1e-3)
    lb = corpus.init_logbeta(c.n_terms, K, seed=K)
    with capi.Context(K, c.n_terms, fp32_mode=1) as ctx:
        ctx.set_corpus_csr(c.row_ptr, c.term_id, c.count)
        ctx.set_alpha(alpha)
        ctx.set_logbeta(lb)
        g_alpha, g_lb, g_gamma = alpha, lb, None
        for it in range(3):
            r = ctx.iterate()
            ref = oracle.iterate(K, c.n_terms, c.row_ptr, c.term_id, c.count, g_lb, g_alpha,
                                 gamma=g_gamma)
            seen = ref["present"].astype(bool)
            g_lb, g_alpha, g_gamma = ctx.get_logbeta(), ctx.get_alpha(), ctx.get_gamma()
            assert r["n_docs"] == c.n_docs and r["n_tokens"] == c.n_tokens
            assert np.array_equal(ctx.get_beta()[2], ref["present"])
            assert abs(r["log_likelihood"] - ref["log_likelihood"]) <= 1e-6 * abs(ref["log_likelihood"])
            assert _lb_err(g_lb[seen], ref["logbeta"][seen]) < (5e-3 if it < 2 else 1e-4), it
            assert rel_err(g_alpha, ref["alpha"]) < 1e-5
    # and it really ran in single precision: not bit-identical to the fp64 mode
    out64, _ = _run_pair(capi, oracle, c, alpha, lb)
    out32, _ = _run_pair(capi, oracle, c, alpha, lb, fp32_mode=1)
    assert rel_err(out32["gamma"], out64["gamma"]) > 1e-12


def test_alpha_known_answers_on_device(capi):
    """The reference's own KATs (VariationalInferenceTest.java:27-62) through the device kernels."""
    with capi.Context(3, 10) as ctx:
        out = ctx.update_vector_alpha(
            3, 112, [0.4736839726180464, 9.928726975283879, 8.319361678447014],
            [-23792.9569126969113, -22519.9434073184025, -23973.2360888324797])
        want = [0.4736839726180464, 9.92872697528388, 8.319361678447015]
        assert np.max(np.abs(out - want)) < 1e-10
        kats = [(5, -40100.9192398908126052, 0.2958548131184747),
                (5, -34828.2371112336259102, 0.3731832583179411),
                (5, -37309.1699276268700487, 0.3319329678764105),
                (5, -44085.8660385293114814, 0.2568195157403902),
                (10, -155990.5727383689954877, 0.1531475153565107),
                (10, -196359.2521305996051524, 0.1150183709445565),
                (10, -226577.3570433593704365, 0.0972395316113154),
                (10, -256318.9209672076685820, 0.0845206104885002)]
        for K, ss, want in kats:
            assert abs(ctx.update_scalar_alpha(K, 2246, 100, ss) - want) < 1e-10


def test_device_special_functions_match_oracle(capi, oracle):
    x = np.concatenate([np.logspace(-12, 6, 400), [1e-3, 1.0, 1.4616321449683623, 2.5]])
    with capi.Context(3, 10) as ctx:
        dg = ctx.eval_special("digamma", x)
        tg = ctx.eval_special("trigamma", x)
        lg = ctx.eval_special("lngamma", x)
        ed = ctx.eval_special("exp_digamma", x)
        edn = ctx.eval_special("exp_digamma_n", x)
    want_dg = np.array([oracle.digamma(v) for v in x])
    assert np.array_equal(dg, want_dg) or rel_err(dg, want_dg) < 1e-13
    assert rel_err(tg, [oracle.trigamma(v) for v in x]) < 1e-13
    lgw = np.array([oracle.lngamma(v) for v in x])
    assert np.max(np.abs(lg - lgw) / np.maximum(np.abs(lgw), 1.0)) < 1e-13
    big = x > 2e-3  # below, exp(digamma) underflows towards 0
    assert rel_err(ed[big], np.exp(want_dg[big])) < 1e-12
    # the interleaved short-chain form of the warp-group kernel: same series, same zeros
    assert rel_err(edn[big], np.exp(want_dg[big])) < 1e-12
    assert rel_err(edn[big], ed[big]) < 1e-12  # |u| ulp(1): exp(u) amplifies the rounding of u near -500
    assert np.all(edn[~big] >= 0.0) and np.all(edn[~big] < 1e-200)


def test_all_documents_empty_and_no_documents(capi, oracle):
    from mrlda_b200 import corpus
    K, V = 4, 20
    c = corpus.from_docs([None, {}, None], V, K)
    with capi.Context(K, V) as ctx:
        ctx.set_corpus_csr(c.row_ptr, c.term_id, c.count)
        ctx.set_alpha(np.full(K, 0.5))
        ctx.set_logbeta(corpus.init_logbeta(V, K, seed=1))
        r = ctx.e_step()  # every document has null content: counted, nothing computed
        assert r["n_docs"] == 3 and r["n_tokens"] == 0 and r["ll_counter"] == 0
        assert np.all(ctx.get_gamma() == 0)
        ctx.set_corpus_csr([0], [], [])
        with pytest.raises(capi.MrldaError):
            ctx.e_step()


def test_many_topic_counts_cover_every_kernel_family(capi, oracle):
    """One short E-step per K: every (warps, columns-per-warp) family of the register-stationary
    kernel, the cluster variants and the general kernel's lane layouts."""
    from mrlda_b200 import corpus
    for K in (2, 5, 13, 27, 40, 64, 77, 104, 111, 160, 208, 230, 257, 400, 513, 800, 1024):
        c = small_corpus(D=9, V=120, K=K, mean_len=25, seed=K)
        alpha = np.full(K, 0.01)
        lb = corpus.init_logbeta(c.n_terms, K, seed=3)
        out, ref = _run_pair(capi, oracle, c, alpha, lb, sweep_cap=6)
        _check(out, ref, c)


def test_error_paths(capi):
    with pytest.raises(capi.MrldaError):
        capi.Context(0, 10)
    with pytest.raises(capi.MrldaError):
        capi.Context(5000, 10)
    with capi.Context(4, 10) as ctx:
        with pytest.raises(capi.MrldaError):
            ctx.iterate()  # nothing set
        with pytest.raises(capi.MrldaError):
            ctx.set_corpus_csr([0, 1], [11], [1])  # term id out of range
        with pytest.raises(capi.MrldaError):
            ctx.set_alpha([0.1, 0.0, 0.1, 0.1])


@pytest.mark.parametrize("K", [20, 200, 600])
def test_deterministic_mode_is_bit_reproducible(capi, oracle, K):
    """cfg.deterministic: the statistics are accumulated in fixed point with integer atomics, so two
    runs on the same inputs give bit-identical beta, alpha, statistics and likelihood (the default
    mode differs at 1e-16 from run to run through the order of the fp64 atomics) — and they stay
    within the parity tolerance of the oracle.  The documents are also presented in a different
    order (same set): the statistics are a sum over documents, so they must not move at all."""
    from mrlda_b200 import corpus
    c = small_corpus(D=300, V=500, K=K, mean_len=60, seed=K + 2)
    alpha = np.full(K, 1e-3)
    lb = corpus.init_logbeta(c.n_terms, K, seed=K + 3)
    runs = []
    for _ in range(2):
        out, ref = _run_pair(capi, oracle, c, alpha, lb, deterministic=1, sweep_cap=30)
        runs.append(out)
    _check(runs[0], ref, c)
    a, b = runs
    assert a["r"]["ll_counter"] == b["r"]["ll_counter"]
    for key in ("gamma", "alpha_ss", "alpha", "logbeta", "stats"):
        assert np.array_equal(a[key], b[key]), key
    assert all(np.array_equal(x, y) for x, y in zip(a["beta"], b["beta"]))
    tot = np.bincount(c.term_id - 1, weights=c.count, minlength=c.n_terms)
    assert np.allclose(a["stats"].sum(1), tot, rtol=1e-13, atol=1e-13)
    # the same documents in reverse order: per-document results are unchanged, so the sums are too
    order = np.arange(c.n_docs)[::-1]
    n = np.diff(c.row_ptr)
    rp = np.zeros(c.n_docs + 1, dtype=np.int64)
    rp[1:] = np.cumsum(n[order])
    idx = np.concatenate([np.arange(c.row_ptr[d], c.row_ptr[d + 1]) for d in order])
    with capi.Context(K, c.n_terms, deterministic=1, sweep_cap=30) as ctx:
        ctx.set_corpus_csr(rp, c.term_id[idx], c.count[idx])
        ctx.set_alpha(alpha)
        ctx.set_logbeta(lb)
        r = ctx.iterate()
        assert r["ll_counter"] == a["r"]["ll_counter"]
        assert np.array_equal(ctx.get_stats(), a["stats"])
        assert np.array_equal(ctx.get_logbeta(), a["logbeta"])
        assert np.array_equal(ctx.get_alpha(), a["alpha"])
        assert np.array_equal(ctx.get_gamma()[np.argsort(order)], a["gamma"])
