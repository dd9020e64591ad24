"""oracle/mirror.py — second, independently written restatement of the reference's E-step and M-step
arithmetic, in NumPy (float64, same operation order as the Java) and in mpmath (40 digits).

TEST INFRASTRUCTURE ONLY, like everything under oracle/: only tests/ imports it.  Its job is to pin
oracle/mrlda_oracle.c (the C restatement the CUDA kernels are checked against) with something that
was not written by looking at that file: the functions below were written from the Java sources

  * DocumentMapper.map          /root/reference/src/main/java/cc/mrlda/DocumentMapper.java:175-259
  * DocumentMapper.updatePhi    DocumentMapper.java:402-429
  * phi emission (-directemit)  DocumentMapper.java:315-329
  * TermCombiner.reduce         TermCombiner.java:19-35
  * TermReducer.reduce / close  TermReducer.java:134-238
  * DocumentMapper.importBeta   DocumentMapper.java:475-536
  * updateVectorAlpha / updateScalarAlpha   VariationalInference.java:409-511, :573-625

and agree with the C oracle to <= 1e-13 relative (tests/test_oracle_mirror.py).  The reference
itself cannot run in this image (Java/Hadoop, no JVM), so this is an independent-restatement pin,
not a reference-output pin; DESIGN.md section 2 says so.

Third-party arithmetic (edu.umd:cloud9:1.4.17, not under /root/reference): Gamma.digamma is the
lda-c series that the reference's own known-answer tests pin (VariationalInferenceTest.java:27-62);
LogMath.add is taken as the exact log(exp a + exp b) (np.logaddexp); Gamma.lngamma as ln Gamma.

The NumPy path keeps the reference's fold ORDER: ufunc.reduce along an axis is a sequential left
fold, which is what the Java loops do (topics 0..K-1 for the normaliser, terms in document order for
updateLogGamma, documents in order for a (topic, term) cell, terms ascending for a topic).
"""
import math

import numpy as np

ETA = 1e-12  # Settings.java:58, DEFAULT_LOG_ETA = Math.log(1e-12)
COUNTER_SCALE = 1e6  # Settings.java:35


def digamma(x):
    """cloud9 Gamma.digamma (lda-c series), elementwise on float64 arrays, literal operation order."""
    x = np.asarray(x, dtype=np.float64) + 6.0
    p = 1.0 / (x * x)
    p = (((0.004166666666667 * p - 0.003968253986254) * p + 0.008333333333333) * p - 0.083333333333333) * p
    return (p + np.log(x) - 0.5 / x - 1.0 / (x - 1.0) - 1.0 / (x - 2.0) - 1.0 / (x - 3.0) - 1.0 / (x - 4.0)
            - 1.0 / (x - 5.0) - 1.0 / (x - 6.0))


def _lngamma(x):
    return math.lgamma(float(x))


def map_document(term_id, count, logbeta, alpha, gamma0=None, sweeps=99):
    """One call of DocumentMapper.map on a non-empty document.

    term_id (1-based) / count: the document's HMapII content in iteration order; logbeta: V x K,
    row v-1 = the double[K] of term v.  Returns (gamma[K], log_phi[n, K] of the last sweep =
    log(n_w phi_wk), document log likelihood, counter = (long)(-LL * 1e6))."""
    term_id = np.asarray(term_id)
    n_w = np.asarray(count, dtype=np.int64)
    K = len(alpha)
    alpha = np.asarray(alpha, dtype=np.float64)
    if gamma0 is not None:  # DocumentMapper.java:184-187
        temp_gamma = np.array(gamma0, dtype=np.float64)
    else:  # :189-192, float arithmetic: 1.0f * tokens / K, then widened and added to alpha
        share = np.float32(1.0) * np.float32(int(n_w.sum())) / np.float32(K)
        temp_gamma = alpha + np.float64(share)
    rows = logbeta[term_id - 1]  # n x K (retrieveBeta, :446-463, for terms present in the model)
    log_count = np.log(n_w.astype(np.float64))[:, None]
    fcount = n_w.astype(np.float64)[:, None]
    log_alpha = np.log(alpha)
    for _ in range(sweeps):  # :204-242: the counter starts at 1 and the loop runs while ++c < 100
        dg = digamma(temp_gamma)  # :209
        log_phi = rows + dg[None, :]  # updatePhi :408,413
        nf = np.logaddexp.reduce(log_phi, axis=1)  # :409-415, left fold over the topics
        log_phi = log_phi - nf[:, None]  # :420
        # :421, summed over the topics inside updatePhi, then over the terms in map (:225)
        conv = (fcount * np.exp(log_phi)) * (rows - log_phi)
        lik_phi = 0.0
        for w in range(conv.shape[0]):
            c = 0.0
            for v in conv[w]:
                c += v
            lik_phi += c
        log_phi = log_phi + log_count  # :422
        # :425 updateLogGamma[i] = LogMath.add(updateLogGamma[i], logPhi[i]), terms in order, seeded
        # with log(alpha[i]) (:210)
        ulg = np.logaddexp.reduce(np.vstack([log_alpha[None, :], log_phi]), axis=0)
        temp_gamma = np.exp(ulg)  # :228-230
    sum_gamma = 0.0
    lik_gamma = 0.0
    for g in temp_gamma:  # :245-251
        sum_gamma += g
        lik_gamma += _lngamma(g)
    lik_gamma -= _lngamma(sum_gamma)
    sum_alpha = 0.0
    sum_lng = 0.0
    for a in alpha:  # :121-126
        sum_lng += _lngamma(a)
        sum_alpha += a
    lik_alpha = _lngamma(sum_alpha) - sum_lng
    ll = lik_alpha + lik_gamma + lik_phi  # :252
    return temp_gamma, log_phi, ll, int(-ll * COUNTER_SCALE)  # Java (long) truncates toward zero


def map_phase(K, row_ptr, term_id, count, logbeta, alpha, gamma=None, sweeps=99):
    """The whole map phase over a CSR corpus: dict(gamma, log_phi, counter, doc_counter, alpha_ss).
    Documents with no content are counted and skipped (DocumentMapper.java:197-201)."""
    D = len(row_ptr) - 1
    out_gamma = np.zeros((D, K))
    log_phi = np.zeros((int(row_ptr[-1]), K))
    doc_counter = np.zeros(D, dtype=np.int64)
    ss = np.zeros(K)
    for d in range(D):
        b, e = int(row_ptr[d]), int(row_ptr[d + 1])
        if e <= b:
            if gamma is not None:
                out_gamma[d] = gamma[d]
            continue
        g, lp, _, c = map_document(term_id[b:e], count[b:e], logbeta, alpha,
                                   None if gamma is None else gamma[d], sweeps)
        out_gamma[d] = g
        log_phi[b:e] = lp
        doc_counter[d] = c
        sg = 0.0
        for x in g:
            sg += x
        ss += digamma(g) - digamma(np.float64(sg))  # :256-259
    return dict(gamma=out_gamma, log_phi=log_phi, counter=int(doc_counter.sum()),
                doc_counter=doc_counter, alpha_ss=ss)


def reduce_terms(K, V, row_ptr, term_id, log_phi, log_eta=None):
    """TermCombiner + TermReducer over the emitted (topic, term) -> log(n phi) records.

    Values of one key are folded in document order (TermCombiner.java:27-31, TermReducer.java:157-160),
    the prior is added (TermReducer.java:162-167), the beta file value is digamma(exp(.))
    (:195,213) and the topic key carries (float) digamma(exp(fold over the terms, ascending))
    (:173,189,212,233).  Returns (digamma_lambda V x K, normaliser float32[K], present uint8[V],
    lambda V x K)."""
    le = math.log(ETA) if log_eta is None else log_eta
    acc = np.full((V, K), -np.inf)
    present = np.zeros(V, dtype=np.uint8)
    for z in range(int(row_ptr[-1])):
        v = int(term_id[z]) - 1
        acc[v] = log_phi[z] if not present[v] else np.logaddexp(acc[v], log_phi[z])
        present[v] = 1
    seen = np.nonzero(present)[0]
    ln_lambda = np.logaddexp(le, acc[seen])  # LogMath.add(DEFAULT_LOG_ETA, logPhiValue)
    dl = np.zeros((V, K))
    lam = np.zeros((V, K))
    lam[seen] = np.exp(ln_lambda)
    dl[seen] = digamma(lam[seen])
    ln_norm = np.logaddexp.reduce(ln_lambda, axis=0)  # keys sorted by (topic, term): terms ascending
    nrm = digamma(np.exp(ln_norm)).astype(np.float32)
    return dl, nrm, present, lam


def import_beta(dl, nrm, present):
    """DocumentMapper.importBeta: value - (double) normaliserFloat; rows absent from the file stay 0."""
    out = dl - nrm.astype(np.float64)[None, :]
    out[present == 0] = 0.0
    return out


# ---------------------------------------------------------------------------------------------
# mpmath: the same formulas in 40-digit arithmetic (what the Java computes, minus its rounding)
# ---------------------------------------------------------------------------------------------
def map_document_mp(term_id, count, logbeta, alpha, sweeps=99, dps=40):
    import mpmath as mp

    mp.mp.dps = dps

    def dig(x):
        x = x + 6
        p = 1 / (x * x)
        p = (((mp.mpf("0.004166666666667") * p - mp.mpf("0.003968253986254")) * p
              + mp.mpf("0.008333333333333")) * p - mp.mpf("0.083333333333333")) * p
        return p + mp.log(x) - mp.mpf(1) / 2 / x - sum(1 / (x - j) for j in range(1, 7))

    K = len(alpha)
    al = [mp.mpf(float(a)) for a in alpha]
    n = [int(c) for c in count]
    share = float(np.float32(1.0) * np.float32(sum(n)) / np.float32(K))
    gam = [a + mp.mpf(share) for a in al]
    rows = [[mp.mpf(float(x)) for x in logbeta[int(t) - 1]] for t in term_id]
    lik_phi = mp.mpf(0)
    log_phi = None
    for _ in range(sweeps):
        dg = [dig(g) for g in gam]
        acc = list(al)  # exp(updateLogGamma), kept in linear space: exact arithmetic, same value
        lik_phi = mp.mpf(0)
        log_phi = []
        for w, row in enumerate(rows):
            x = [row[k] + dg[k] for k in range(K)]
            m = max(x)
            nf = m + mp.log(sum(mp.exp(v - m) for v in x))
            lp = [v - nf for v in x]
            lik_phi += sum(n[w] * mp.exp(lp[k]) * (row[k] - lp[k]) for k in range(K))
            lp = [v + mp.log(n[w]) for v in lp]
            for k in range(K):
                acc[k] += mp.exp(lp[k])
            log_phi.append(lp)
        gam = acc
    sg = sum(gam)
    ll = (mp.loggamma(sum(al)) - sum(mp.loggamma(a) for a in al)) + (sum(mp.loggamma(g) for g in gam)
                                                                      - mp.loggamma(sg)) + lik_phi
    return ([float(g) for g in gam], [[float(v) for v in r] for r in log_phi], float(ll))


# ---------------------------------------------------------------------------------------------
# Driver-side alpha update, written from VariationalInference.java:409-511 (updateVectorAlpha) and
# :573-625 (updateScalarAlpha) alone.  Java arrays are references: `alphaVectorUpdate =
# alphaVector` (:471) and `alphaVector = alphaVectorUpdate` (:500) make the two names one array, so
# Python lists (also references) restate the aliasing literally.  Constants are the float literals
# of Settings.java:60-63,68 widened to double, as Java widens them.
ALPHA_CONVERGE = float(np.float32(0.000001))  # DEFAULT_ALPHA_UPDATE_CONVERGE_THRESHOLD
ALPHA_MAX_ITER = 1000
ALPHA_MAX_DECAY = 10
ALPHA_DECAY = float(np.float32(0.8))  # DEFAULT_ALPHA_UPDATE_DECAY_FACTOR
ALPHA_SCALE = 10


def trigamma(x):
    """cloud9 Gamma.trigamma (lda-c series), scalar, literal operation order."""
    x = float(x) + 6.0
    p = 1.0 / (x * x)
    p = (((((0.075757575757576 * p - 0.033333333333333) * p + 0.0238095238095238) * p - 0.033333333333333) * p
          + 0.166666666666667) * p + 1.0) / x + 0.5 * p
    for _ in range(6):
        x = x - 1.0
        p = 1.0 / (x * x) + p
    return p


def _digamma1(x):
    return float(digamma(np.float64(x)))


def update_vector_alpha(n_topics, n_docs, alpha, ss):
    """VariationalInference.updateVectorAlpha as written, aliasing included."""
    f = np.float64
    alpha_vector = [f(a) for a in alpha]
    ss = [f(s) for s in ss]
    update = [f(0.0)] * n_topics
    grad = [f(0.0)] * n_topics
    hess = [f(0.0)] * n_topics
    count = 0
    keep_going = True
    decay = 0
    with np.errstate(all="ignore"):
        alpha_sum = f(0.0)
        for j in range(n_topics):
            alpha_sum = alpha_sum + alpha_vector[j]
        while keep_going:
            sum_g_h = f(0.0)
            sum_1_h = f(0.0)
            for i in range(n_topics):
                grad[i] = f(n_docs) * (f(_digamma1(alpha_sum)) - f(_digamma1(alpha_vector[i]))) + ss[i]
                hess[i] = f(-n_docs) * f(trigamma(alpha_vector[i]))
                if np.isinf(grad[i]):
                    return np.array(alpha_vector, dtype=np.float64)  # ArithmeticException, caught at :505
                sum_g_h = sum_g_h + grad[i] / hess[i]
                sum_1_h = sum_1_h + f(1.0) / hess[i]
            z = f(n_docs) * f(trigamma(alpha_sum))
            c = sum_g_h / (f(1.0) / z + sum_1_h)
            while True:
                singular = False
                for i in range(n_topics):
                    step = f(math.pow(ALPHA_DECAY, decay)) * (grad[i] - c) / hess[i]
                    if alpha_vector[i] <= step:
                        singular = True
                        break
                    update[i] = alpha_vector[i] - step
                if singular:
                    decay += 1
                    update = alpha_vector  # :471, the two names are one array from here on
                    if decay > ALPHA_MAX_DECAY:
                        break
                else:
                    break
            alpha_sum = f(0.0)
            keep_going = False
            for j in range(n_topics):
                alpha_sum = alpha_sum + update[j]
                if abs((update[j] - alpha_vector[j]) / alpha_vector[j]) >= ALPHA_CONVERGE:
                    keep_going = True
            if count >= ALPHA_MAX_ITER:
                keep_going = False
            if decay > ALPHA_MAX_DECAY:
                break
            count += 1
            alpha_vector = update  # :500
    return np.array(alpha_vector, dtype=np.float64)


def update_scalar_alpha(n_topics, n_docs, alpha_init, ss):
    """VariationalInference.updateScalarAlpha as written (log-space Newton on a symmetric alpha)."""
    f = np.float64
    alpha_init = f(alpha_init)
    ss = f(ss)
    count = 0
    update = alpha_init
    K = f(n_topics)
    with np.errstate(all="ignore"):
        while True:
            count += 1
            if np.isnan(update) or np.isinf(update):
                alpha_init = alpha_init * f(ALPHA_SCALE)
                update = alpha_init
            alpha_sum = update * K
            g = f(n_docs) * (K * f(_digamma1(alpha_sum)) - K * f(_digamma1(update))) + ss
            h = f(n_docs) * (K * K * f(trigamma(alpha_sum)) - K * f(trigamma(update)))
            update = np.exp(np.log(update) - g / (h * update + g))
            if abs(g) < ALPHA_CONVERGE:
                break
            if count > ALPHA_MAX_ITER:
                break
    return float(update)
