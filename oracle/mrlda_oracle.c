/*
 * mrlda_oracle.c — CPU restatement of Mr.LDA's variational-inference iteration.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker or the timed CPU arm.  The product path (mrlda_b200/) never calls it.
 *
 * What is restated (reference = /root/reference/src/main/java/cc/mrlda, "DM" = DocumentMapper.java,
 * "TR" = TermReducer.java, "TC" = TermCombiner.java, "VI" = VariationalInference.java,
 * "Set" = Settings.java):
 *   - DM.java:184-259   per-document E-step (gamma init, 99 log-space sweeps, ELBO, alpha stats)
 *   - DM.java:402-429   updatePhi (log-space phi normalisation, LogMath.add fold)
 *   - DM.java:315-329   phi emission (the -directemit semantics: every (doc,term,topic) is summed)
 *   - TC.java:19-35 / TR.java:134-238  log-space reduce, + eta, digamma, (float) topic normaliser
 *   - DM.java:475-536   importBeta  (logbeta = value - (double)normaliserFloat)
 *   - VI.java:409-511   updateVectorAlpha (with the Java array aliasing)
 *   - VI.java:573-625   updateScalarAlpha; VI.java:295-307 symmetric wrapper
 *   - VI.java:260-268,381-387  likelihood counter decode and convergence test
 *
 * Third-party arithmetic that is NOT under /root/reference: edu.umd:cloud9:1.4.17
 * (pom.xml:123-127), classes edu.umd.cloud9.math.Gamma and edu.umd.cloud9.math.LogMath.
 *   - digamma / trigamma: lda-c style series.  PINNED by the reference's own known-answer tests
 *     (src/test/java/cc/mrlda/VariationalInferenceTest.java:27-62, all nine reproduced to 1e-10;
 *     see tests/test_oracle_kat.py).
 *   - lngamma and LogMath.add: PARITY UNPINNED — no reference test or golden vector touches them.
 *     This file uses C lgamma() (<= 1e-15 relative) and the exact two-operand log-sum
 *     max + log1p(exp(min-max)) without a drop threshold.  A "cutoff" mode
 *     (mrlda_oracle_set_logadd_cutoff) restates the recalled cloud9 behaviour (drop addends more than
 *     T nats below the running sum) so the gap can be measured; it is not the parity target.
 *   - HMapII iteration order (DM.java:213) is unpinned; the oracle walks terms in CSR order.
 *
 * Layouts: term ids are 1-based as in the reference (README.md:284); logbeta is term-major,
 * row v-1 holds K contiguous doubles (this mirrors HMapIV<double[K]> keyed by term, DM.java:43).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MRLDA_SWEEP_CAP 100          /* Set.java:54 MAXIMUM_LOCAL_ITERATION */
#define MRLDA_COUNTER_SCALE 1e6      /* Set.java:35 */
#define MRLDA_ETA 1e-12              /* Set.java:58 DEFAULT_LOG_ETA = log(1e-12) */

static double g_logadd_cutoff = -1.0; /* < 0: exact; >= 0: drop addends more than cutoff below */

void mrlda_oracle_set_logadd_cutoff(double nats) { g_logadd_cutoff = nats; }

int mrlda_oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* Sets the OpenMP team size explicitly (bench.py passes the number of cores the process may run
 * on): torchrun exports OMP_NUM_THREADS=1, which would otherwise turn the timed CPU arm into a
 * single-thread run.  Returns the team size now in effect. */
int mrlda_oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) {
    omp_set_dynamic(0);
    omp_set_num_threads(n);
  }
  return omp_get_max_threads();
#else
  (void)n;
  return 1;
#endif
}

/* cloud9 Gamma.digamma as pinned by VariationalInferenceTest (lda-c series). */
#pragma GCC push_options
#pragma GCC optimize("fp-contract=off")
double mrlda_oracle_digamma(double x) {
  double p;
  x = x + 6;
  p = 1 / (x * x);
  p = (((0.004166666666667 * p - 0.003968253986254) * p + 0.008333333333333) * p -
       0.083333333333333) * p;
  p = p + log(x) - 0.5 / x - 1 / (x - 1) - 1 / (x - 2) - 1 / (x - 3) - 1 / (x - 4) - 1 / (x - 5) -
      1 / (x - 6);
  return p;
}

/* cloud9 Gamma.trigamma as pinned by VariationalInferenceTest (lda-c series). */
double mrlda_oracle_trigamma(double x) {
  double p;
  int i;
  x = x + 6;
  p = 1 / (x * x);
  p = (((((0.075757575757576 * p - 0.033333333333333) * p + 0.0238095238095238) * p -
         0.033333333333333) * p + 0.166666666666667) * p + 1) / x + 0.5 * p;
  for (i = 0; i < 6; i++) {
    x = x - 1;
    p = 1 / (x * x) + p;
  }
  return p;
}
#pragma GCC pop_options

/* cloud9 Gamma.lngamma — unpinned, see header. */
double mrlda_oracle_lngamma(double x) { return lgamma(x); }

/* cloud9 LogMath.add — unpinned, see header. */
double mrlda_oracle_logadd(double a, double b) {
  double hi = a > b ? a : b, lo = a > b ? b : a;
  if (hi == -INFINITY) return hi;
  double d = lo - hi;
  if (g_logadd_cutoff >= 0 && d < -g_logadd_cutoff) return hi;
  return hi + log1p(exp(d));
}

/* DM.java:402-429 updatePhi, verbatim structure. Returns the phi part of the doc likelihood. */
static double update_phi(int K, int term_count, const double *log_beta, const double *digamma_gamma,
                         double *log_phi, double *update_log_gamma) {
  double converge_phi = 0;
  log_phi[0] = log_beta[0] + digamma_gamma[0];
  double normalize_factor = log_phi[0];
  for (int i = 1; i < K; i++) {
    log_phi[i] = log_beta[i] + digamma_gamma[i];
    normalize_factor = mrlda_oracle_logadd(normalize_factor, log_phi[i]);
  }
  double log_count = log((double)term_count);
  for (int i = 0; i < K; i++) {
    log_phi[i] -= normalize_factor;
    converge_phi += term_count * exp(log_phi[i]) * (log_beta[i] - log_phi[i]);
    log_phi[i] += log_count;
    update_log_gamma[i] = mrlda_oracle_logadd(update_log_gamma[i], log_phi[i]);
  }
  return converge_phi;
}

/* DM.java:121-126 likelihoodAlpha = lngamma(sum alpha) - sum lngamma(alpha_k). */
double mrlda_oracle_likelihood_alpha(int K, const double *alpha) {
  double s = 0, l = 0;
  for (int k = 0; k < K; k++) {
    l += mrlda_oracle_lngamma(alpha[k]);
    s += alpha[k];
  }
  return mrlda_oracle_lngamma(s) - l;
}

/*
 * One document through DocumentMapper.map (DM.java:175-259).
 *   gamma[K]      in: stored gamma (used when has_gamma && !random_start), out: updated gamma
 *   log_phi       out: n_terms x K values log(n_w * phi_wk) of the LAST sweep (DM.java:321-327)
 *   alpha_ss[K]   += digamma(gamma_k) - digamma(sum gamma)      (DM.java:256-259)
 * Returns the document log likelihood (DM.java:252); *counter_out gets (long)(-LL * 1e6).
 */
double mrlda_oracle_estep_doc(int K, int sweeps, int n_terms, const int32_t *term_id,
                              const int32_t *count, const double *logbeta, const double *alpha,
                              double likelihood_alpha, int has_gamma, int random_start,
                              double *gamma, double *log_phi, double *alpha_ss,
                              int64_t *counter_out) {
  double *update_log_gamma = (double *)malloc(sizeof(double) * (size_t)K);
  double likelihood_phi = 0;
  long n_tokens = 0;
  for (int w = 0; w < n_terms; w++) n_tokens += count[w];

  if (!(has_gamma && !random_start)) {
    /* DM.java:191: alpha[i] + 1.0f * numberOfTokens / numberOfTopics  (float division) */
    float share = 1.0f * (float)n_tokens / (float)K;
    for (int i = 0; i < K; i++) gamma[i] = alpha[i] + share;
  }

  /* DM.java:204,236,242: count starts at 1, loop while ++count < cap  => cap-1 sweeps. */
  for (int it = 0; it < sweeps; it++) {
    likelihood_phi = 0;
    for (int i = 0; i < K; i++) {
      gamma[i] = mrlda_oracle_digamma(gamma[i]);
      update_log_gamma[i] = log(alpha[i]);
    }
    for (int w = 0; w < n_terms; w++) {
      const double *lb = logbeta + (size_t)(term_id[w] - 1) * (size_t)K;
      likelihood_phi += update_phi(K, count[w], lb, gamma, log_phi + (size_t)w * K,
                                   update_log_gamma);
    }
    for (int i = 0; i < K; i++) gamma[i] = exp(update_log_gamma[i]);
  }

  double sum_gamma = 0, likelihood_gamma = 0;
  for (int i = 0; i < K; i++) {
    sum_gamma += gamma[i];
    likelihood_gamma += mrlda_oracle_lngamma(gamma[i]);
  }
  likelihood_gamma -= mrlda_oracle_lngamma(sum_gamma);
  double ll = likelihood_alpha + likelihood_gamma + likelihood_phi;
  if (counter_out) *counter_out = (int64_t)(-ll * MRLDA_COUNTER_SCALE); /* DM.java:253-254 */

  if (alpha_ss) {
    double dsg = mrlda_oracle_digamma(sum_gamma);
    for (int i = 0; i < K; i++) alpha_ss[i] += mrlda_oracle_digamma(gamma[i]) - dsg;
  }
  free(update_log_gamma);
  return ll;
}

/*
 * E-step over a CSR corpus (the whole map phase).  Outputs:
 *   gamma[D*K]      in/out (in only read when has_gamma)
 *   log_phi[Z*K]    optional (NULL to skip): per-nnz log(n*phi) of the last sweep
 *   alpha_ss[K]     zeroed then accumulated in document order per thread, threads summed in order
 *   doc_counter[D]  optional per-doc int64 counters; return value = their sum (LOG_LIKELIHOOD)
 * OpenMP over documents (static schedule): the CPU baseline leg of bench.py times this call.
 */
int64_t mrlda_oracle_estep(int K, int sweeps, int64_t D, const int64_t *row_ptr,
                           const int32_t *term_id, const int32_t *count, const double *logbeta,
                           const double *alpha, int has_gamma, int random_start, double *gamma,
                           double *log_phi, double *alpha_ss, int64_t *doc_counter) {
  double la = mrlda_oracle_likelihood_alpha(K, alpha);
  int64_t total = 0;
  int nthreads = mrlda_oracle_num_threads();
  double *ss_all = (double *)calloc((size_t)nthreads * (size_t)K, sizeof(double));
#pragma omp parallel reduction(+ : total)
  {
#ifdef _OPENMP
    int tid = omp_get_thread_num();
#else
    int tid = 0;
#endif
    double *ss = ss_all + (size_t)tid * K;
    double *scratch = NULL;
#pragma omp for schedule(dynamic, 4)
    for (int64_t d = 0; d < D; d++) {
      int64_t b = row_ptr[d], e = row_ptr[d + 1];
      int n = (int)(e - b);
      int64_t c = 0;
      if (n <= 0) {
        /* DM.java:178,197-201: null content => doc counted, skipped before any update. */
        if (doc_counter) doc_counter[d] = 0;
        continue;
      }
      double *lp = log_phi ? log_phi + (size_t)b * K : NULL;
      if (!lp) {
        scratch = (double *)realloc(scratch, sizeof(double) * (size_t)n * K);
        lp = scratch;
      }
      mrlda_oracle_estep_doc(K, sweeps, n, term_id + b, count + b, logbeta, alpha, la, has_gamma,
                             random_start, gamma + (size_t)d * K, lp, ss, &c);
      if (doc_counter) doc_counter[d] = c;
      total += c;
    }
    free(scratch);
  }
  if (alpha_ss) {
    for (int k = 0; k < K; k++) {
      double s = 0;
      for (int t = 0; t < nthreads; t++) s += ss_all[(size_t)t * K + k];
      alpha_ss[k] = s;
    }
  }
  free(ss_all);
  return total;
}

/*
 * M-step for beta: TermCombiner + TermReducer (TC.java:19-35, TR.java:134-238).
 * For each topic k (1-based key k+1) and each term v that occurs in the corpus, ascending v:
 *   s      = LogMath.add-fold of log(n*phi) over the nnz of term v, in document order
 *   lnlam  = LogMath.add(log eta, s)                                    (TR.java:166)
 *   out[v][k] = digamma(exp(lnlam))                                     (TR.java:195,213)
 *   lnnorm_k  = LogMath.add-fold of lnlam over v ascending              (TR.java:189,212)
 *   normaliser[k] = (float) digamma(exp(lnnorm_k))                      (TR.java:173,233)
 * present[v-1] = 1 for terms that were emitted. Returns TOTAL_TERMS / K = #seen terms (VI.java:267).
 * digamma_lambda is term-major V x K; rows of unseen terms are left untouched.
 */
int64_t mrlda_oracle_mstep(int K, int V, int64_t D, const int64_t *row_ptr, const int32_t *term_id,
                           const double *log_phi, double *digamma_lambda, float *normaliser,
                           uint8_t *present, const uint8_t *seed) {
  int64_t Z = row_ptr[D];
  size_t cells = (size_t)V * (size_t)K;
  double *acc = (double *)malloc(sizeof(double) * cells);
  memset(present, 0, (size_t)V);
  for (int64_t z = 0; z < Z; z++) {
    int v = term_id[z] - 1;
    const double *lp = log_phi + (size_t)z * K;
    double *a = acc + (size_t)v * K;
    if (!present[v]) {
      present[v] = 1;
      for (int k = 0; k < K; k++) a[k] = lp[k];
    } else {
      for (int k = 0; k < K; k++) a[k] = mrlda_oracle_logadd(a[k], lp[k]);
    }
  }
  double log_eta = log(MRLDA_ETA);
  /* -informedprior (InformedPrior.java:43-44,172-177; hook TermReducer.java:162-164): seed[v*K+k]
   * says term v+1 is a seed of topic k+1; log eta is a FLOAT, log 1000 for seeds, log 0.001 else.
   * Deviation: the reference looks the seed set up under the previously reduced key's topic for the
   * first term of every topic (TermReducer.java:164 runs before :188); the topic itself is used. */
  const double log_eta_hi = (double)(float)log(1000.0), log_eta_lo = (double)(float)log(0.001);
  int64_t seen = 0;
  for (int v = 0; v < V; v++) seen += present[v] ? 1 : 0;
#pragma omp parallel for schedule(static)
  for (int k = 0; k < K; k++) {
    int first = 1;
    double lnnorm = 0;
    for (int v = 0; v < V; v++) {
      if (!present[v]) continue;
      double le = seed ? (seed[(size_t)v * K + k] ? log_eta_hi : log_eta_lo) : log_eta;
      double lnlam = mrlda_oracle_logadd(le, acc[(size_t)v * K + k]);
      digamma_lambda[(size_t)v * K + k] = mrlda_oracle_digamma(exp(lnlam));
      if (first) {
        lnnorm = lnlam;
        first = 0;
      } else {
        lnnorm = mrlda_oracle_logadd(lnnorm, lnlam);
      }
    }
    normaliser[k] = first ? 0.0f : (float)mrlda_oracle_digamma(exp(lnnorm));
  }
  free(acc);
  return seen;
}

/* DM.java:475-536 importBeta: logbeta[v][k] = value - (double)normaliserFloat; absent rows stay 0. */
void mrlda_oracle_import_beta(int K, int V, const double *digamma_lambda, const float *normaliser,
                              const uint8_t *present, double *logbeta) {
  for (int v = 0; v < V; v++) {
    for (int k = 0; k < K; k++) {
      size_t i = (size_t)v * K + k;
      logbeta[i] = present[v] ? digamma_lambda[i] - (double)normaliser[k] : 0.0;
    }
  }
}

/*
 * VI.java:409-511 updateVectorAlpha, restated with the Java reference semantics: alpha_vector
 * and alpha_update are *references*; `alphaVectorUpdate = alphaVector` (VI.java:471) and
 * `alphaVector = alphaVectorUpdate` (VI.java:500) alias them.  out[K] receives the returned vector.
 */
void mrlda_oracle_update_vector_alpha(int K, int n_docs, const double *alpha_in, const double *ss,
                                      double *out) {
  double *buf_a = (double *)malloc(sizeof(double) * (size_t)K);
  double *buf_b = (double *)calloc((size_t)K, sizeof(double));
  double *grad = (double *)malloc(sizeof(double) * (size_t)K);
  double *hess = (double *)malloc(sizeof(double) * (size_t)K);
  memcpy(buf_a, alpha_in, sizeof(double) * (size_t)K);
  double *alpha_vector = buf_a, *alpha_update = buf_b;
  int iter = 0, keep_going = 1, decay = 0;
  double alpha_sum = 0;
  for (int j = 0; j < K; j++) alpha_sum += alpha_vector[j];

  while (keep_going) {
    double sum_g_h = 0, sum_1_h = 0;
    int invalid = 0;
    for (int i = 0; i < K; i++) {
      grad[i] = n_docs * (mrlda_oracle_digamma(alpha_sum) - mrlda_oracle_digamma(alpha_vector[i])) +
                ss[i];
      hess[i] = -n_docs * mrlda_oracle_trigamma(alpha_vector[i]);
      if (isinf(grad[i])) { /* VI.java:440-443 throws ArithmeticException -> caught, return */
        invalid = 1;
        break;
      }
      sum_g_h += grad[i] / hess[i];
      sum_1_h += 1 / hess[i];
    }
    if (invalid) break;
    double z = n_docs * mrlda_oracle_trigamma(alpha_sum);
    double c = sum_g_h / (1 / z + sum_1_h);

    for (;;) {
      int singular = 0;
      /* Set.java:63: 0.8f promoted to double by Math.pow */
      double scale = pow((double)0.8f, (double)decay);
      for (int i = 0; i < K; i++) {
        double step = scale * (grad[i] - c) / hess[i];
        if (alpha_vector[i] <= step) {
          singular = 1;
          break;
        }
        alpha_update[i] = alpha_vector[i] - step;
      }
      if (singular) {
        decay++;
        alpha_update = alpha_vector; /* VI.java:471 (aliasing) */
        if (decay > 10) break;       /* Set.java:62 */
      } else {
        break;
      }
    }

    alpha_sum = 0;
    keep_going = 0;
    for (int j = 0; j < K; j++) {
      alpha_sum += alpha_update[j];
      /* Set.java:60: threshold 0.000001f promoted to double */
      if (fabs((alpha_update[j] - alpha_vector[j]) / alpha_vector[j]) >= (double)0.000001f)
        keep_going = 1;
    }
    if (iter >= 1000) keep_going = 0; /* Set.java:61 */
    if (decay > 10) break;
    iter++;
    alpha_vector = alpha_update; /* VI.java:500 (aliasing) */
  }
  memcpy(out, alpha_vector, sizeof(double) * (size_t)K);
  free(buf_a);
  free(buf_b);
  free(grad);
  free(hess);
}

/* VI.java:573-625 updateScalarAlpha. */
double mrlda_oracle_update_scalar_alpha(int K, int n_docs, double alpha_init, double ss) {
  int iter = 0;
  double alpha_update = alpha_init;
  for (;;) {
    iter++;
    if (isnan(alpha_update) || isinf(alpha_update)) {
      alpha_init *= 10; /* Set.java:68 */
      alpha_update = alpha_init;
    }
    double alpha_sum = alpha_update * K;
    double g = n_docs * (K * mrlda_oracle_digamma(alpha_sum) - K * mrlda_oracle_digamma(alpha_update)) +
               ss;
    double h = n_docs * ((double)(K * K) * mrlda_oracle_trigamma(alpha_sum) -
                         K * mrlda_oracle_trigamma(alpha_update));
    alpha_update = exp(log(alpha_update) - g / (h * alpha_update + g));
    if (fabs(g) < (double)0.000001f) break;
    if (iter > 1000) break;
  }
  return alpha_update;
}

/* VI.java:278-311: the driver-side alpha update (vector, or symmetric wrapper VI.java:295-307). */
void mrlda_oracle_update_alpha(int K, int n_docs, int symmetric, const double *alpha_in,
                               const double *ss, double *out) {
  if (symmetric) {
    double tot = 0, old = 0;
    for (int i = 0; i < K; i++) {
      tot += ss[i];
      old += alpha_in[i];
    }
    old /= K;
    double a = mrlda_oracle_update_scalar_alpha(K, n_docs, old, tot);
    for (int i = 0; i < K; i++) out[i] = a;
  } else {
    mrlda_oracle_update_vector_alpha(K, n_docs, alpha_in, ss, out);
  }
}

/*
 * One full EM iteration in the intended (cluster) semantics, SURVEY.md section 3.6.
 * In: logbeta (V x K), alpha[K], gamma (D x K, read iff has_gamma).  Out: new digamma_lambda,
 * normaliser, present, new alpha, gamma, alpha_ss; returns the LOG_LIKELIHOOD counter (int64);
 * decode as -counter/1e6 (VI.java:261-262).  n_docs counts every document, empty ones included
 * (DM.java:178 increments TOTAL_DOCS before the null-content return).
 */
int64_t mrlda_oracle_iterate(int K, int V, int sweeps, int64_t D, const int64_t *row_ptr,
                             const int32_t *term_id, const int32_t *count, const double *logbeta,
                             const double *alpha, int has_gamma, int random_start, int learning,
                             int symmetric, double *gamma, double *digamma_lambda,
                             float *normaliser, uint8_t *present, double *alpha_out,
                             double *alpha_ss, int64_t *n_terms_seen, const uint8_t *seed) {
  int64_t Z = row_ptr[D];
  double *log_phi = learning ? (double *)malloc(sizeof(double) * (size_t)Z * K) : NULL;
  int64_t counter = mrlda_oracle_estep(K, sweeps, D, row_ptr, term_id, count, logbeta, alpha,
                                       has_gamma, random_start, gamma, log_phi, alpha_ss, NULL);
  if (learning) {
    int64_t seen =
        mrlda_oracle_mstep(K, V, D, row_ptr, term_id, log_phi, digamma_lambda, normaliser, present, seed);
    if (n_terms_seen) *n_terms_seen = seen;
    mrlda_oracle_update_alpha(K, (int)D, symmetric, alpha, alpha_ss, alpha_out);
    free(log_phi);
  }
  return counter;
}

/* VI.java:383 convergence test. */
int mrlda_oracle_converged(double last_ll, double ll) {
  return fabs((last_ll - ll) / last_ll) <= 0.000001; /* Set.java:56 */
}
