/*
 * mrlda_b200.h — C ABI of libmrlda_b200.so: Mr.LDA's variational-inference iteration on B200.
 *
 * This is the drop-in boundary for the one hot path of lintool/Mr.LDA: the work that
 * cc.mrlda.VariationalInference performs per EM iteration through JobClient.runJob
 * (VariationalInference.java:256) with DocumentMapper / TermCombiner / TermPartitioner /
 * TermReducer (VariationalInference.java:237-240) plus the driver-side alpha update
 * (VariationalInference.java:278-326).  A Java shim binds these entry points over JNI/Panama
 * (see INTEGRATION.md); the native driver host/mrlda_vi.cpp and the ctypes binding
 * mrlda_b200/capi.py call exactly the same symbols.
 *
 * Conventions
 *  - plain C types only; every pointer is caller-owned HOST memory borrowed for the duration of
 *    the call (the library copies to/from the device); no callbacks; no exceptions cross the ABI.
 *  - every function returns MRLDA_OK (0) or a negative MRLDA_E* code; mrlda_last_error() gives the
 *    text for the last failure on that context (or the last global failure when ctx is NULL).
 *  - term ids are 1-based as written by ParseCorpus (README.md:284, ParseCorpus.java:484-487);
 *    topic k of the API is the reference's topic key k+1.
 *  - K x V matrices cross the ABI term-major: element (term v, topic k) at [(v-1)*K + k], the
 *    layout of the reference's HMapIV<double[K]> keyed by term (DocumentMapper.java:43,475-536).
 *  - one context per process per GPU; calls on a context must be serialised by the caller
 *    (same rule as the reference's static mapper state, DocumentMapper.java:43-54).
 *  - there is no CPU fallback: without a usable CUDA device mrlda_create fails with MRLDA_ECUDA.
 */
#ifndef MRLDA_B200_H
#define MRLDA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRLDA_OK 0
#define MRLDA_EINVAL (-1)       /* bad argument (Preconditions.checkArgument in the reference) */
#define MRLDA_ECUDA (-2)        /* CUDA runtime error / no device */
#define MRLDA_ESTATE (-3)       /* call order: corpus / alpha / beta not set */
#define MRLDA_ENCCL (-4)        /* NCCL not loadable or a collective failed */
#define MRLDA_EUNSUPPORTED (-5) /* e.g. number of topics above the compiled kernel range */
#define MRLDA_ENOMEM (-6)

#define MRLDA_MAX_TOPICS 1024

typedef struct mrlda_ctx mrlda_ctx;

/* Mirrors the JobConf keys VariationalInference sets per iteration (VariationalInference.java:207-216)
 * and the constants of Settings.java. */
typedef struct mrlda_config {
  int32_t n_topics;        /* cc.mrlda.model.topics  (-topic)                                  */
  int32_t n_terms;         /* cc.mrlda.corpus.terms  (-term); term ids are 1..n_terms          */
  int32_t sweep_cap;       /* cc.mrlda.model.mapper.converge.iteration = 100 (Settings.java:54);
                              the mapper runs sweep_cap-1 = 99 sweeps (DocumentMapper.java:204-242) */
  int32_t learning;        /* cc.mrlda.model.train (false for -test)                           */
  int32_t random_start;    /* cc.mrlda.model.random.start (-randomstart): ignore stored gamma   */
  int32_t symmetric_alpha; /* -symmetricalpha: updateScalarAlpha instead of updateVectorAlpha   */
  int32_t device;          /* CUDA device ordinal                                              */
  int32_t rank;            /* document-shard rank (0 when single GPU)                          */
  int32_t world_size;      /* number of document shards / GPUs (1 when single GPU)             */
  int32_t update_alpha;    /* 1 = run the alpha update inside mrlda_iterate (default)          */
  int32_t fp32_mode;       /* 0 = fp64 (default, 1e-9 parity); 1 = optional fp32 E-step tile:
                              single-precision E / dot products / S, everything else fp64.
                              Stated tolerance against the fp64 reference on the same inputs:
                              ELBO 1e-6; logbeta 1e-4 from the third EM iteration on and 5e-3 in
                              the first two iterations from a random initial beta (near-identical
                              topics make the split of a document's mass ill-conditioned)        */
  double eta;              /* Settings.DEFAULT_LOG_ETA = log(1e-12) -> eta = 1e-12              */
  uint64_t seed;           /* seed for mrlda_init_beta_random (the reference uses Math.random)  */
  int32_t deterministic;   /* 1 = deterministic statistics mode (debugging aid): the phi-weighted
                              counts and the alpha statistics are accumulated in 192-bit fixed
                              point with integer atomics and all-reduced as integers, so beta,
                              alpha and the likelihood are bit-identical from run to run and for
                              any number of ranks / any sharding of the documents (the Hadoop
                              shuffle + TermReducer sum in key order, TermReducer.java:137-160).
                              Every document takes the general kernel: about 2.5x slower         */
  int32_t reserved;
} mrlda_config;

/* Counters and per-iteration outputs the driver reads after JobClient.runJob
 * (VariationalInference.java:260-275) plus device timings. */
typedef struct mrlda_iter_result {
  int64_t ll_counter;     /* LOG_LIKELIHOOD counter: sum_d (long)(-LL_d * 1e6) (DocumentMapper.java:253) */
  double log_likelihood;  /* -ll_counter / 1e6 (VariationalInference.java:261-262)             */
  int64_t n_docs;         /* TOTAL_DOCS over all shards                                         */
  int64_t n_terms_seen;   /* TOTAL_TERMS / K (VariationalInference.java:267); 0 when !learning  */
  int64_t n_tokens;       /* tokens processed by this E-step over all shards                    */
  double e_step_ms;       /* device time of the E-step kernel (CUDA events), this rank          */
  double m_step_ms;       /* device time of the beta finalize kernels                           */
  double alpha_ms;        /* device time of the alpha update kernel                             */
  double allreduce_ms;    /* device time of the NCCL exchange (0 when world_size == 1)          */
  int32_t estep_launches; /* kernels launched by this call, E-step part                         */
  int32_t total_launches; /* all kernels launched by this call                                  */
  int32_t slow_path_rows; /* (doc,term) rows that took the exact log-space path this iteration   */
  int32_t reserved;
} mrlda_iter_result;

/* Fills *cfg with the reference defaults (Settings.java): sweep_cap 100, learning 1, eta 1e-12,
 * random_start 0, symmetric_alpha 0, world_size 1, update_alpha 1. */
void mrlda_default_config(mrlda_config *cfg);

/* Replaces DocumentMapper.configure / TermReducer.configure (DocumentMapper.java:71-172). */
int mrlda_create(const mrlda_config *cfg, mrlda_ctx **out);
void mrlda_destroy(mrlda_ctx *ctx);
const char *mrlda_last_error(const mrlda_ctx *ctx);

/* The input split of this rank: D documents in CSR form, the content of the
 * SequenceFile<IntWritable, cc.mrlda.Document> records (Document.java:147-172).
 *   row_ptr[D+1]; term_id[Z] 1-based, count[Z] > 0; doc_id[D] (the IntWritable keys, may be NULL);
 *   gamma0[D*K] stored gamma (Document.gamma) or NULL for none (DocumentMapper.java:184-193).
 * A document with row_ptr[d+1]==row_ptr[d] is the reference's "content == null" case: counted in
 * TOTAL_DOCS and skipped (DocumentMapper.java:178,197-201). */
int mrlda_set_corpus_csr(mrlda_ctx *ctx, int64_t n_docs, const int64_t *row_ptr,
                         const int32_t *term_id, const int32_t *count, const int32_t *doc_id,
                         const double *gamma0);

/* importAlpha (VariationalInference.java:521-540): K values, all must be > 0. */
int mrlda_set_alpha(mrlda_ctx *ctx, const double *alpha);
int mrlda_get_alpha(mrlda_ctx *ctx, double *alpha);

/* importBeta (DocumentMapper.java:475-536): digamma_lambda[V*K] term-major are the HMapIDW values
 * of the beta file, normaliser[K] the PairOfIntFloat right elements, present[V] says which terms
 * have a row in the file.  logbeta = value - (double)normaliser; absent rows are initialised as in
 * retrieveBeta (DocumentMapper.java:446-463) but from cfg.seed instead of Math.random. */
int mrlda_set_beta(mrlda_ctx *ctx, const double *digamma_lambda, const float *normaliser,
                   const uint8_t *present);
/* The content of the next beta file (TermReducer output, TermReducer.java:173-195,232-235). */
int mrlda_get_beta(mrlda_ctx *ctx, double *digamma_lambda, float *normaliser, uint8_t *present);

/* -informedprior (InformedPrior.java:43-44,172-177; hook TermReducer.java:162-164): n (topic, term)
 * pairs, both 1-based, name the seed terms of each topic.  With a prior set, the finalize uses
 * log eta = (float)log(1000) for seed cells and (float)log(0.001) for all others instead of
 * log(1e-12).  n = 0 removes the prior.  Deviation: the reference looks the seed set up under the
 * previously reduced key's topic for the first term of each topic (TermReducer.java:164 runs before
 * :188, and "previous" depends on -reducer); the topic itself is used here. */
int mrlda_set_informed_prior(mrlda_ctx *ctx, int64_t n, const int32_t *topic, const int32_t *term);

/* Direct injection / read-back of the K x V log-beta the mapper works with (term-major). */
int mrlda_set_logbeta(mrlda_ctx *ctx, const double *logbeta);
int mrlda_get_logbeta(mrlda_ctx *ctx, double *logbeta);
/* Iteration-0 initialisation in the reference's form log(2*U1/V + U2) (DocumentMapper.java:456),
 * U from a counter-based generator seeded with cfg.seed (deterministic across ranks). */
int mrlda_init_beta_random(mrlda_ctx *ctx);

/* One EM iteration = one JobClient.runJob (VariationalInference.java:256) + the alpha update
 * (:278-326): E-step over the resident documents, exchange of the K x V statistics across ranks,
 * beta finalize (TermReducer), alpha Newton update.  With learning == 0 (-test) only the E-step and
 * the likelihood run (DocumentMapper.java:264,316,353).  The updated gamma stays resident and is the
 * next iteration's starting point, as the gamma-N directory is in the reference (VI.java:359-379). */
int mrlda_iterate(mrlda_ctx *ctx, mrlda_iter_result *out);

/* The two halves of mrlda_iterate, for callers that time or interleave them.  mrlda_e_step leaves
 * the (all-reduced) statistics on the device; mrlda_m_step finalises beta and alpha from them. */
int mrlda_e_step(mrlda_ctx *ctx, mrlda_iter_result *out);
int mrlda_m_step(mrlda_ctx *ctx, mrlda_iter_result *out);

/* gamma[D*K] of this rank's documents, document order as given to mrlda_set_corpus_csr
 * (the Document.gamma written to the "gamma" named output, DocumentMapper.java:342-346). */
int mrlda_get_gamma(mrlda_ctx *ctx, double *gamma);
/* Alpha sufficient statistics of the last E-step, the content of tempDir/part-00000
 * (DocumentMapper.java:354-361, VariationalInference.java:148,288-289). */
int mrlda_get_alpha_ss(mrlda_ctx *ctx, double *alpha_ss);
/* Linear-space K x V statistics sum_d n_dw*phi_dwk of the last E-step (term-major). */
int mrlda_get_stats(mrlda_ctx *ctx, double *stats);

/* updateVectorAlpha / updateScalarAlpha as stand-alone device calls
 * (VariationalInference.java:409-511, 573-625) — same arguments as the Java statics. */
int mrlda_update_vector_alpha(mrlda_ctx *ctx, int32_t n_topics, int32_t n_docs,
                              const double *alpha_in, const double *alpha_ss, double *alpha_out);
int mrlda_update_scalar_alpha(mrlda_ctx *ctx, int32_t n_topics, int32_t n_docs, double alpha_init,
                              double alpha_ss_total, double *alpha_out);
/* cloud9 Gamma.digamma / trigamma / lngamma evaluated by the device functions the kernels use. */
int mrlda_eval_special(mrlda_ctx *ctx, int32_t which /*0 digamma,1 trigamma,2 lngamma,3 E-step exp(digamma)*/,
                       int64_t n, const double *x, double *y);

/* Multi-GPU (one process per GPU, documents sharded across ranks; replaces Hadoop's shuffle with one
 * NCCL all-reduce over NVLink).  Rank 0 calls mrlda_nccl_unique_id, the host side ships the 128
 * bytes to every rank (any transport), then every rank calls mrlda_comm_init. */
#define MRLDA_NCCL_ID_BYTES 128
int mrlda_nccl_unique_id(uint8_t id[MRLDA_NCCL_ID_BYTES]);
int mrlda_comm_init(mrlda_ctx *ctx, const uint8_t id[MRLDA_NCCL_ID_BYTES]);

/* Measured DFMA throughput of the device (TFLOP/s), the second roofline denominator bench.py
 * reports: at the reference's fixed 99 sweeps the E-step is FP64-pipe bound. */
int mrlda_measure_fp64_peak(mrlda_ctx *ctx, double *tflops);

/* Profiling aid: per-phase cycle counters of the register-stationary E-step kernel, accumulated by
 * one thread; all zero unless the library was built with -DMRLDA_PHASE_TIMING.  Reading resets. */
int mrlda_debug_phase_cycles(mrlda_ctx *ctx, int64_t out[16]);

/* Number of this library's kernels launched on the context since creation. */
int64_t mrlda_kernel_launches(const mrlda_ctx *ctx);
/* Which E-step kernels a context with this many topics uses (fp64 mode), longest documents last, as
 * one line of text: kernel template instance, cluster size and the largest document (distinct
 * terms) each holds on a B200.  Pure host code, no CUDA call: answers without a GPU.  Returns the
 * length written (excluding the terminator), or MRLDA_EINVAL / MRLDA_EUNSUPPORTED. */
int mrlda_describe_kernels(int n_topics, char *buf, size_t len);
/* Library/build description, e.g. "mrlda_b200 0.1 sm_100a". */
const char *mrlda_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MRLDA_B200_H */
