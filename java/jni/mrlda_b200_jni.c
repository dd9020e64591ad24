/*
 * mrlda_b200_jni.c — JNI glue between cc.mrlda.b200.MrLdaB200 and the C ABI of libmrlda_b200.so
 * (include/mrlda_b200.h).  One native method = one ABI call; Java arrays are pinned with
 * GetPrimitiveArrayCritical for the duration of the call (the library copies to / from the device,
 * nothing outlives the call) and every failure becomes a java.io.IOException carrying
 * mrlda_last_error().  The C ABI takes plain pointers and trusts their extents, so the glue checks
 * every array length against K, V, D and Z BEFORE the call: a short Java array is an IOException,
 * never an out-of-bounds access.  No JNI function is called while an array is pinned (the JNI
 * rule for critical regions): failures found inside one are thrown after the last release.
 *
 * Build (needs a JDK, which this repository's image does not have):
 *   gcc -shared -fPIC -I"$JAVA_HOME/include" -I"$JAVA_HOME/include/linux" -Iinclude \
 *       java/jni/mrlda_b200_jni.c -Lmrlda_b200 -lmrlda_b200 -o libmrlda_b200_jni.so
 * The CPU test suite syntax-checks this file against a minimal jni.h (tests/jni_stub/).
 */
#include <jni.h>
#include <stdint.h>
#include <string.h>

#include "mrlda_b200.h"

static void throw_io(JNIEnv *env, const char *msg) {
  jclass cls = (*env)->FindClass(env, "java/io/IOException");
  if (cls) (*env)->ThrowNew(env, cls, msg ? msg : "mrlda_b200: unknown error");
}

static mrlda_ctx *ctx_of(JNIEnv *env, jobject self) {
  jclass cls = (*env)->GetObjectClass(env, self);
  jfieldID fid = (*env)->GetFieldID(env, cls, "handle", "J");
  if (!fid) return NULL;
  mrlda_ctx *ctx = (mrlda_ctx *)(intptr_t)(*env)->GetLongField(env, self, fid);
  if (!ctx) throw_io(env, "mrlda_b200: context is closed");
  return ctx;
}

static int check(JNIEnv *env, mrlda_ctx *ctx, int rc) {
  if (rc == MRLDA_OK) return 0;
  throw_io(env, mrlda_last_error(ctx));
  return 1;
}

static const char *const kPinFailed = "mrlda_b200: could not pin a Java array";

/* pins a (possibly null) primitive array; returns 0 on success.  Does not throw: other arrays may
 * be pinned already, and a critical region allows no other JNI call. */
static int pin(JNIEnv *env, jarray arr, void **out) {
  *out = NULL;
  if (!arr) return 0;
  *out = (*env)->GetPrimitiveArrayCritical(env, arr, NULL);
  return *out ? 0 : 1;
}

/* arr must be non-null (unless optional) and hold exactly `need` elements; throws otherwise */
static int bad_length(JNIEnv *env, jarray arr, int64_t need, int optional, const char *what) {
  if (!arr) {
    if (optional) return 0;
    throw_io(env, what);
    return 1;
  }
  if ((int64_t)(*env)->GetArrayLength(env, arr) != need) {
    throw_io(env, what);
    return 1;
  }
  return 0;
}

static jint int_field(JNIEnv *env, jobject self, const char *name) {
  jclass cls = (*env)->GetObjectClass(env, self);
  jfieldID fid = (*env)->GetFieldID(env, cls, name, "I");
  return fid ? (*env)->GetIntField(env, self, fid) : 0;
}
/* K and V of the Java object (final int fields "topics" and "terms") */
static jint topics_of(JNIEnv *env, jobject self) { return int_field(env, self, "topics"); }
static jint terms_of(JNIEnv *env, jobject self) { return int_field(env, self, "terms"); }

/* D of the corpus the last successful setCorpus loaded (long field "docs") */
static int64_t docs_of(JNIEnv *env, jobject self) {
  jclass cls = (*env)->GetObjectClass(env, self);
  jfieldID fid = (*env)->GetFieldID(env, cls, "docs", "J");
  return fid ? (int64_t)(*env)->GetLongField(env, self, fid) : 0;
}
static void set_docs(JNIEnv *env, jobject self, int64_t d) {
  jclass cls = (*env)->GetObjectClass(env, self);
  jfieldID fid = (*env)->GetFieldID(env, cls, "docs", "J");
  if (fid) (*env)->SetLongField(env, self, fid, (jlong)d);
}
static void unpin(JNIEnv *env, jarray arr, void *p, jint mode) {
  if (arr && p) (*env)->ReleasePrimitiveArrayCritical(env, arr, p, mode);
}

JNIEXPORT jlong JNICALL Java_cc_mrlda_b200_MrLdaB200_create(JNIEnv *env, jclass cls, jint topics, jint terms,
                                                            jboolean training, jboolean random_start,
                                                            jboolean symmetric_alpha, jint device, jint rank,
                                                            jint world_size) {
  (void)cls;
  mrlda_config cfg;
  mrlda_default_config(&cfg);
  cfg.n_topics = topics;
  cfg.n_terms = terms;
  cfg.learning = training ? 1 : 0;
  cfg.random_start = random_start ? 1 : 0;
  cfg.symmetric_alpha = symmetric_alpha ? 1 : 0;
  cfg.device = device;
  cfg.rank = rank;
  cfg.world_size = world_size;
  mrlda_ctx *ctx = NULL;
  if (mrlda_create(&cfg, &ctx) != MRLDA_OK) {
    throw_io(env, mrlda_last_error(NULL));
    return 0;
  }
  return (jlong)(intptr_t)ctx;
}

JNIEXPORT void JNICALL Java_cc_mrlda_b200_MrLdaB200_close(JNIEnv *env, jobject self) {
  jclass cls = (*env)->GetObjectClass(env, self);
  jfieldID fid = (*env)->GetFieldID(env, cls, "handle", "J");
  if (!fid) return;
  mrlda_ctx *ctx = (mrlda_ctx *)(intptr_t)(*env)->GetLongField(env, self, fid);
  if (ctx) {
    mrlda_destroy(ctx);
    (*env)->SetLongField(env, self, fid, (jlong)0);
  }
}

JNIEXPORT void JNICALL Java_cc_mrlda_b200_MrLdaB200_setCorpus(JNIEnv *env, jobject self, jlongArray row_ptr,
                                                              jintArray term_id, jintArray count,
                                                              jintArray doc_id, jdoubleArray gamma0) {
  mrlda_ctx *ctx = ctx_of(env, self);
  if (!ctx) return;
  if (!row_ptr || (*env)->GetArrayLength(env, row_ptr) < 1) {
    throw_io(env, "mrlda_b200: rowPtr must hold D + 1 entries");
    return;
  }
  const int64_t D = (int64_t)(*env)->GetArrayLength(env, row_ptr) - 1;
  const int64_t K = topics_of(env, self);
  const int64_t n_term = term_id ? (*env)->GetArrayLength(env, term_id) : 0;
  const int64_t n_count = count ? (*env)->GetArrayLength(env, count) : 0;
  if (bad_length(env, doc_id, D, 1, "mrlda_b200: docId must hold D entries") ||
      bad_length(env, gamma0, D * K, 1, "mrlda_b200: gamma0 must hold D * K entries"))
    return;
  void *rp = NULL, *ti = NULL, *ct = NULL, *di = NULL, *g0 = NULL;
  int rc = MRLDA_EINVAL, called = 0;
  const char *fail = NULL;
  if (pin(env, row_ptr, &rp) || pin(env, term_id, &ti) || pin(env, count, &ct) || pin(env, doc_id, &di) ||
      pin(env, gamma0, &g0)) {
    fail = kPinFailed;
  } else if (((const int64_t *)rp)[D] < 0 || ((const int64_t *)rp)[D] > n_term ||
             ((const int64_t *)rp)[D] > n_count) {
    fail = "mrlda_b200: termId / count are shorter than rowPtr[D]";
  } else {
    called = 1;
    /* jlong / jint / jdouble are int64_t / int32_t / double on every JVM this library targets */
    rc = mrlda_set_corpus_csr(ctx, D, (const int64_t *)rp, (const int32_t *)ti, (const int32_t *)ct,
                              (const int32_t *)di, (const double *)g0);
  }
  unpin(env, gamma0, g0, JNI_ABORT);
  unpin(env, doc_id, di, JNI_ABORT);
  unpin(env, count, ct, JNI_ABORT);
  unpin(env, term_id, ti, JNI_ABORT);
  unpin(env, row_ptr, rp, JNI_ABORT);
  if (fail) {
    throw_io(env, fail);
    return;
  }
  if (called && !check(env, ctx, rc)) set_docs(env, self, D);
}

JNIEXPORT void JNICALL Java_cc_mrlda_b200_MrLdaB200_setAlpha(JNIEnv *env, jobject self, jdoubleArray alpha) {
  mrlda_ctx *ctx = ctx_of(env, self);
  if (!ctx) return;
  if (bad_length(env, alpha, topics_of(env, self), 0, "mrlda_b200: alpha must hold K entries")) return;
  void *a = NULL;
  if (pin(env, alpha, &a)) {
    throw_io(env, kPinFailed);
    return;
  }
  const int rc = mrlda_set_alpha(ctx, (const double *)a);
  unpin(env, alpha, a, JNI_ABORT);
  check(env, ctx, rc);
}

JNIEXPORT void JNICALL Java_cc_mrlda_b200_MrLdaB200_setBeta(JNIEnv *env, jobject self, jdoubleArray digamma_lambda,
                                                            jfloatArray normaliser, jbyteArray present) {
  mrlda_ctx *ctx = ctx_of(env, self);
  if (!ctx) return;
  const int64_t K = topics_of(env, self), V = terms_of(env, self);
  if (bad_length(env, digamma_lambda, V * K, 0, "mrlda_b200: digammaLambda must hold V * K entries") ||
      bad_length(env, normaliser, K, 0, "mrlda_b200: normaliser must hold K entries") ||
      bad_length(env, present, V, 0, "mrlda_b200: present must hold V entries"))
    return;
  void *dl = NULL, *nm = NULL, *pr = NULL;
  int rc = MRLDA_EINVAL, called = 0;
  if (!pin(env, digamma_lambda, &dl) && !pin(env, normaliser, &nm) && !pin(env, present, &pr)) {
    called = 1;
    rc = mrlda_set_beta(ctx, (const double *)dl, (const float *)nm, (const uint8_t *)pr);
  }
  unpin(env, present, pr, JNI_ABORT);
  unpin(env, normaliser, nm, JNI_ABORT);
  unpin(env, digamma_lambda, dl, JNI_ABORT);
  if (called) check(env, ctx, rc);
  else throw_io(env, kPinFailed);
}

JNIEXPORT void JNICALL Java_cc_mrlda_b200_MrLdaB200_initBetaRandom(JNIEnv *env, jobject self) {
  mrlda_ctx *ctx = ctx_of(env, self);
  if (!ctx) return;
  check(env, ctx, mrlda_init_beta_random(ctx));
}

JNIEXPORT void JNICALL Java_cc_mrlda_b200_MrLdaB200_setInformedPrior(JNIEnv *env, jobject self, jintArray topic,
                                                                     jintArray term) {
  mrlda_ctx *ctx = ctx_of(env, self);
  if (!ctx) return;
  const jsize n = topic ? (*env)->GetArrayLength(env, topic) : 0;
  if (n != (term ? (*env)->GetArrayLength(env, term) : 0)) {
    throw_io(env, "mrlda_b200: topic and term arrays differ in length");
    return;
  }
  void *tp = NULL, *tm = NULL;
  int rc = MRLDA_EINVAL, called = 0;
  if (!pin(env, topic, &tp) && !pin(env, term, &tm)) {
    called = 1;
    rc = mrlda_set_informed_prior(ctx, (int64_t)n, (const int32_t *)tp, (const int32_t *)tm);
  }
  unpin(env, term, tm, JNI_ABORT);
  unpin(env, topic, tp, JNI_ABORT);
  if (called) check(env, ctx, rc);
  else throw_io(env, kPinFailed);
}

JNIEXPORT jlongArray JNICALL Java_cc_mrlda_b200_MrLdaB200_iterate(JNIEnv *env, jobject self) {
  mrlda_ctx *ctx = ctx_of(env, self);
  if (!ctx) return NULL;
  mrlda_iter_result r;
  memset(&r, 0, sizeof r);
  if (check(env, ctx, mrlda_iterate(ctx, &r))) return NULL;
  const jlong out[4] = {(jlong)r.ll_counter, (jlong)r.n_docs, (jlong)r.n_terms_seen, (jlong)r.n_tokens};
  jlongArray a = (*env)->NewLongArray(env, 4);
  if (a) (*env)->SetLongArrayRegion(env, a, 0, 4, out);
  return a;
}

JNIEXPORT jdoubleArray JNICALL Java_cc_mrlda_b200_MrLdaB200_getAlpha(JNIEnv *env, jobject self) {
  mrlda_ctx *ctx = ctx_of(env, self);
  if (!ctx) return NULL;
  jdoubleArray arr = (*env)->NewDoubleArray(env, topics_of(env, self));
  void *a = NULL;
  if (!arr) return NULL; /* OutOfMemoryError is pending */
  if (pin(env, arr, &a)) {
    throw_io(env, kPinFailed);
    return NULL;
  }
  const int rc = mrlda_get_alpha(ctx, (double *)a);
  unpin(env, arr, a, 0);
  return check(env, ctx, rc) ? NULL : arr;
}

JNIEXPORT jdoubleArray JNICALL Java_cc_mrlda_b200_MrLdaB200_getAlphaSufficientStatistics(JNIEnv *env, jobject self) {
  mrlda_ctx *ctx = ctx_of(env, self);
  if (!ctx) return NULL;
  jdoubleArray arr = (*env)->NewDoubleArray(env, topics_of(env, self));
  void *a = NULL;
  if (!arr) return NULL; /* OutOfMemoryError is pending */
  if (pin(env, arr, &a)) {
    throw_io(env, kPinFailed);
    return NULL;
  }
  const int rc = mrlda_get_alpha_ss(ctx, (double *)a);
  unpin(env, arr, a, 0);
  return check(env, ctx, rc) ? NULL : arr;
}

JNIEXPORT void JNICALL Java_cc_mrlda_b200_MrLdaB200_getBeta(JNIEnv *env, jobject self, jdoubleArray digamma_lambda,
                                                            jfloatArray normaliser, jbyteArray present) {
  mrlda_ctx *ctx = ctx_of(env, self);
  if (!ctx) return;
  const int64_t K = topics_of(env, self), V = terms_of(env, self);
  if (bad_length(env, digamma_lambda, V * K, 0, "mrlda_b200: digammaLambda must hold V * K entries") ||
      bad_length(env, normaliser, K, 0, "mrlda_b200: normaliser must hold K entries") ||
      bad_length(env, present, V, 0, "mrlda_b200: present must hold V entries"))
    return;
  void *dl = NULL, *nm = NULL, *pr = NULL;
  int rc = MRLDA_EINVAL, called = 0;
  if (!pin(env, digamma_lambda, &dl) && !pin(env, normaliser, &nm) && !pin(env, present, &pr)) {
    called = 1;
    rc = mrlda_get_beta(ctx, (double *)dl, (float *)nm, (uint8_t *)pr);
  }
  unpin(env, present, pr, 0);
  unpin(env, normaliser, nm, 0);
  unpin(env, digamma_lambda, dl, 0);
  if (called) check(env, ctx, rc);
  else throw_io(env, kPinFailed);
}

JNIEXPORT void JNICALL Java_cc_mrlda_b200_MrLdaB200_getGamma(JNIEnv *env, jobject self, jdoubleArray gamma) {
  mrlda_ctx *ctx = ctx_of(env, self);
  if (!ctx) return;
  if (bad_length(env, gamma, docs_of(env, self) * (int64_t)topics_of(env, self), 0,
                 "mrlda_b200: gamma must hold D * K entries (D of the last setCorpus)"))
    return;
  void *g = NULL;
  if (pin(env, gamma, &g)) {
    throw_io(env, kPinFailed);
    return;
  }
  const int rc = mrlda_get_gamma(ctx, (double *)g);
  unpin(env, gamma, g, 0);
  check(env, ctx, rc);
}

JNIEXPORT jbyteArray JNICALL Java_cc_mrlda_b200_MrLdaB200_ncclUniqueId(JNIEnv *env, jclass cls) {
  (void)cls;
  uint8_t id[MRLDA_NCCL_ID_BYTES];
  if (mrlda_nccl_unique_id(id) != MRLDA_OK) {
    throw_io(env, mrlda_last_error(NULL));
    return NULL;
  }
  jbyteArray a = (*env)->NewByteArray(env, MRLDA_NCCL_ID_BYTES);
  if (a) (*env)->SetByteArrayRegion(env, a, 0, MRLDA_NCCL_ID_BYTES, (const jbyte *)id);
  return a;
}

JNIEXPORT void JNICALL Java_cc_mrlda_b200_MrLdaB200_commInit(JNIEnv *env, jobject self, jbyteArray id) {
  mrlda_ctx *ctx = ctx_of(env, self);
  if (!ctx) return;
  if (!id || (*env)->GetArrayLength(env, id) != MRLDA_NCCL_ID_BYTES) {
    throw_io(env, "mrlda_b200: the NCCL id is 128 bytes");
    return;
  }
  uint8_t buf[MRLDA_NCCL_ID_BYTES];
  (*env)->GetByteArrayRegion(env, id, 0, MRLDA_NCCL_ID_BYTES, (jbyte *)buf);
  check(env, ctx, mrlda_comm_init(ctx, buf));
}
