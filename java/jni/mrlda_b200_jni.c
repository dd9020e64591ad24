/*
 * mrlda_b200_jni.c — JNI glue between cc.mrlda.b200.MrLdaB200 and the C ABI of libmrlda_b200.so
 * (include/mrlda_b200.h).  One native method = one ABI call; Java arrays are pinned with
 * GetPrimitiveArrayCritical for the duration of the call (the library copies to / from the device,
 * nothing outlives the call) and every failure becomes a java.io.IOException carrying
 * mrlda_last_error().
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

/* pins a (possibly null) primitive array; returns 0 on success */
static int pin(JNIEnv *env, jarray arr, void **out) {
  *out = NULL;
  if (!arr) return 0;
  *out = (*env)->GetPrimitiveArrayCritical(env, arr, NULL);
  if (!*out) {
    throw_io(env, "mrlda_b200: could not pin a Java array");
    return 1;
  }
  return 0;
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
  if (!row_ptr) {
    throw_io(env, "mrlda_b200: rowPtr is null");
    return;
  }
  const jsize n = (*env)->GetArrayLength(env, row_ptr);
  void *rp = NULL, *ti = NULL, *ct = NULL, *di = NULL, *g0 = NULL;
  int rc = MRLDA_EINVAL, pinned = 0;
  if (!pin(env, row_ptr, &rp) && !pin(env, term_id, &ti) && !pin(env, count, &ct) && !pin(env, doc_id, &di) &&
      !pin(env, gamma0, &g0)) {
    pinned = 1;
    /* jlong / jint / jdouble are int64_t / int32_t / double on every JVM this library targets */
    rc = mrlda_set_corpus_csr(ctx, (int64_t)n - 1, (const int64_t *)rp, (const int32_t *)ti,
                              (const int32_t *)ct, (const int32_t *)di, (const double *)g0);
  }
  unpin(env, gamma0, g0, JNI_ABORT);
  unpin(env, doc_id, di, JNI_ABORT);
  unpin(env, count, ct, JNI_ABORT);
  unpin(env, term_id, ti, JNI_ABORT);
  unpin(env, row_ptr, rp, JNI_ABORT);
  if (pinned) check(env, ctx, rc);
}

JNIEXPORT void JNICALL Java_cc_mrlda_b200_MrLdaB200_setAlpha(JNIEnv *env, jobject self, jdoubleArray alpha) {
  mrlda_ctx *ctx = ctx_of(env, self);
  if (!ctx) return;
  void *a = NULL;
  if (pin(env, alpha, &a)) return;
  const int rc = mrlda_set_alpha(ctx, (const double *)a);
  unpin(env, alpha, a, JNI_ABORT);
  check(env, ctx, rc);
}

JNIEXPORT void JNICALL Java_cc_mrlda_b200_MrLdaB200_setBeta(JNIEnv *env, jobject self, jdoubleArray digamma_lambda,
                                                            jfloatArray normaliser, jbyteArray present) {
  mrlda_ctx *ctx = ctx_of(env, self);
  if (!ctx) return;
  void *dl = NULL, *nm = NULL, *pr = NULL;
  int rc = MRLDA_EINVAL, pinned = 0;
  if (!pin(env, digamma_lambda, &dl) && !pin(env, normaliser, &nm) && !pin(env, present, &pr)) {
    pinned = 1;
    rc = mrlda_set_beta(ctx, (const double *)dl, (const float *)nm, (const uint8_t *)pr);
  }
  unpin(env, present, pr, JNI_ABORT);
  unpin(env, normaliser, nm, JNI_ABORT);
  unpin(env, digamma_lambda, dl, JNI_ABORT);
  if (pinned) check(env, ctx, rc);
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
  int rc = MRLDA_EINVAL, pinned = 0;
  if (!pin(env, topic, &tp) && !pin(env, term, &tm)) {
    pinned = 1;
    rc = mrlda_set_informed_prior(ctx, (int64_t)n, (const int32_t *)tp, (const int32_t *)tm);
  }
  unpin(env, term, tm, JNI_ABORT);
  unpin(env, topic, tp, JNI_ABORT);
  if (pinned) check(env, ctx, rc);
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

/* K of the Java object (final int field "topics") */
static jint topics_of(JNIEnv *env, jobject self) {
  jclass cls = (*env)->GetObjectClass(env, self);
  jfieldID fid = (*env)->GetFieldID(env, cls, "topics", "I");
  return fid ? (*env)->GetIntField(env, self, fid) : 0;
}

JNIEXPORT jdoubleArray JNICALL Java_cc_mrlda_b200_MrLdaB200_getAlpha(JNIEnv *env, jobject self) {
  mrlda_ctx *ctx = ctx_of(env, self);
  if (!ctx) return NULL;
  jdoubleArray arr = (*env)->NewDoubleArray(env, topics_of(env, self));
  void *a = NULL;
  if (!arr || pin(env, arr, &a)) return NULL;
  const int rc = mrlda_get_alpha(ctx, (double *)a);
  unpin(env, arr, a, 0);
  return check(env, ctx, rc) ? NULL : arr;
}

JNIEXPORT jdoubleArray JNICALL Java_cc_mrlda_b200_MrLdaB200_getAlphaSufficientStatistics(JNIEnv *env, jobject self) {
  mrlda_ctx *ctx = ctx_of(env, self);
  if (!ctx) return NULL;
  jdoubleArray arr = (*env)->NewDoubleArray(env, topics_of(env, self));
  void *a = NULL;
  if (!arr || pin(env, arr, &a)) return NULL;
  const int rc = mrlda_get_alpha_ss(ctx, (double *)a);
  unpin(env, arr, a, 0);
  return check(env, ctx, rc) ? NULL : arr;
}

JNIEXPORT void JNICALL Java_cc_mrlda_b200_MrLdaB200_getBeta(JNIEnv *env, jobject self, jdoubleArray digamma_lambda,
                                                            jfloatArray normaliser, jbyteArray present) {
  mrlda_ctx *ctx = ctx_of(env, self);
  if (!ctx) return;
  void *dl = NULL, *nm = NULL, *pr = NULL;
  int rc = MRLDA_EINVAL, pinned = 0;
  if (!pin(env, digamma_lambda, &dl) && !pin(env, normaliser, &nm) && !pin(env, present, &pr)) {
    pinned = 1;
    rc = mrlda_get_beta(ctx, (double *)dl, (float *)nm, (uint8_t *)pr);
  }
  unpin(env, present, pr, 0);
  unpin(env, normaliser, nm, 0);
  unpin(env, digamma_lambda, dl, 0);
  if (pinned) check(env, ctx, rc);
}

JNIEXPORT void JNICALL Java_cc_mrlda_b200_MrLdaB200_getGamma(JNIEnv *env, jobject self, jdoubleArray gamma) {
  mrlda_ctx *ctx = ctx_of(env, self);
  if (!ctx) return;
  void *g = NULL;
  if (pin(env, gamma, &g)) return;
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
