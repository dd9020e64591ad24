/* Minimal stand-in for the JDK's <jni.h>: just the types and JNIEnv members that
 * java/jni/mrlda_b200_jni.c uses, so that the CPU test suite can run `gcc -fsyntax-only` on the glue
 * in an image without a JDK (tests/test_java_shim.py).  Signatures follow the JNI specification. */
#ifndef MRLDA_TEST_JNI_STUB_H
#define MRLDA_TEST_JNI_STUB_H
#include <stdint.h>

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2

typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef uint8_t jboolean;
typedef float jfloat;
typedef double jdouble;
typedef jint jsize;

struct _jobject;
typedef struct _jobject *jobject;
typedef jobject jclass;
typedef jobject jarray;
typedef jarray jlongArray;
typedef jarray jintArray;
typedef jarray jdoubleArray;
typedef jarray jfloatArray;
typedef jarray jbyteArray;
struct _jfieldID;
typedef struct _jfieldID *jfieldID;

struct JNINativeInterface_;
typedef const struct JNINativeInterface_ *JNIEnv;

struct JNINativeInterface_ {
  jclass (*FindClass)(JNIEnv *, const char *);
  jint (*ThrowNew)(JNIEnv *, jclass, const char *);
  jclass (*GetObjectClass)(JNIEnv *, jobject);
  jfieldID (*GetFieldID)(JNIEnv *, jclass, const char *, const char *);
  jlong (*GetLongField)(JNIEnv *, jobject, jfieldID);
  void (*SetLongField)(JNIEnv *, jobject, jfieldID, jlong);
  jint (*GetIntField)(JNIEnv *, jobject, jfieldID);
  jsize (*GetArrayLength)(JNIEnv *, jarray);
  void *(*GetPrimitiveArrayCritical)(JNIEnv *, jarray, jboolean *);
  void (*ReleasePrimitiveArrayCritical)(JNIEnv *, jarray, void *, jint);
  jlongArray (*NewLongArray)(JNIEnv *, jsize);
  jdoubleArray (*NewDoubleArray)(JNIEnv *, jsize);
  jbyteArray (*NewByteArray)(JNIEnv *, jsize);
  void (*SetLongArrayRegion)(JNIEnv *, jlongArray, jsize, jsize, const jlong *);
  void (*SetByteArrayRegion)(JNIEnv *, jbyteArray, jsize, jsize, const jbyte *);
  void (*GetByteArrayRegion)(JNIEnv *, jbyteArray, jsize, jsize, jbyte *);
};
#endif
