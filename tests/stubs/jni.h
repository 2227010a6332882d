/* Minimal stand-in for the JDK's <jni.h>: ONLY for the syntax check of integration/jni/flashfry_b200_jni.c in an image
 * without a JDK.  It declares the JNI types and the handful of JNIEnv functions the shim uses, with the signatures of
 * the JNI specification; it is not a usable JNI header and nothing links against it. */
#ifndef FF_TEST_STUB_JNI_H
#define FF_TEST_STUB_JNI_H
#include <stdint.h>

typedef int32_t jint;
typedef int64_t jlong;
typedef uint8_t jboolean;
typedef double jdouble;
typedef jint jsize;
typedef void *jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef jobject jarray;
typedef jarray jlongArray;
typedef jarray jdoubleArray;
typedef jarray jobjectArray;
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2

struct JNINativeInterface_;
typedef const struct JNINativeInterface_ *JNIEnv;
struct JNINativeInterface_ {
  jclass (*FindClass)(JNIEnv *, const char *);
  jint (*ThrowNew)(JNIEnv *, jclass, const char *);
  const char *(*GetStringUTFChars)(JNIEnv *, jstring, jboolean *);
  void (*ReleaseStringUTFChars)(JNIEnv *, jstring, const char *);
  jsize (*GetArrayLength)(JNIEnv *, jarray);
  void *(*GetPrimitiveArrayCritical)(JNIEnv *, jarray, jboolean *);
  void (*ReleasePrimitiveArrayCritical)(JNIEnv *, jarray, void *, jint);
  jlongArray (*NewLongArray)(JNIEnv *, jsize);
  void (*SetLongArrayRegion)(JNIEnv *, jlongArray, jsize, jsize, const jlong *);
  jlong *(*GetLongArrayElements)(JNIEnv *, jlongArray, jboolean *);
  void (*ReleaseLongArrayElements)(JNIEnv *, jlongArray, jlong *, jint);
  jdoubleArray (*NewDoubleArray)(JNIEnv *, jsize);
  jdouble *(*GetDoubleArrayElements)(JNIEnv *, jdoubleArray, jboolean *);
  void (*ReleaseDoubleArrayElements)(JNIEnv *, jdoubleArray, jdouble *, jint);
  jobjectArray (*NewObjectArray)(JNIEnv *, jsize, jclass, jobject);
  void (*SetObjectArrayElement)(JNIEnv *, jobjectArray, jsize, jobject);
};
#endif
