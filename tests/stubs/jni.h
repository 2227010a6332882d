/* Stand-in for the JDK's <jni.h> in an image without a JDK.  It declares the JNI types and the JNIEnv functions the shim
 * (integration/jni/flashfry_b200_jni.c) uses, with the signatures of the JNI specification.  tests/stubs/jni_mock.c
 * implements them over heap objects, so the shim can be EXECUTED by a C driver (tests/stubs/jni_exec.c); only the
 * member ORDER of the function table differs from a real JNIEnv (nothing here is binary-compatible with a JVM). */
#ifndef FF_TEST_STUB_JNI_H
#define FF_TEST_STUB_JNI_H
#include <stdint.h>

typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef uint8_t jboolean;
typedef double jdouble;
typedef jint jsize;
typedef void *jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef jobject jarray;
typedef jarray jlongArray;
typedef jarray jintArray;
typedef jarray jbyteArray;
typedef jarray jbooleanArray;
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
  jlongArray (*NewLongArray)(JNIEnv *, jsize);
  void (*SetLongArrayRegion)(JNIEnv *, jlongArray, jsize, jsize, const jlong *);
  void (*GetLongArrayRegion)(JNIEnv *, jlongArray, jsize, jsize, jlong *);
  jintArray (*NewIntArray)(JNIEnv *, jsize);
  void (*SetIntArrayRegion)(JNIEnv *, jintArray, jsize, jsize, const jint *);
  void (*GetIntArrayRegion)(JNIEnv *, jintArray, jsize, jsize, jint *);
  jbyteArray (*NewByteArray)(JNIEnv *, jsize);
  void (*SetByteArrayRegion)(JNIEnv *, jbyteArray, jsize, jsize, const jbyte *);
  void (*GetBooleanArrayRegion)(JNIEnv *, jbooleanArray, jsize, jsize, jboolean *);
  jdoubleArray (*NewDoubleArray)(JNIEnv *, jsize);
  void (*SetDoubleArrayRegion)(JNIEnv *, jdoubleArray, jsize, jsize, const jdouble *);
  jobjectArray (*NewObjectArray)(JNIEnv *, jsize, jclass, jobject);
  void (*SetObjectArrayElement)(JNIEnv *, jobjectArray, jsize, jobject);
  jobject (*GetObjectArrayElement)(JNIEnv *, jobjectArray, jsize);
  jobject (*NewDirectByteBuffer)(JNIEnv *, void *, jlong);
};
#endif
