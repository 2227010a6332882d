/* JNI shim over include/flashfry_b200.h.
 * Build where a JDK exists:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude integration/jni/flashfry_b200_jni.c \
 *       -Lflashfry_b200 -lflashfry_b200 -o libflashfry_b200_jni.so
 * This repository's image has no JDK: tests/test_host_cpu.py only syntax-checks this file against a minimal stand-in
 * for <jni.h> (tests/stubs/jni.h), which proves that every call into flashfry_b200.h is type-correct. */
#include <jni.h>
#include <stddef.h>
#include <stdint.h>
#include "flashfry_b200.h"

static void throw_ise(JNIEnv *env) {   /* the reference fails with IllegalStateException / assertion errors */
  (*env)->ThrowNew(env, (*env)->FindClass(env, "java/lang/IllegalStateException"), ff_last_error());
}

JNIEXPORT jlong JNICALL Java_flashfry_NativeBridge_create(JNIEnv *env, jclass c, jint dev) {
  ff_ctx *ctx = NULL;
  if (ff_create(&ctx, dev) != FF_OK) { throw_ise(env); return 0; }
  return (jlong)(intptr_t)ctx;
}
JNIEXPORT void JNICALL Java_flashfry_NativeBridge_destroy(JNIEnv *env, jclass c, jlong ctx) { ff_destroy((ff_ctx *)(intptr_t)ctx); }

JNIEXPORT void JNICALL Java_flashfry_NativeBridge_loadDatabase(JNIEnv *env, jclass c, jlong ctx, jstring path) {
  const char *p = (*env)->GetStringUTFChars(env, path, NULL);
  int rc = ff_load_database((ff_ctx *)(intptr_t)ctx, p, NULL);      /* NULL header path => p + ".header" */
  (*env)->ReleaseStringUTFChars(env, path, p);
  if (rc != FF_OK) throw_ise(env);
}

JNIEXPORT jlong JNICALL Java_flashfry_NativeBridge_discover(JNIEnv *env, jclass c, jlong ctx, jlongArray guides,
                                                            jint k, jint maxOT, jboolean pos) {
  jsize n = (*env)->GetArrayLength(env, guides);
  jlong *g = (*env)->GetPrimitiveArrayCritical(env, guides, NULL);   /* Java long == the reference's target long */
  ff_hits *h = NULL;
  int rc = ff_discover((ff_ctx *)(intptr_t)ctx, (const uint64_t *)g, n, k, maxOT, pos ? 1 : 0, &h);
  (*env)->ReleasePrimitiveArrayCritical(env, guides, g, JNI_ABORT);
  if (rc != FF_OK) { throw_ise(env); return 0; }
  return (jlong)(intptr_t)h;
}

static jlongArray to_jlongs(JNIEnv *env, const void *src, jsize n) {
  jlongArray a = (*env)->NewLongArray(env, n);
  if (a && n) (*env)->SetLongArrayRegion(env, a, 0, n, (const jlong *)src);
  return a;
}
JNIEXPORT jlongArray JNICALL Java_flashfry_NativeBridge_hitsRowPtr(JNIEnv *env, jclass c, jlong h) {
  const ff_hits *x = (const ff_hits *)(intptr_t)h; return to_jlongs(env, x->row_ptr, (jsize)x->n_guides + 1);
}
JNIEXPORT jlongArray JNICALL Java_flashfry_NativeBridge_hitsTargets(JNIEnv *env, jclass c, jlong h) {
  const ff_hits *x = (const ff_hits *)(intptr_t)h; return to_jlongs(env, x->targets, (jsize)x->n_hits);
}
JNIEXPORT jlongArray JNICALL Java_flashfry_NativeBridge_hitsPosPtr(JNIEnv *env, jclass c, jlong h) {
  const ff_hits *x = (const ff_hits *)(intptr_t)h; return x->pos_ptr ? to_jlongs(env, x->pos_ptr, (jsize)x->n_hits + 1) : NULL;
}
JNIEXPORT jlongArray JNICALL Java_flashfry_NativeBridge_hitsPositions(JNIEnv *env, jclass c, jlong h) {
  const ff_hits *x = (const ff_hits *)(intptr_t)h;
  return x->pos_ptr ? to_jlongs(env, x->positions, (jsize)x->pos_ptr[x->n_hits]) : NULL;
}
JNIEXPORT jlong JNICALL Java_flashfry_NativeBridge_hitsCompares(JNIEnv *env, jclass c, jlong h) { return (jlong)((const ff_hits *)(intptr_t)h)->n_compares; }
JNIEXPORT void JNICALL Java_flashfry_NativeBridge_hitsFree(JNIEnv *env, jclass c, jlong h) { ff_hits_free((ff_hits *)(intptr_t)h); }

JNIEXPORT jobjectArray JNICALL Java_flashfry_NativeBridge_score(JNIEnv *env, jclass c, jlong ctx, jlongArray guides,
                                                                jlongArray rowPtr, jlongArray targets, jint metrics) {
  jsize n = (*env)->GetArrayLength(env, guides), nh = (*env)->GetArrayLength(env, targets);
  jlong *g = (*env)->GetLongArrayElements(env, guides, NULL), *rp = (*env)->GetLongArrayElements(env, rowPtr, NULL),
        *t = (*env)->GetLongArrayElements(env, targets, NULL);
  ff_hits h = {0};
  h.n_guides = n; h.n_hits = nh; h.row_ptr = (const int64_t *)rp; h.targets = (const uint64_t *)t;
  jdoubleArray out[4] = {(*env)->NewDoubleArray(env, n), (*env)->NewDoubleArray(env, n), (*env)->NewDoubleArray(env, n),
                         (*env)->NewDoubleArray(env, nh)};
  jdouble *o[4];
  for (int i = 0; i < 4; ++i) o[i] = (*env)->GetDoubleArrayElements(env, out[i], NULL);
  int rc = ff_score((ff_ctx *)(intptr_t)ctx, (const uint64_t *)g, &h, (uint32_t)metrics, o[0], o[1], o[2], o[3]);
  for (int i = 0; i < 4; ++i) (*env)->ReleaseDoubleArrayElements(env, out[i], o[i], 0);
  (*env)->ReleaseLongArrayElements(env, guides, g, JNI_ABORT);
  (*env)->ReleaseLongArrayElements(env, rowPtr, rp, JNI_ABORT);
  (*env)->ReleaseLongArrayElements(env, targets, t, JNI_ABORT);
  if (rc != FF_OK) { throw_ise(env); return NULL; }
  jobjectArray res = (*env)->NewObjectArray(env, 4, (*env)->FindClass(env, "[D"), NULL);
  for (int i = 0; i < 4; ++i) (*env)->SetObjectArrayElement(env, res, i, out[i]);
  return res;
}
