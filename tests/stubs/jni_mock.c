#include "jni_mock.h"

#include <stdlib.h>
#include <string.h>

static char g_exc[1024];
static int g_has_exc;

static size_t elem_size(int kind) {
  switch (kind) { case 'J': return 8; case 'I': return 4; case 'B': case 'Z': return 1; case 'D': return 8; case 'L': return sizeof(void *); default: return 1; }
}
jarray mock_array(int kind, jsize n, const void *init) {
  mock_obj *o = (mock_obj *)calloc(1, sizeof *o);
  o->kind = kind; o->n = n; o->elem = elem_size(kind);
  o->data = calloc((size_t)n + 1, o->elem);
  if (init && n) memcpy(o->data, init, (size_t)n * o->elem);
  return o;
}
jstring mock_string(const char *s) {
  mock_obj *o = (mock_obj *)calloc(1, sizeof *o);
  o->kind = 'S'; o->n = (jsize)strlen(s); o->elem = 1;
  o->data = strdup(s);
  return o;
}
void mock_free(jobject x) { if (x) { if (MOCK(x)->kind != 'P') free(MOCK(x)->data); free(x); } }
const char *mock_pending_exception(void) { return g_has_exc ? g_exc : NULL; }
void mock_clear_exception(void) { g_has_exc = 0; }

static jclass FindClass(JNIEnv *e, const char *name) { (void)e; return mock_string(name); }
static jint ThrowNew(JNIEnv *e, jclass c, const char *msg) { (void)e; (void)c; strncpy(g_exc, msg ? msg : "", sizeof g_exc - 1); g_has_exc = 1; return 0; }
static const char *GetStringUTFChars(JNIEnv *e, jstring s, jboolean *copy) { (void)e; if (copy) *copy = 0; return (const char *)MOCK(s)->data; }
static void ReleaseStringUTFChars(JNIEnv *e, jstring s, const char *p) { (void)e; (void)s; (void)p; }
static jsize GetArrayLength(JNIEnv *e, jarray a) { (void)e; return MOCK(a)->n; }
#define REGION(NAME, T, KIND)                                                                                             \
  static jarray New##NAME##Array(JNIEnv *e, jsize n) { (void)e; return mock_array(KIND, n, NULL); }                         \
  static void Set##NAME##ArrayRegion(JNIEnv *e, jarray a, jsize s, jsize n, const T *src) {                                 \
    (void)e; if (s < 0 || n < 0 || s + n > MOCK(a)->n) { ThrowNew(e, NULL, "ArrayIndexOutOfBoundsException"); return; }      \
    memcpy((T *)MOCK(a)->data + s, src, (size_t)n * sizeof(T));                                                             \
  }                                                                                                                         \
  static void Get##NAME##ArrayRegion(JNIEnv *e, jarray a, jsize s, jsize n, T *dst) {                                       \
    (void)e; if (s < 0 || n < 0 || s + n > MOCK(a)->n) { ThrowNew(e, NULL, "ArrayIndexOutOfBoundsException"); return; }      \
    memcpy(dst, (const T *)MOCK(a)->data + s, (size_t)n * sizeof(T));                                                       \
  }
REGION(Long, jlong, 'J')
REGION(Int, jint, 'I')
REGION(Byte, jbyte, 'B')
REGION(Boolean, jboolean, 'Z')
REGION(Double, jdouble, 'D')
static jobjectArray NewObjectArray(JNIEnv *e, jsize n, jclass c, jobject init) { (void)e; (void)c; (void)init; return mock_array('L', n, NULL); }
static void SetObjectArrayElement(JNIEnv *e, jobjectArray a, jsize i, jobject v) { (void)e; ((jobject *)MOCK(a)->data)[i] = v; }
static jobject GetObjectArrayElement(JNIEnv *e, jobjectArray a, jsize i) { (void)e; return ((jobject *)MOCK(a)->data)[i]; }
static jobject NewDirectByteBuffer(JNIEnv *e, void *p, jlong cap) {
  (void)e;
  mock_obj *o = (mock_obj *)calloc(1, sizeof *o);
  o->kind = 'P'; o->n = (jsize)(cap > 0x7fffffff ? 0x7fffffff : cap); o->elem = 1; o->data = p;
  return o;
}

static const struct JNINativeInterface_ g_table = {
    FindClass, ThrowNew, GetStringUTFChars, ReleaseStringUTFChars, GetArrayLength,
    NewLongArray, SetLongArrayRegion, GetLongArrayRegion, NewIntArray, SetIntArrayRegion, GetIntArrayRegion,
    NewByteArray, SetByteArrayRegion, GetBooleanArrayRegion, NewDoubleArray, SetDoubleArrayRegion,
    NewObjectArray, SetObjectArrayElement, GetObjectArrayElement, NewDirectByteBuffer};
static const struct JNINativeInterface_ *g_env = &g_table;
JNIEnv *mock_env(void) { return &g_env; }
