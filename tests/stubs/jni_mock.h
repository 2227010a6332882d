/* A functional mock of the JNIEnv functions in tests/stubs/jni.h: Java arrays, strings and direct buffers as heap objects. */
#ifndef FF_JNI_MOCK_H
#define FF_JNI_MOCK_H
#include <jni.h>
#include <stddef.h>

typedef struct {
  int kind;        /* 'J' long[], 'I' int[], 'B' byte[], 'Z' boolean[], 'D' double[], 'L' Object[], 'S' String, 'P' direct buffer */
  jsize n;
  size_t elem;
  void *data;
} mock_obj;

JNIEnv *mock_env(void);                         /* the function table (one static instance) */
const char *mock_pending_exception(void);       /* message of the last ThrowNew, or NULL; mock_clear_exception resets */
void mock_clear_exception(void);
jstring mock_string(const char *s);
jarray mock_array(int kind, jsize n, const void *init); /* init may be NULL (zeroed) */
void mock_free(jobject o);                      /* (Object[] elements are not freed recursively) */
#define MOCK(o) ((mock_obj *)(o))
#endif
