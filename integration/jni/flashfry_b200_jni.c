/* JNI shim over include/flashfry_b200.h.
 * Build where a JDK exists:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude integration/jni/flashfry_b200_jni.c \
 *       -Lflashfry_b200 -lflashfry_b200 -o libflashfry_b200_jni.so
 * This repository's image has no JDK.  tests/stubs/jni.h + tests/stubs/jni_mock.c are a FUNCTIONAL stand-in for the few
 * JNIEnv functions used here (heap-backed arrays and strings), and tests/stubs/jni_exec.c calls every
 * Java_flashfry_NativeBridge_* function below through it against the real library on the GPU
 * (tests/test_gpu_jni.py) -- so the shim's logic is executed, not only type-checked.
 *
 * JNI rules kept here: no GetPrimitiveArrayCritical around a GPU call (the critical region would stall every JVM thread
 * that needs a GC for the whole discover); arrays are copied in with Get<T>ArrayRegion into malloc'd buffers that are
 * freed before returning; array lengths are validated before the native side reads them. */
#include <jni.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "flashfry_b200.h"

static void throw_ise(JNIEnv *env, const char *msg) {   /* the reference fails with IllegalStateException / assertion errors */
  (*env)->ThrowNew(env, (*env)->FindClass(env, "java/lang/IllegalStateException"), msg ? msg : ff_last_error());
}
#define CTX(x) ((ff_ctx *)(intptr_t)(x))
#define HITS(x) ((ff_hits *)(intptr_t)(x))

/* copy a long[] into a malloc'd buffer (caller frees); NULL + exception on failure */
static int64_t *copy_longs(JNIEnv *env, jlongArray a, jsize *n_out) {
  jsize n = a ? (*env)->GetArrayLength(env, a) : 0;
  int64_t *buf = (int64_t *)malloc(((size_t)n + 1) * sizeof(int64_t));
  if (!buf) { throw_ise(env, "out of host memory"); return NULL; }
  if (n) (*env)->GetLongArrayRegion(env, a, 0, n, (jlong *)buf);
  *n_out = n;
  return buf;
}

JNIEXPORT jlong JNICALL Java_flashfry_NativeBridge_create(JNIEnv *env, jclass c, jint dev) {
  ff_ctx *ctx = NULL;
  if (ff_create(&ctx, dev) != FF_OK) { throw_ise(env, NULL); return 0; }
  return (jlong)(intptr_t)ctx;
}
JNIEXPORT void JNICALL Java_flashfry_NativeBridge_destroy(JNIEnv *env, jclass c, jlong ctx) { ff_destroy(CTX(ctx)); }

JNIEXPORT void JNICALL Java_flashfry_NativeBridge_setOption(JNIEnv *env, jclass c, jlong ctx, jstring key, jlong value) {
  const char *k = (*env)->GetStringUTFChars(env, key, NULL);
  int rc = ff_set_option(CTX(ctx), k, (long long)value);
  (*env)->ReleaseStringUTFChars(env, key, k);
  if (rc != FF_OK) throw_ise(env, NULL);
}

JNIEXPORT void JNICALL Java_flashfry_NativeBridge_loadDatabase(JNIEnv *env, jclass c, jlong ctx, jstring path) {
  const char *p = (*env)->GetStringUTFChars(env, path, NULL);
  int rc = ff_load_database(CTX(ctx), p, NULL);      /* NULL header path => p + ".header" */
  (*env)->ReleaseStringUTFChars(env, path, p);
  if (rc != FF_OK) throw_ise(env, NULL);
}
JNIEXPORT void JNICALL Java_flashfry_NativeBridge_loadImage(JNIEnv *env, jclass c, jlong ctx, jstring path) {
  const char *p = (*env)->GetStringUTFChars(env, path, NULL);
  int rc = ff_load_image(CTX(ctx), p);
  (*env)->ReleaseStringUTFChars(env, path, p);
  if (rc != FF_OK) throw_ise(env, NULL);
}
JNIEXPORT void JNICALL Java_flashfry_NativeBridge_saveImage(JNIEnv *env, jclass c, jlong ctx, jstring path) {
  const char *p = (*env)->GetStringUTFChars(env, path, NULL);
  int rc = ff_save_image(CTX(ctx), p);
  (*env)->ReleaseStringUTFChars(env, path, p);
  if (rc != FF_OK) throw_ise(env, NULL);
}

JNIEXPORT jlong JNICALL Java_flashfry_NativeBridge_discover(JNIEnv *env, jclass c, jlong ctx, jlongArray guides,
                                                            jint k, jint maxOT, jboolean pos) {
  jsize n = 0;
  int64_t *g = copy_longs(env, guides, &n);           /* Java long == the reference's target long */
  if (!g) return 0;
  ff_hits *h = NULL;
  int rc = ff_discover(CTX(ctx), (const uint64_t *)g, n, k, maxOT, pos ? 1 : 0, &h);
  free(g);
  if (rc != FF_OK) { throw_ise(env, NULL); return 0; }
  return (jlong)(intptr_t)h;
}

/* ff_discover_score: scores = {cfdMax[n], cfdSpecificity[n], hsu2013[n]} filled in place; returns the ff_hits* */
JNIEXPORT jlong JNICALL Java_flashfry_NativeBridge_discoverScore(JNIEnv *env, jclass c, jlong ctx, jlongArray guides, jint k,
                                                                 jint maxOT, jboolean pos, jint metrics, jdoubleArray cfdMax,
                                                                 jdoubleArray cfdSpec, jdoubleArray hsu) {
  jsize n = 0;
  int64_t *g = copy_longs(env, guides, &n);
  if (!g) return 0;
  if ((*env)->GetArrayLength(env, cfdMax) < n || (*env)->GetArrayLength(env, cfdSpec) < n || (*env)->GetArrayLength(env, hsu) < n) {
    free(g);
    throw_ise(env, "score arrays are shorter than the guide array");
    return 0;
  }
  double *s = (double *)malloc(((size_t)n + 1) * 3 * sizeof(double));
  if (!s) { free(g); throw_ise(env, "out of host memory"); return 0; }
  ff_hits *h = NULL;
  int rc = ff_discover_score(CTX(ctx), (const uint64_t *)g, n, k, maxOT, pos ? 1 : 0, (uint32_t)metrics, &h, s, s + n, s + 2 * (size_t)n);
  free(g);
  if (rc == FF_OK && n) {
    (*env)->SetDoubleArrayRegion(env, cfdMax, 0, n, s);
    (*env)->SetDoubleArrayRegion(env, cfdSpec, 0, n, s + n);
    (*env)->SetDoubleArrayRegion(env, hsu, 0, n, s + 2 * (size_t)n);
  }
  free(s);
  if (rc != FF_OK) { throw_ise(env, NULL); return 0; }
  return (jlong)(intptr_t)h;
}

static jlongArray to_jlongs(JNIEnv *env, const void *src, jsize n) {
  jlongArray a = (*env)->NewLongArray(env, n);
  if (a && n) (*env)->SetLongArrayRegion(env, a, 0, n, (const jlong *)src);
  return a;
}
JNIEXPORT jlongArray JNICALL Java_flashfry_NativeBridge_hitsRowPtr(JNIEnv *env, jclass c, jlong h) {
  return to_jlongs(env, HITS(h)->row_ptr, (jsize)HITS(h)->n_guides + 1);
}
JNIEXPORT jlongArray JNICALL Java_flashfry_NativeBridge_hitsTargets(JNIEnv *env, jclass c, jlong h) {
  if (!HITS(h)->targets && HITS(h)->n_hits > 0) { throw_ise(env, "compact hit list: call hitsResolve first or use hitsTargetIndex"); return NULL; }
  return to_jlongs(env, HITS(h)->targets, (jsize)HITS(h)->n_hits);
}
JNIEXPORT jbyteArray JNICALL Java_flashfry_NativeBridge_hitsMismatches(JNIEnv *env, jclass c, jlong h) {
  jbyteArray a = (*env)->NewByteArray(env, (jsize)HITS(h)->n_hits);
  if (a && HITS(h)->n_hits) (*env)->SetByteArrayRegion(env, a, 0, (jsize)HITS(h)->n_hits, (const jbyte *)HITS(h)->mismatches);
  return a;
}
JNIEXPORT jintArray JNICALL Java_flashfry_NativeBridge_hitsTargetIndex(JNIEnv *env, jclass c, jlong h) {  /* option compact_hits */
  if (!HITS(h)->target_index) return NULL;
  jintArray a = (*env)->NewIntArray(env, (jsize)HITS(h)->n_hits);
  if (a && HITS(h)->n_hits) (*env)->SetIntArrayRegion(env, a, 0, (jsize)HITS(h)->n_hits, (const jint *)HITS(h)->target_index);
  return a;
}
JNIEXPORT void JNICALL Java_flashfry_NativeBridge_hitsResolve(JNIEnv *env, jclass c, jlong ctx, jlong h) {
  if (ff_hits_resolve(CTX(ctx), HITS(h)) != FF_OK) throw_ise(env, NULL);
}
/* the host mirror of the target array as a direct LongBuffer-able region: a CRISPRHit's long = mirror[targetIndex] */
JNIEXPORT jobject JNICALL Java_flashfry_NativeBridge_dbHostTargets(JNIEnv *env, jclass c, jlong ctx) {
  ff_db_info_t info;
  const uint64_t *m = ff_db_host_targets(CTX(ctx));
  if (!m || ff_db_info(CTX(ctx), &info) != FF_OK) { throw_ise(env, NULL); return NULL; }
  return (*env)->NewDirectByteBuffer(env, (void *)m, (jlong)(info.n_targets * 8));
}
JNIEXPORT jintArray JNICALL Java_flashfry_NativeBridge_hitsTotalCount(JNIEnv *env, jclass c, jlong h) {
  jintArray a = (*env)->NewIntArray(env, (jsize)HITS(h)->n_guides);
  if (a && HITS(h)->n_guides) (*env)->SetIntArrayRegion(env, a, 0, (jsize)HITS(h)->n_guides, (const jint *)HITS(h)->total_count);
  return a;
}
JNIEXPORT jbyteArray JNICALL Java_flashfry_NativeBridge_hitsOverflowed(JNIEnv *env, jclass c, jlong h) {
  jbyteArray a = (*env)->NewByteArray(env, (jsize)HITS(h)->n_guides);
  if (a && HITS(h)->n_guides) (*env)->SetByteArrayRegion(env, a, 0, (jsize)HITS(h)->n_guides, (const jbyte *)HITS(h)->overflowed);
  return a;
}
JNIEXPORT jlongArray JNICALL Java_flashfry_NativeBridge_hitsPosPtr(JNIEnv *env, jclass c, jlong h) {
  return HITS(h)->pos_ptr ? to_jlongs(env, HITS(h)->pos_ptr, (jsize)HITS(h)->n_hits + 1) : NULL;
}
JNIEXPORT jlongArray JNICALL Java_flashfry_NativeBridge_hitsPositions(JNIEnv *env, jclass c, jlong h) {
  return HITS(h)->pos_ptr ? to_jlongs(env, HITS(h)->positions, (jsize)HITS(h)->pos_ptr[HITS(h)->n_hits]) : NULL;
}
JNIEXPORT jlong JNICALL Java_flashfry_NativeBridge_hitsCompares(JNIEnv *env, jclass c, jlong h) { return (jlong)HITS(h)->n_compares; }
JNIEXPORT void JNICALL Java_flashfry_NativeBridge_hitsFree(JNIEnv *env, jclass c, jlong h) { ff_hits_free(HITS(h)); }

/* ff_hits_write_tsv: contigs / bases / contexts are String[] (context entries may be null = "NONE") */
JNIEXPORT void JNICALL Java_flashfry_NativeBridge_hitsWriteTsv(JNIEnv *env, jclass c, jlong ctx, jlong h, jstring path, jobjectArray contigs,
                                                               jintArray starts, jobjectArray bases, jobjectArray contexts,
                                                               jbooleanArray forward, jboolean positions) {
  const jsize n = (jsize)HITS(h)->n_guides;
  if ((*env)->GetArrayLength(env, contigs) != n || (*env)->GetArrayLength(env, starts) != n || (*env)->GetArrayLength(env, bases) != n ||
      (*env)->GetArrayLength(env, contexts) != n || (*env)->GetArrayLength(env, forward) != n) {
    throw_ise(env, "guide description arrays must have one entry per row of the hit list");
    return;
  }
  ff_tsv_guide *rows = (ff_tsv_guide *)calloc((size_t)n + 1, sizeof(ff_tsv_guide));
  jint *st = (jint *)malloc(((size_t)n + 1) * sizeof(jint));
  jboolean *fw = (jboolean *)malloc((size_t)n + 1);
  if (!rows || !st || !fw) { free(rows); free(st); free(fw); throw_ise(env, "out of host memory"); return; }
  if (n) { (*env)->GetIntArrayRegion(env, starts, 0, n, st); (*env)->GetBooleanArrayRegion(env, forward, 0, n, fw); }
  for (jsize i = 0; i < n; ++i) {
    jstring sc = (jstring)(*env)->GetObjectArrayElement(env, contigs, i), sb = (jstring)(*env)->GetObjectArrayElement(env, bases, i),
            sx = (jstring)(*env)->GetObjectArrayElement(env, contexts, i);
    rows[i].contig = (*env)->GetStringUTFChars(env, sc, NULL);
    rows[i].bases = (*env)->GetStringUTFChars(env, sb, NULL);
    rows[i].context = sx ? (*env)->GetStringUTFChars(env, sx, NULL) : NULL;
    rows[i].start = st[i]; rows[i].forward = fw[i] ? 1 : 0;
  }
  const char *p = (*env)->GetStringUTFChars(env, path, NULL);
  int rc = ff_hits_write_tsv(CTX(ctx), p, rows, HITS(h), positions ? 1 : 0);
  (*env)->ReleaseStringUTFChars(env, path, p);
  for (jsize i = 0; i < n; ++i) {
    (*env)->ReleaseStringUTFChars(env, (jstring)(*env)->GetObjectArrayElement(env, contigs, i), rows[i].contig);
    (*env)->ReleaseStringUTFChars(env, (jstring)(*env)->GetObjectArrayElement(env, bases, i), rows[i].bases);
    if (rows[i].context) (*env)->ReleaseStringUTFChars(env, (jstring)(*env)->GetObjectArrayElement(env, contexts, i), rows[i].context);
  }
  free(rows); free(st); free(fw);
  if (rc != FF_OK) throw_ise(env, NULL);
}

/* ff_score_enzyme over guides + CSR hit list; out = {cfdMax[n], cfdSpecificity[n], hsu2013[n], perOtCfd[nHits]} */
JNIEXPORT jobjectArray JNICALL Java_flashfry_NativeBridge_score(JNIEnv *env, jclass c, jlong ctx, jint enzymeIndex, jlongArray guides,
                                                                jlongArray rowPtr, jlongArray targets, jint metrics) {
  jsize n = 0, nrp = 0, nh = 0;
  int64_t *g = copy_longs(env, guides, &n), *rp = copy_longs(env, rowPtr, &nrp), *t = copy_longs(env, targets, &nh);
  double *o = NULL;
  jobjectArray res = NULL;
  if (!g || !rp || !t) goto done;
  if (nrp != n + 1) { throw_ise(env, "rowPtr must hold nGuides + 1 entries"); goto done; }  /* ff_score reads row_ptr[nGuides] */
  o = (double *)malloc((3 * ((size_t)n + 1) + (size_t)nh + 1) * sizeof(double));
  if (!o) { throw_ise(env, "out of host memory"); goto done; }
  {
    ff_hits h;
    memset(&h, 0, sizeof h);
    h.n_guides = n; h.n_hits = nh; h.row_ptr = rp; h.targets = (const uint64_t *)t;
    double *o0 = o, *o1 = o + n + 1, *o2 = o + 2 * ((size_t)n + 1), *o3 = o + 3 * ((size_t)n + 1);
    if (ff_score_enzyme(CTX(ctx), enzymeIndex, (const uint64_t *)g, &h, (uint32_t)metrics, o0, o1, o2, o3) != FF_OK) { throw_ise(env, NULL); goto done; }
    res = (*env)->NewObjectArray(env, 4, (*env)->FindClass(env, "[D"), NULL);
    const double *src[4] = {o0, o1, o2, o3};
    const jsize len[4] = {n, n, n, nh};
    for (int i = 0; i < 4; ++i) {
      jdoubleArray a = (*env)->NewDoubleArray(env, len[i]);
      if (len[i]) (*env)->SetDoubleArrayRegion(env, a, 0, len[i], src[i]);
      (*env)->SetObjectArrayElement(env, res, i, a);
    }
  }
done:
  free(g); free(rp); free(t); free(o);
  return res;
}

/* ff_hit_aggregates ("minot" + dangerous_in_genome): out = int[n * 8] = {closest, closestCount, hist[5], inGenome} per guide */
JNIEXPORT jintArray JNICALL Java_flashfry_NativeBridge_hitAggregates(JNIEnv *env, jclass c, jlong ctx, jint enzymeIndex, jlongArray guides,
                                                                    jlongArray rowPtr, jlongArray targets) {
  jsize n = 0, nrp = 0, nh = 0;
  int64_t *g = copy_longs(env, guides, &n), *rp = copy_longs(env, rowPtr, &nrp), *t = copy_longs(env, targets, &nh);
  int32_t *o = NULL;
  jintArray res = NULL;
  if (!g || !rp || !t) goto done;
  if (nrp != n + 1) { throw_ise(env, "rowPtr must hold nGuides + 1 entries"); goto done; }
  o = (int32_t *)calloc(8 * ((size_t)n + 1), sizeof(int32_t));
  if (!o) { throw_ise(env, "out of host memory"); goto done; }
  {
    ff_hits h;
    memset(&h, 0, sizeof h);
    h.n_guides = n; h.n_hits = nh; h.row_ptr = rp; h.targets = (const uint64_t *)t;
    int32_t *closest = o, *cnt = o + (n + 1), *hist = o + 2 * ((size_t)n + 1), *ing = o + 7 * ((size_t)n + 1);
    if (ff_hit_aggregates(CTX(ctx), enzymeIndex, (const uint64_t *)g, &h, closest, cnt, hist, ing) != FF_OK) { throw_ise(env, NULL); goto done; }
    res = (*env)->NewIntArray(env, n * 8);
    for (jsize i = 0; i < n; ++i) {
      jint row[8] = {closest[i], cnt[i], hist[5 * i], hist[5 * i + 1], hist[5 * i + 2], hist[5 * i + 3], hist[5 * i + 4], ing[i]};
      (*env)->SetIntArrayRegion(env, res, i * 8, 8, row);
    }
  }
done:
  free(g); free(rp); free(t); free(o);
  return res;
}

/* ---- several GPUs behind the one JVM process (ff_multi) ---- */
JNIEXPORT jlong JNICALL Java_flashfry_NativeBridge_multiCreate(JNIEnv *env, jclass c, jintArray devices) {
  jsize n = (*env)->GetArrayLength(env, devices);
  jint *d = (jint *)malloc(((size_t)n + 1) * sizeof(jint));
  if (!d) { throw_ise(env, "out of host memory"); return 0; }
  if (n) (*env)->GetIntArrayRegion(env, devices, 0, n, d);
  ff_multi *m = NULL;
  int rc = ff_multi_create(&m, (const int *)d, n);
  free(d);
  if (rc != FF_OK) { throw_ise(env, NULL); return 0; }
  return (jlong)(intptr_t)m;
}
JNIEXPORT void JNICALL Java_flashfry_NativeBridge_multiDestroy(JNIEnv *env, jclass c, jlong m) { ff_multi_destroy((ff_multi *)(intptr_t)m); }
JNIEXPORT void JNICALL Java_flashfry_NativeBridge_multiSetOption(JNIEnv *env, jclass c, jlong m, jstring key, jlong value) {
  const char *k = (*env)->GetStringUTFChars(env, key, NULL);
  int rc = ff_multi_set_option((ff_multi *)(intptr_t)m, k, (long long)value);
  (*env)->ReleaseStringUTFChars(env, key, k);
  if (rc != FF_OK) throw_ise(env, NULL);
}
JNIEXPORT void JNICALL Java_flashfry_NativeBridge_multiLoadDatabase(JNIEnv *env, jclass c, jlong m, jstring path) {
  const char *p = (*env)->GetStringUTFChars(env, path, NULL);
  int rc = ff_multi_load_database((ff_multi *)(intptr_t)m, p, NULL);
  (*env)->ReleaseStringUTFChars(env, path, p);
  if (rc != FF_OK) throw_ise(env, NULL);
}
/* returns long[nDevices] of ff_hits* handles (shard r = guides ff_shard_range(r)); totals[nGuides] filled with the
 * all-gathered CRISPRSiteOT.currentTotal of every guide */
JNIEXPORT jlongArray JNICALL Java_flashfry_NativeBridge_multiDiscover(JNIEnv *env, jclass c, jlong m, jlongArray guides, jint k, jint maxOT,
                                                                     jboolean pos, jintArray totals) {
  jsize n = 0;
  int64_t *g = copy_longs(env, guides, &n);
  if (!g) return NULL;
  const int nd = ff_multi_size((ff_multi *)(intptr_t)m);
  ff_hits **out = (ff_hits **)calloc((size_t)nd + 1, sizeof(ff_hits *));
  int32_t *tot = (int32_t *)calloc((size_t)n + 1, sizeof(int32_t));
  jlongArray res = NULL;
  if (!out || !tot) { throw_ise(env, "out of host memory"); goto done; }
  if (totals && (*env)->GetArrayLength(env, totals) < n) { throw_ise(env, "totals is shorter than the guide array"); goto done; }
  if (ff_multi_discover((ff_multi *)(intptr_t)m, (const uint64_t *)g, n, k, maxOT, pos ? 1 : 0, out, tot) != FF_OK) { throw_ise(env, NULL); goto done; }
  if (totals && n) (*env)->SetIntArrayRegion(env, totals, 0, n, (const jint *)tot);
  {
    jlong handles[64];
    for (int r = 0; r < nd && r < 64; ++r) handles[r] = (jlong)(intptr_t)out[r];
    res = (*env)->NewLongArray(env, nd);
    (*env)->SetLongArrayRegion(env, res, 0, nd, handles);
  }
done:
  free(g); free(out); free(tot);
  return res;
}
