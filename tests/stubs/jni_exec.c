/* Executes every Java_flashfry_NativeBridge_* function of integration/jni/flashfry_b200_jni.c through the mock JNIEnv
 * (tests/stubs/jni_mock.c) against the real libflashfry_b200.so on the GPU, and checks the results against direct
 * calls of the C ABI.  Usage: jni_exec <flashfry database> <scratch dir> <n_gpus>.  Prints JNI_EXEC_OK on success. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "jni_mock.h"
#include "../../integration/jni/flashfry_b200_jni.c"

static int g_checks = 0;
#define CHECK(cond)                                                                                          \
  do {                                                                                                       \
    ++g_checks;                                                                                              \
    if (!(cond)) { fprintf(stderr, "CHECK FAILED %s:%d: %s (pending exception: %s)\n", __FILE__, __LINE__, #cond, \
                           mock_pending_exception() ? mock_pending_exception() : "none"); exit(1); }       \
  } while (0)
#define NOEXC() CHECK(mock_pending_exception() == NULL)
#define LONGS(a) ((jlong *)MOCK(a)->data)
#define INTS(a) ((jint *)MOCK(a)->data)
#define BYTES(a) ((jbyte *)MOCK(a)->data)
#define DOUBLES(a) ((jdouble *)MOCK(a)->data)

int main(int argc, char **argv) {
  if (argc < 4) { fprintf(stderr, "usage: jni_exec DB SCRATCH_DIR N_GPUS\n"); return 2; }
  const char *db = argv[1];
  const int n_gpus = atoi(argv[3]);
  char image[1024], tsv_a[1024], tsv_b[1024];
  snprintf(image, sizeof image, "%s/jni.ffimage", argv[2]);
  snprintf(tsv_a, sizeof tsv_a, "%s/jni_a.tsv", argv[2]);
  snprintf(tsv_b, sizeof tsv_b, "%s/jni_b.tsv", argv[2]);
  JNIEnv *env = mock_env();

  jlong ctx = Java_flashfry_NativeBridge_create(env, NULL, 0); NOEXC(); CHECK(ctx != 0);
  Java_flashfry_NativeBridge_loadDatabase(env, NULL, ctx, mock_string(db)); NOEXC();
  Java_flashfry_NativeBridge_loadDatabase(env, NULL, ctx, mock_string("/nonexistent/db"));
  CHECK(mock_pending_exception() != NULL); mock_clear_exception();      /* a failure surfaces as IllegalStateException */
  Java_flashfry_NativeBridge_loadDatabase(env, NULL, ctx, mock_string(db)); NOEXC();

  /* guides: database targets with one substitution each (count reset to 1) */
  ff_db_info_t info;
  CHECK(ff_db_info(CTX(ctx), &info) == FF_OK && info.n_targets > 200);
  enum { G = 96 };
  uint64_t t[G];
  jlong g[G];
  for (int i = 0; i < G; ++i) {
    CHECK(ff_db_copy_targets(CTX(ctx), (info.n_targets / G) * (uint64_t)i, 1, &t[i]) == FF_OK);
    g[i] = (jlong)(((t[i] & 0xFFFFFFFFFFFFull) ^ ((uint64_t)(1 + i % 3) << (2 * (4 + i % 18)))) | (1ull << 48));
  }
  jlongArray jg = mock_array('J', G, g);

  /* discover through the shim == ff_discover directly */
  ff_hits *ref = NULL;
  CHECK(ff_discover(CTX(ctx), (const uint64_t *)g, G, 4, 2000, 1, &ref) == FF_OK && ref->n_hits > G / 2);
  jlong h = Java_flashfry_NativeBridge_discover(env, NULL, ctx, jg, 4, 2000, 1); NOEXC(); CHECK(h != 0);
  jlongArray rp = Java_flashfry_NativeBridge_hitsRowPtr(env, NULL, h), tg = Java_flashfry_NativeBridge_hitsTargets(env, NULL, h);
  jbyteArray mm = Java_flashfry_NativeBridge_hitsMismatches(env, NULL, h), ov = Java_flashfry_NativeBridge_hitsOverflowed(env, NULL, h);
  jintArray tc = Java_flashfry_NativeBridge_hitsTotalCount(env, NULL, h);
  jlongArray pp = Java_flashfry_NativeBridge_hitsPosPtr(env, NULL, h), ps = Java_flashfry_NativeBridge_hitsPositions(env, NULL, h);
  NOEXC();
  CHECK(MOCK(rp)->n == G + 1 && MOCK(tg)->n == ref->n_hits && MOCK(mm)->n == ref->n_hits && MOCK(tc)->n == G && MOCK(ov)->n == G);
  CHECK(memcmp(LONGS(rp), ref->row_ptr, (G + 1) * 8) == 0 && memcmp(LONGS(tg), ref->targets, ref->n_hits * 8) == 0);
  CHECK(memcmp(BYTES(mm), ref->mismatches, ref->n_hits) == 0 && memcmp(INTS(tc), ref->total_count, G * 4) == 0);
  CHECK(memcmp(BYTES(ov), ref->overflowed, G) == 0);
  CHECK(pp && ps && MOCK(pp)->n == ref->n_hits + 1 && memcmp(LONGS(pp), ref->pos_ptr, (ref->n_hits + 1) * 8) == 0);
  CHECK(MOCK(ps)->n == ref->pos_ptr[ref->n_hits] && memcmp(LONGS(ps), ref->positions, MOCK(ps)->n * 8) == 0);
  CHECK(Java_flashfry_NativeBridge_hitsCompares(env, NULL, h) == (jlong)ref->n_compares);

  /* TSV fast path through the shim == ff_hits_write_tsv directly */
  {
    jobjectArray contigs = mock_array('L', G, NULL), bases = mock_array('L', G, NULL), ctxs = mock_array('L', G, NULL);
    jint starts[G];
    jboolean fwd[G];
    ff_tsv_guide rows[G];
    static char seqs[G][24];
    for (int i = 0; i < G; ++i) {
      for (int b = 0; b < 23; ++b) seqs[i][b] = "ACGT"[((uint64_t)g[i] >> (2 * (22 - b))) & 3];
      seqs[i][23] = 0;
      starts[i] = 10 * i; fwd[i] = (jboolean)(i & 1);
      ((jobject *)MOCK(contigs)->data)[i] = mock_string("guides");
      ((jobject *)MOCK(bases)->data)[i] = mock_string(seqs[i]);
      ((jobject *)MOCK(ctxs)->data)[i] = (i % 4) ? mock_string("NNNNNN") : NULL;
      rows[i].contig = "guides"; rows[i].start = starts[i]; rows[i].bases = seqs[i]; rows[i].context = (i % 4) ? "NNNNNN" : NULL; rows[i].forward = i & 1;
    }
    Java_flashfry_NativeBridge_hitsWriteTsv(env, NULL, ctx, h, mock_string(tsv_a), contigs, mock_array('I', G, starts), bases, ctxs,
                                            mock_array('Z', G, fwd), 1);
    NOEXC();
    CHECK(ff_hits_write_tsv(CTX(ctx), tsv_b, rows, ref, 1) == FF_OK);
    FILE *fa = fopen(tsv_a, "rb"), *fb = fopen(tsv_b, "rb");
    CHECK(fa && fb);
    int ca, cb, nbytes = 0;
    do { ca = fgetc(fa); cb = fgetc(fb); ++nbytes; CHECK(ca == cb); } while (ca != EOF);
    CHECK(nbytes > 1000);
    fclose(fa); fclose(fb);
    Java_flashfry_NativeBridge_hitsWriteTsv(env, NULL, ctx, h, mock_string(tsv_a), mock_array('L', 3, NULL), mock_array('I', G, starts), bases, ctxs,
                                            mock_array('Z', G, fwd), 1);
    CHECK(mock_pending_exception() != NULL); mock_clear_exception();    /* array lengths are validated */
  }

  /* scores: discoverScore == score over the CSR == ff_score directly */
  jdoubleArray c0 = mock_array('D', G, NULL), c1 = mock_array('D', G, NULL), c2 = mock_array('D', G, NULL);
  jlong hs = Java_flashfry_NativeBridge_discoverScore(env, NULL, ctx, jg, 4, 2000, 0, 3, c0, c1, c2); NOEXC(); CHECK(hs != 0);
  jobjectArray sc = Java_flashfry_NativeBridge_score(env, NULL, ctx, info.enzyme_index, jg, rp, tg, 3); NOEXC(); CHECK(sc != NULL);
  double d0[G], d1[G], d2[G];
  CHECK(ff_score(CTX(ctx), (const uint64_t *)g, ref, 3, d0, d1, d2, NULL) == FF_OK);
  for (int i = 0; i < G; ++i) {
    jdoubleArray *o = (jdoubleArray *)MOCK(sc)->data;
    CHECK(DOUBLES(c0)[i] == d0[i] && DOUBLES(c1)[i] == d1[i] && DOUBLES(c2)[i] == d2[i]);
    CHECK(DOUBLES(o[0])[i] == d0[i] && DOUBLES(o[1])[i] == d1[i] && DOUBLES(o[2])[i] == d2[i]);
  }
  CHECK(MOCK(((jdoubleArray *)MOCK(sc)->data)[3])->n == ref->n_hits);
  Java_flashfry_NativeBridge_hitsFree(env, NULL, hs);
  CHECK(Java_flashfry_NativeBridge_score(env, NULL, ctx, info.enzyme_index, jg, mock_array('J', G, LONGS(rp)), tg, 3) == NULL);
  CHECK(mock_pending_exception() != NULL); mock_clear_exception();      /* rowPtr one entry short: refused before the native read */
  CHECK(Java_flashfry_NativeBridge_score(env, NULL, ctx, 1 /* Cpf1 */, jg, rp, tg, 3) == NULL);
  CHECK(mock_pending_exception() != NULL); mock_clear_exception();      /* validOverEnzyme */

  /* minot / dangerous_in_genome */
  jintArray ag = Java_flashfry_NativeBridge_hitAggregates(env, NULL, ctx, info.enzyme_index, jg, rp, tg); NOEXC(); CHECK(ag && MOCK(ag)->n == 8 * G);
  int32_t cl[G], cn[G], hist[5 * G], ing[G];
  CHECK(ff_hit_aggregates(CTX(ctx), info.enzyme_index, (const uint64_t *)g, ref, cl, cn, hist, ing) == FF_OK);
  for (int i = 0; i < G; ++i) {
    const jint *r = INTS(ag) + 8 * i;
    CHECK(r[0] == cl[i] && r[1] == cn[i] && r[7] == ing[i]);
    for (int m = 0; m < 5; ++m) CHECK(r[2 + m] == hist[5 * i + m]);
  }

  /* compact hit lists: indices + host mirror, resolve */
  Java_flashfry_NativeBridge_setOption(env, NULL, ctx, mock_string("compact_hits"), 1); NOEXC();
  Java_flashfry_NativeBridge_setOption(env, NULL, ctx, mock_string("no_such_option"), 1);
  CHECK(mock_pending_exception() != NULL); mock_clear_exception();
  jlong hc = Java_flashfry_NativeBridge_discover(env, NULL, ctx, jg, 4, 2000, 0); NOEXC();
  jintArray ti = Java_flashfry_NativeBridge_hitsTargetIndex(env, NULL, hc); NOEXC(); CHECK(ti && MOCK(ti)->n == ref->n_hits);
  CHECK(Java_flashfry_NativeBridge_hitsTargets(env, NULL, hc) == NULL && mock_pending_exception() != NULL); mock_clear_exception();
  jobject mirror = Java_flashfry_NativeBridge_dbHostTargets(env, NULL, ctx); NOEXC(); CHECK(mirror != NULL);
  for (int64_t i = 0; i < ref->n_hits; ++i) CHECK(((const uint64_t *)MOCK(mirror)->data)[(uint32_t)INTS(ti)[i]] == ref->targets[i]);
  Java_flashfry_NativeBridge_hitsResolve(env, NULL, ctx, hc); NOEXC();
  jlongArray tr = Java_flashfry_NativeBridge_hitsTargets(env, NULL, hc); NOEXC();
  CHECK(memcmp(LONGS(tr), ref->targets, ref->n_hits * 8) == 0);
  Java_flashfry_NativeBridge_hitsFree(env, NULL, hc);
  Java_flashfry_NativeBridge_setOption(env, NULL, ctx, mock_string("compact_hits"), 0); NOEXC();

  /* image side-car: save, load into a second context, same answers */
  Java_flashfry_NativeBridge_saveImage(env, NULL, ctx, mock_string(image)); NOEXC();
  jlong ctx2 = Java_flashfry_NativeBridge_create(env, NULL, 0); NOEXC();
  Java_flashfry_NativeBridge_loadImage(env, NULL, ctx2, mock_string(image)); NOEXC();
  jlong h2 = Java_flashfry_NativeBridge_discover(env, NULL, ctx2, jg, 4, 2000, 1); NOEXC();
  jlongArray tg2 = Java_flashfry_NativeBridge_hitsTargets(env, NULL, h2);
  CHECK(MOCK(tg2)->n == ref->n_hits && memcmp(LONGS(tg2), ref->targets, ref->n_hits * 8) == 0);
  Java_flashfry_NativeBridge_hitsFree(env, NULL, h2);
  Java_flashfry_NativeBridge_destroy(env, NULL, ctx2);

  /* several GPUs behind the one process */
  {
    jint devs[2] = {0, 1};
    const int nd = n_gpus >= 2 ? 2 : 1;
    jlong m = Java_flashfry_NativeBridge_multiCreate(env, NULL, mock_array('I', nd, devs)); NOEXC(); CHECK(m != 0);
    Java_flashfry_NativeBridge_multiLoadDatabase(env, NULL, m, mock_string(db)); NOEXC();
    jintArray totals = mock_array('I', G, NULL);
    jlongArray hh = Java_flashfry_NativeBridge_multiDiscover(env, NULL, m, jg, 4, 2000, 0, totals); NOEXC(); CHECK(hh && MOCK(hh)->n == nd);
    CHECK(memcmp(INTS(totals), ref->total_count, G * 4) == 0);
    int64_t hits = 0;
    for (int r = 0; r < nd; ++r) { hits += HITS(LONGS(hh)[r])->n_hits; Java_flashfry_NativeBridge_hitsFree(env, NULL, LONGS(hh)[r]); }
    CHECK(hits == ref->n_hits);
    Java_flashfry_NativeBridge_multiDestroy(env, NULL, m);
  }
  /* the same with the index work sharded (shard_mode = 1: candidates over peer memory); two ranks, on one device if need be */
  {
    jint devs[2] = {0, n_gpus >= 2 ? 1 : 0};
    jlong m = Java_flashfry_NativeBridge_multiCreate(env, NULL, mock_array('I', 2, devs)); NOEXC(); CHECK(m != 0);
    Java_flashfry_NativeBridge_multiSetOption(env, NULL, m, mock_string("shard_mode"), 1); NOEXC();
    Java_flashfry_NativeBridge_multiLoadDatabase(env, NULL, m, mock_string(db)); NOEXC();
    jintArray totals = mock_array('I', G, NULL);
    jlongArray hh = Java_flashfry_NativeBridge_multiDiscover(env, NULL, m, jg, 4, 2000, 0, totals); NOEXC(); CHECK(hh && MOCK(hh)->n == 2);
    CHECK(memcmp(INTS(totals), ref->total_count, G * 4) == 0);
    int64_t hits = 0;
    for (int r = 0; r < 2; ++r) { hits += HITS(LONGS(hh)[r])->n_hits; Java_flashfry_NativeBridge_hitsFree(env, NULL, LONGS(hh)[r]); }
    CHECK(hits == ref->n_hits);
    Java_flashfry_NativeBridge_multiDestroy(env, NULL, m);
  }

  Java_flashfry_NativeBridge_hitsFree(env, NULL, h);
  ff_hits_free(ref);
  Java_flashfry_NativeBridge_destroy(env, NULL, ctx);
  printf("JNI_EXEC_OK %d checks, %d guides\n", g_checks, G);
  return 0;
}
