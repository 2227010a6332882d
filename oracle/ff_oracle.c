/*
 * ff_oracle.c -- CPU restatement of FlashFry's discovery + CFD/Hsu2013 hot path (plain C, gcc).
 *
 * TEST INFRASTRUCTURE ONLY: the checker for the CUDA path and the CPU baseline for bench.py.  It is
 * never linked into or called from the product library.  Parity pin status: PINNED against the
 * reference's integration-test md5s and unit-test known answers (tests/test_oracle_pins.py).
 *
 * The loops deliberately keep the reference's order of work (bin -> sub-bin filter -> target x guide),
 * not an optimised one: it doubles as the "C restatement of the reference's loop order" CPU baseline.
 * Citations are file:line in the FlashFry checkout (src/main/scala/...).
 */
#include "ff_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define FF_TABLE_QUAL static const
#include "score_tables.h"

/* bitcoding/BitEncoding.scala:192-210 */
#define STRING_MASK 0xFFFFFFFFFFFFULL
#define UPPER_BITS 0xAAAAAAAAAAAAULL

/* ---------------------------------------------------------------------------------------------- */
/* enzyme parameter packs: standards/StandardScanParameters.scala:61-70,90-215                     */
int ffo_pack_from_index(int idx, ffo_pack *o) {
  o->enzyme_index = idx;
  switch (idx) {
    case 1: o->scan_len = 24; o->pam_len = 4; o->five_prime = 1; o->cmp_mask = 0x00FFFFFFFFFFULL; return 0; /* :199-215 */
    case 2: /* :90-109 */
    case 3: /* :134-153 */
    case 4: /* :178-197 */
      o->scan_len = 23; o->pam_len = 3; o->five_prime = 0; o->cmp_mask = 0x3FFFFFFFFFC0ULL; return 0;
    case 5: /* :112-131 */
    case 6: /* :156-175 */
      o->scan_len = 22; o->pam_len = 3; o->five_prime = 0; o->cmp_mask = 0x0FFFFFFFFFC0ULL; return 0;
    default: return -1;
  }
}

/* ---------------------------------------------------------------------------------------------- */
/* BitEncoding.scala:46-67 */
uint64_t ffo_encode(const char *s, int len, int count, int *err) {
  uint64_t enc = 0;
  if (err) *err = 0;
  if (len > 24 || count < 1 || count > 32767) { if (err) *err = 1; return 0; }
  for (int i = 0; i < len; i++) {
    enc <<= 2;
    switch (s[i]) {
      case 'A': case 'a': break;
      case 'C': case 'c': enc |= 1; break;
      case 'G': case 'g': enc |= 2; break;
      case 'T': case 't': enc |= 3; break;
      default: if (err) *err = 1; return 0;
    }
  }
  return enc | ((uint64_t)count << 48);
}

/* BitEncoding.scala:85-99 */
int ffo_decode(uint64_t enc, int len, char *out) {
  static const char L[4] = {'A', 'C', 'G', 'T'};
  for (int i = 0; i < len; i++) out[len - 1 - i] = L[(enc >> (2 * i)) & 3];
  out[len] = 0;
  return (int)(int16_t)(enc >> 48);
}

/* BitEncoding.scala:127-132 */
static inline int mismatches_raw(uint64_t cmp_mask, uint64_t a, uint64_t b, uint64_t additional) {
  uint64_t first = (a ^ b) & additional & cmp_mask;
  return __builtin_popcountll((first & UPPER_BITS) | ((first << 1) & UPPER_BITS));
}
int ffo_mismatches(const ffo_pack *p, uint64_t a, uint64_t b, uint64_t additional) {
  return mismatches_raw(p->cmp_mask, a, b, additional);
}

/* BitEncoding.scala:178-185 binShift */
static inline uint64_t bin_shift(const ffo_pack *p, int bin_size, uint64_t base, int right_shift) {
  int sh = p->five_prime ? 2 * (p->scan_len - (bin_size + p->pam_len + right_shift))
                         : 2 * (p->scan_len - (bin_size + right_shift));
  return (base << sh) & STRING_MASK;
}
/* BitEncoding.scala:153-170 */
void ffo_bin_comparitor(const ffo_pack *p, uint64_t bin_code, int bin_size, int right_shift,
                        uint64_t *bin_long, uint64_t *guide_mask) {
  /* bitEncodeString(bin) carries count=1 in bit 48; binShift's "& stringMask" drops it (:181,:183) */
  *bin_long = bin_shift(p, bin_size, bin_code | (1ULL << 48), right_shift);
  *guide_mask = bin_shift(p, bin_size, STRING_MASK >> (48 - 2 * bin_size), right_shift);
}
/* BitEncoding.scala:142-144 */
int ffo_mismatch_bin(const ffo_pack *p, uint64_t bin_long, uint64_t guide_mask, uint64_t guide) {
  return mismatches_raw(p->cmp_mask, bin_long, guide & guide_mask, STRING_MASK);
}

/* ---------------------------------------------------------------------------------------------- */
/* growable hit list per guide == CRISPRSiteOT.offTargets (crispr/CRISPRSiteOT.scala:31-46)        */
typedef struct {
  uint64_t target;
  int64_t pos_start; /* index into db of the first position long, -1 if no positions */
} hit_t;

typedef struct {
  hit_t *h;
  int64_t n, cap;
  int32_t total; /* currentTotal */
} gstate_t;

static void gs_push(gstate_t *g, uint64_t target, int64_t pos_start) {
  if (g->n == g->cap) {
    g->cap = g->cap ? g->cap * 2 : 16;
    g->h = (hit_t *)realloc(g->h, (size_t)g->cap * sizeof(hit_t));
  }
  g->h[g->n].target = target;
  g->h[g->n].pos_start = pos_start;
  g->n++;
}

/* ResultsAggregator.scala:61-69 updateOT + CRISPRSiteOT.scala:39-46; returns 1 if the guide just became full */
static inline int update_ot(gstate_t *g, int max_ot, uint64_t target, int64_t pos_start) {
  if (g->total >= max_ot) return 0; /* full */
  gs_push(g, target, pos_start);
  g->total += (int)(int16_t)(target >> 48); /* positions.length == count */
  return g->total >= max_ot;
}

typedef struct { int64_t *v; int64_t n, cap; } ivec_t;
static void iv_push(ivec_t *a, int64_t x) {
  if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 64; a->v = (int64_t *)realloc(a->v, (size_t)a->cap * 8); }
  a->v[a->n++] = x;
}

typedef struct {
  const ffo_pack *p;
  int max_mm, max_ot;
  const uint64_t *guides;
  gstate_t *gs;
  uint64_t n_all, n_tc, n_targets;
  /* traversal overflow callback state (OrderedBinTraversalFactory.scala:90 / LinearTraversal.scala:63-76) */
  uint8_t *excluded;
  uint64_t sub_long[256], sub_mask[256]; /* BlockManager.scala:46-49 blockDescriptorLookup */
} run_t;

/* BlockManager.scala:212-254 compareLinearBlock over block[0..n_longs) ; base = index of block[0] inside db */
static void compare_linear_block(run_t *r, const uint64_t *block, int64_t n_longs, int64_t base, int has_pos,
                                 const int64_t *gl, int64_t ng) {
  int64_t off = 0;
  while (off < n_longs) {
    uint64_t t = block[off];
    int count = has_pos ? (int)(int16_t)(t >> 48) : 0;
    r->n_targets++;
    for (int64_t j = 0; j < ng; j++) {
      int64_t gi = gl[j];
      r->n_tc++;
      r->n_all++;
      if (mismatches_raw(r->p->cmp_mask, r->guides[gi], t, STRING_MASK) <= r->max_mm) {
        if (update_ot(&r->gs[gi], r->max_ot, t, has_pos ? base + off + 1 : -1)) r->excluded[gi] = 1;
      }
    }
    off += count + 1;
  }
}

/* BlockManager.scala:143-201 compareIndexedBlock; block points just after the type long */
static void compare_indexed_block(run_t *r, const uint64_t *block, int64_t base, uint64_t pb_long, uint64_t pb_mask,
                                  const int64_t *gl, int64_t ng, int64_t *scratch) {
  for (int bi = 0; bi < 256; bi++) {
    uint64_t ps = block[bi];
    int32_t pos = (int32_t)(ps >> 32);
    int32_t size = (int32_t)(ps & 0xFFFFFFFFULL);
    if (pos >= 0 && size > 0) {
      uint64_t full_mask = r->sub_mask[bi] | pb_mask;
      uint64_t full_long = pb_long | r->sub_long[bi];
      int64_t nn = 0;
      for (int64_t j = 0; j < ng; j++) { /* :186-191 */
        r->n_all++;
        if (mismatches_raw(r->p->cmp_mask, r->guides[gl[j]], full_long, full_mask) <= r->max_mm) scratch[nn++] = gl[j];
      }
      if (nn > 0) compare_linear_block(r, block + 256 + pos, size, base + 256 + pos, 1, scratch, nn);
    }
  }
}

/* BlockManager.scala:63-90 compareBlock */
static int compare_block(run_t *r, const uint64_t *db, int64_t lo, int64_t hi, uint64_t pb_long, uint64_t pb_mask,
                         const int64_t *gl, int64_t ng, int64_t *scratch) {
  uint64_t first = db[lo];
  if (first == 1) compare_linear_block(r, db + lo + 1, hi - lo - 1, lo + 1, 1, gl, ng);
  else if (first == 2) compare_indexed_block(r, db + lo + 1, lo + 1, pb_long, pb_mask, gl, ng, scratch);
  else return -2; /* IllegalStateException("Invalid bin type") :85-87 */
  return 0;
}

static ffo_hits *collect(run_t *r, int64_t n_guides, const uint64_t *db, int with_pos) {
  ffo_hits *h = (ffo_hits *)calloc(1, sizeof(ffo_hits));
  h->n_guides = n_guides;
  h->row_ptr = (int64_t *)calloc((size_t)n_guides + 1, 8);
  h->total_count = (int32_t *)calloc((size_t)n_guides + 1, 4);
  h->overflowed = (uint8_t *)calloc((size_t)n_guides + 1, 1);
  int64_t nh = 0, np = 0;
  for (int64_t g = 0; g < n_guides; g++) {
    h->row_ptr[g] = nh;
    nh += r->gs[g].n;
    for (int64_t i = 0; i < r->gs[g].n; i++) np += (int)(int16_t)(r->gs[g].h[i].target >> 48);
    h->total_count[g] = r->gs[g].total;
    h->overflowed[g] = r->gs[g].total >= r->max_ot; /* CRISPRSiteOT.full :39 */
  }
  h->row_ptr[n_guides] = nh;
  h->targets = (uint64_t *)malloc((size_t)(nh + 1) * 8);
  h->mismatches = (uint8_t *)malloc((size_t)nh + 1);
  if (with_pos) {
    h->pos_ptr = (int64_t *)malloc((size_t)(nh + 1) * 8);
    h->positions = (uint64_t *)malloc((size_t)(np + 1) * 8);
  }
  int64_t k = 0, pk = 0;
  for (int64_t g = 0; g < n_guides; g++) {
    for (int64_t i = 0; i < r->gs[g].n; i++, k++) {
      uint64_t t = r->gs[g].h[i].target;
      h->targets[k] = t;
      h->mismatches[k] = (uint8_t)mismatches_raw(r->p->cmp_mask, r->guides[g], t, STRING_MASK);
      if (with_pos) {
        int c = (int)(int16_t)(t >> 48);
        h->pos_ptr[k] = pk;
        memcpy(h->positions + pk, db + r->gs[g].h[i].pos_start, (size_t)c * 8);
        pk += c;
      }
    }
    free(r->gs[g].h);
  }
  if (with_pos) h->pos_ptr[nh] = pk;
  h->n_compares = r->n_all;
  h->n_target_compares = r->n_tc;
  h->n_targets_scanned = r->n_targets;
  return h;
}

void ffo_hits_free(ffo_hits *h) {
  if (!h) return;
  free(h->row_ptr); free(h->targets); free(h->mismatches); free(h->pos_ptr); free(h->positions);
  free(h->total_count); free(h->overflowed); free(h->bulge); free(h);
}

static void init_sub_bins(run_t *r, int bin_width) {
  /* BlockManager.scala:46-49: BaseCombinationGenerator(4) x binToLongComparitor(bin, offset=binWidth) */
  for (int s = 0; s < 256; s++) ffo_bin_comparitor(r->p, (uint64_t)s, 4, bin_width, &r->sub_long[s], &r->sub_mask[s]);
}

int ffo_discover_blocks(const ffo_pack *p, int bin_width, const uint64_t *db, const int64_t *bin_off,
                        const uint64_t *guides, int64_t n_guides, int max_mm, int max_ot, int force_linear,
                        ffo_hits **out) {
  const int64_t n_bins = 1LL << (2 * bin_width);
  run_t r;
  memset(&r, 0, sizeof r);
  r.p = p; r.max_mm = max_mm; r.max_ot = max_ot; r.guides = guides;
  r.gs = (gstate_t *)calloc((size_t)n_guides + 1, sizeof(gstate_t));
  r.excluded = (uint8_t *)calloc((size_t)n_guides + 1, 1);
  init_sub_bins(&r, bin_width);
  int64_t *scratch = (int64_t *)malloc((size_t)(n_guides + 1) * 8);
  int64_t *cur = (int64_t *)malloc((size_t)(n_guides + 1) * 8);
  int rc = 0, saturated = 0, bins_visited = 0;

  /* OrderedBinTraversalFactory.scala:146-177 : bin -> guides precompute with early saturation exit */
  ivec_t lists = {0, 0, 0};
  int64_t *list_off = NULL;
  int64_t n_listed_bins = 0;
  if (!force_linear) {
    list_off = (int64_t *)malloc((size_t)(n_bins + 1) * 8);
    int64_t index = 0;
    for (; index < n_bins; index++) {
      uint64_t bl, bm;
      ffo_bin_comparitor(p, (uint64_t)index, bin_width, 0, &bl, &bm);
      list_off[index] = lists.n;
      int64_t before = lists.n;
      for (int64_t g = 0; g < n_guides; g++) {
        r.n_all++;
        if (mismatches_raw(p->cmp_mask, bl, guides[g] & bm, STRING_MASK) <= max_mm) iv_push(&lists, g);
      }
      if (lists.n > before) n_listed_bins++;
      if (index % 500 == 0) {
        double sat = (double)n_listed_bins / (index > 0 ? (double)(index + 1) : 1.0);
        if (sat >= 0.95 && index >= 500) { saturated = 1; index++; break; }
      }
    }
    for (int64_t i = index; i <= n_bins; i++) list_off[i] = lists.n;
    if ((double)n_listed_bins / (double)n_bins >= 0.95) saturated = 1;
  }

  if (force_linear || saturated) {
    /* LinearTraversal.scala:82-97 + LinearTraverser.scala:85-110: every bin, guide subset recomputed on the fly
       from the guides that have not overflowed (:63-76) */
    for (int64_t b = 0; b < n_bins && rc == 0; b++) {
      uint64_t bl, bm;
      ffo_bin_comparitor(p, (uint64_t)b, bin_width, 0, &bl, &bm);
      int64_t ng = 0;
      for (int64_t g = 0; g < n_guides; g++) {
        if (r.excluded[g]) continue;
        r.n_all++;
        if (mismatches_raw(p->cmp_mask, bl, guides[g] & bm, STRING_MASK) <= max_mm) cur[ng++] = g;
      }
      rc = compare_block(&r, db, bin_off[b], bin_off[b + 1], bl, bm, cur, ng, scratch);
      bins_visited++;
    }
  } else {
    /* OrderedBinTraversalFactory.scala:74-134 iterator + SeekTraverser.scala:78-102.
       The iterator prepares bin i+1's (filtered) guide list while handing out bin i (:100-117), so a guide that
       overflows inside bin i is still listed for bin i+1; updateOT ignores it there.  excluded_prev models that. */
    uint8_t *excl_snapshot = (uint8_t *)calloc((size_t)n_guides + 1, 1);
    int first = 1;
    for (int64_t b = 0; b < n_bins && rc == 0; b++) {
      int64_t lo = list_off[b], hi = list_off[b + 1];
      if (hi == lo) continue; /* bin not in binToTargets */
      int64_t ng = 0;
      for (int64_t i = lo; i < hi; i++) {
        int64_t g = lists.v[i];
        if (!first && excl_snapshot[g]) continue; /* the first bin is cached unfiltered (:79-86) */
        cur[ng++] = g;
      }
      /* snapshot of guidesToExclude taken when this bin was handed out == state before processing it */
      memcpy(excl_snapshot, r.excluded, (size_t)n_guides);
      first = 0;
      uint64_t bl, bm;
      ffo_bin_comparitor(p, (uint64_t)b, bin_width, 0, &bl, &bm);
      rc = compare_block(&r, db, bin_off[b], bin_off[b + 1], bl, bm, cur, ng, scratch);
      bins_visited++;
    }
    free(excl_snapshot);
  }

  ffo_hits *h = collect(&r, n_guides, db, 1);
  h->saturated = saturated;
  h->bins_visited = bins_visited;
  free(scratch); free(cur); free(lists.v); free(list_off); free(r.gs); free(r.excluded);
  if (rc != 0) { ffo_hits_free(h); return rc; }
  *out = h;
  return 0;
}

/* ---------------------------------------------------------------------------------------------- */
/* SoA variant: same loop order, positions-free, optional bin-parallel host threads                 */
typedef struct { int64_t g; uint64_t t; } bhit_t;
typedef struct { bhit_t *v; int64_t n, cap; } bvec_t;

int ffo_discover_soa(const ffo_pack *p, int bin_width, const uint64_t *targets, const int64_t *bin_off,
                     const uint64_t *guides, int64_t n_guides, int max_mm, int max_ot, int n_threads,
                     ffo_hits **out) {
  const int64_t n_bins = 1LL << (2 * bin_width);
  run_t r;
  memset(&r, 0, sizeof r);
  r.p = p; r.max_mm = max_mm; r.max_ot = max_ot; r.guides = guides;
  r.gs = (gstate_t *)calloc((size_t)n_guides + 1, sizeof(gstate_t));
  r.excluded = (uint8_t *)calloc((size_t)n_guides + 1, 1);
  init_sub_bins(&r, bin_width);
  if (n_threads < 1) n_threads = 1;
  const int sub_shift = 2 * (p->scan_len - (bin_width + 4)); /* 3' PAM only; 5' PAM bins are never indexed */
  uint64_t n_all = 0, n_tc = 0, n_targets = 0;

  /* per-bin hit buffers so that bins can run on any thread and still be merged in database order */
  bvec_t *per_bin = (bvec_t *)calloc((size_t)n_bins, sizeof(bvec_t));
#ifdef _OPENMP
#pragma omp parallel num_threads(n_threads) reduction(+ : n_all, n_tc, n_targets)
#endif
  {
    int64_t *cur = (int64_t *)malloc((size_t)(n_guides + 1) * 8);
    int64_t *sub = (int64_t *)malloc((size_t)(n_guides + 1) * 8);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 8)
#endif
    for (int64_t b = 0; b < n_bins; b++) {
      uint64_t bl, bm;
      ffo_bin_comparitor(p, (uint64_t)b, bin_width, 0, &bl, &bm);
      int64_t ng = 0;
      for (int64_t g = 0; g < n_guides; g++) { /* LinearTraversal.scala:88-94 */
        n_all++;
        if (mismatches_raw(p->cmp_mask, bl, guides[g] & bm, STRING_MASK) <= max_mm) cur[ng++] = g;
      }
      const int64_t lo = bin_off[b], hi = bin_off[b + 1];
      bvec_t *bv = &per_bin[b];
      if (hi - lo > 500 && !p->five_prime) { /* DatabaseWriter.scala:85 -> indexed block */
        int64_t i = lo;
        while (i < hi) {
          int s = (int)((targets[i] >> sub_shift) & 0xFF);
          int64_t e = i;
          while (e < hi && (int)((targets[e] >> sub_shift) & 0xFF) == s) e++;
          uint64_t full_mask = r.sub_mask[s] | bm, full_long = bl | r.sub_long[s];
          int64_t nn = 0;
          for (int64_t j = 0; j < ng; j++) { /* BlockManager.scala:186-191 */
            n_all++;
            if (mismatches_raw(p->cmp_mask, guides[cur[j]], full_long, full_mask) <= max_mm) sub[nn++] = cur[j];
          }
          if (nn > 0) {
            for (int64_t t = i; t < e; t++) { /* BlockManager.scala:225-253 */
              n_targets++;
              for (int64_t j = 0; j < nn; j++) {
                n_tc++; n_all++;
                if (mismatches_raw(p->cmp_mask, guides[sub[j]], targets[t], STRING_MASK) <= max_mm) {
                  if (bv->n == bv->cap) { bv->cap = bv->cap ? bv->cap * 2 : 16; bv->v = (bhit_t *)realloc(bv->v, (size_t)bv->cap * sizeof(bhit_t)); }
                  bv->v[bv->n].g = sub[j]; bv->v[bv->n].t = targets[t]; bv->n++;
                }
              }
            }
          }
          i = e;
        }
      } else {
        for (int64_t t = lo; t < hi; t++) {
          n_targets++;
          for (int64_t j = 0; j < ng; j++) {
            n_tc++; n_all++;
            if (mismatches_raw(p->cmp_mask, guides[cur[j]], targets[t], STRING_MASK) <= max_mm) {
              if (bv->n == bv->cap) { bv->cap = bv->cap ? bv->cap * 2 : 16; bv->v = (bhit_t *)realloc(bv->v, (size_t)bv->cap * sizeof(bhit_t)); }
              bv->v[bv->n].g = cur[j]; bv->v[bv->n].t = targets[t]; bv->n++;
            }
          }
        }
      }
    }
    free(cur); free(sub);
  }
  /* ordered merge == the sequential updateOT stream (ResultsAggregator.scala:61-69) */
  for (int64_t b = 0; b < n_bins; b++) {
    for (int64_t i = 0; i < per_bin[b].n; i++) update_ot(&r.gs[per_bin[b].v[i].g], max_ot, per_bin[b].v[i].t, -1);
    free(per_bin[b].v);
  }
  free(per_bin);
  r.n_all = n_all; r.n_tc = n_tc; r.n_targets = n_targets;
  ffo_hits *h = collect(&r, n_guides, NULL, 0);
  h->saturated = 1;
  h->bins_visited = (int)n_bins;
  free(r.gs); free(r.excluded);
  *out = h;
  return 0;
}

/* ---------------------------------------------------------------------------------------------- */
/* scorers                                                                                          */
static inline int base_at(uint64_t enc, int scan_len, int i) { return (int)((enc >> (2 * (scan_len - 1 - i))) & 3); }

/* Doench2016CFDScore.scala:132-151: product, in ascending position order, of the mismatch weights */
double ffo_cfd_pair(uint64_t guide, uint64_t ot) {
  double score = 1.0;
  for (int i = 0; i < 20; i++) {
    int g = base_at(guide, 23, i), o = base_at(ot, 23, i);
    if (g != o) score *= FF_CFD_MM[i][g][o];
  }
  return score;
}

/* Doench2016CFDScore.scala:53-88 */
void ffo_cfd_guide(uint64_t guide, const uint64_t *ots, int64_t n, double *max_out, double *spec_out, double *per_ot) {
  const uint64_t proto = 0x3FFFFFFFFFC0ULL;
  double sum = 0.0, mx = 0.0;
  int64_t n_scored = 0;
  for (int64_t i = 0; i < n; i++) {
    if (((ots[i] ^ guide) & proto) == 0) { /* :67 first 20 bases equal -> excluded */
      if (per_ot) per_ot[i] = NAN;
      continue;
    }
    double pam = FF_CFD_PAM[base_at(ots[i], 23, 21)][base_at(ots[i], 23, 22)]; /* :69 last two bases */
    double s = pam * ffo_cfd_pair(guide, ots[i]);                               /* :71-73 */
    if (per_ot) per_ot[i] = s;
    int count = (int)(int16_t)(ots[i] >> 48);
    sum += s * (double)count; /* :79 scores.map(score*count).sum, left to right from 0.0 */
    if (n_scored == 0 || s > mx) mx = s;
    n_scored++;
  }
  *spec_out = n_scored > 0 ? 1.0 / (1.0 + sum) : 1.0;
  if (n_scored == 0) mx = 0.0;
  *max_out = mx >= 0.023 ? mx : 0.0; /* :83-87 */
}

/* CrisprMitEduOffTarget.scala:107-148 */
double ffo_hsu_offtarget(uint64_t guide, uint64_t ot) {
  int mm = 0, last = -1, dist_sum = 0, n_dist = 0;
  double p1 = 1.0;
  for (int i = 0; i < 20; i++) {
    if (base_at(ot, 23, i) != base_at(guide, 23, i)) {
      p1 = p1 * (1.0 - FF_HSU_COEF[i]);
      mm++;
      if (last >= 0) { dist_sum += i - last; n_dist++; }
      last = i;
    }
  }
  double p2 = 1.0;
  if (mm >= 2) {
    double avg = (double)dist_sum / (double)n_dist;
    p2 = 1.0 / ((((19 - avg) / 19.0) * 4.0) + 1.0);
  }
  double p3 = mm == 0 ? 1.0 : 1.0 / pow((double)mm, 2);
  double total = p1 * p2 * p3 * 100.0;
  return total * FF_HSU_PAM[base_at(ot, 23, 21)][base_at(ot, 23, 22)];
}

/* CrisprMitEduOffTarget.scala:60,85-105 */
double ffo_hsu_guide(const ffo_pack *p, uint64_t guide, const uint64_t *ots, int64_t n) {
  double sum = 0.0;
  for (int64_t i = 0; i < n; i++)
    if (mismatches_raw(p->cmp_mask, guide, ots[i], STRING_MASK) != 0) sum += ffo_hsu_offtarget(guide, ots[i]); /* :90 */
  return (100.0 / (100.0 + sum)) * 100.0;
}

/* ================================================================================================ */
/* EXTENSION: 1-bp bulge mode.  Not in the reference (PARITY UNPINNED); semantics defined in ff_oracle.h.
 * Deliberately written base by base (no bit tricks) so that it shares nothing with the CUDA implementation. */
void ffo_bulge_align(uint64_t guide, uint64_t target, int flags, int *mm_out, int *type_out, int *pos_out) {
  int g[20], t[20];
  for (int j = 0; j < 20; j++) { g[j] = base_at(guide, 23, j); t[j] = base_at(target, 23, j); }
  int best = 0, btype = 0, bpos = 0;
  for (int j = 0; j < 20; j++) best += g[j] != t[j];
  if (flags & FFO_BULGE_RNA)
    for (int q = 1; q <= 18; q++) {
      int mm = 0;
      for (int j = 1; j < 20; j++) mm += t[j] != (j <= q ? g[j - 1] : g[j]);
      if (mm < best) { best = mm; btype = 1; bpos = q; }
    }
  if (flags & FFO_BULGE_DNA)
    for (int q = 1; q <= 18; q++) {
      int mm = 0;
      for (int j = 0; j < 20; j++) if (j != q) mm += t[j] != (j < q ? g[j + 1] : g[j]);
      if (mm < best) { best = mm; btype = 2; bpos = q; }
    }
  *mm_out = best; *type_out = btype; *pos_out = bpos;
}

int ffo_discover_bulge(const ffo_pack *p, const uint64_t *targets, int64_t n_targets, const uint64_t *guides,
                       int64_t n_guides, int max_mismatch, int max_off_targets, int flags, int n_threads,
                       ffo_hits **out) {
  if (p->scan_len != 23 || p->five_prime) return -1;
  ffo_hits *h = (ffo_hits *)calloc(1, sizeof(ffo_hits));
  h->n_guides = n_guides;
  h->row_ptr = (int64_t *)calloc((size_t)n_guides + 1, 8);
  h->total_count = (int32_t *)calloc((size_t)n_guides + 1, 4);
  h->overflowed = (uint8_t *)calloc((size_t)n_guides + 1, 1);
  /* per guide: hit lists gathered independently (guides never interact), then concatenated */
  int64_t **idx = (int64_t **)calloc((size_t)n_guides + 1, sizeof(int64_t *));
  uint8_t **info = (uint8_t **)calloc((size_t)n_guides + 1, sizeof(uint8_t *));
  int64_t *cnt = (int64_t *)calloc((size_t)n_guides + 1, 8);
  (void)n_threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads > 0 ? n_threads : 1)
#endif
  for (int64_t g = 0; g < n_guides; g++) {
    int64_t cap = 64, n = 0;
    int64_t *ix = (int64_t *)malloc((size_t)cap * 8);
    uint8_t *inf = (uint8_t *)malloc((size_t)cap * 2);
    int32_t total = 0;
    for (int64_t i = 0; i < n_targets && total < max_off_targets; i++) { /* append while currentTotal < overflow */
      int mm, type, pos;
      ffo_bulge_align(guides[g], targets[i], flags, &mm, &type, &pos);
      if (mm > max_mismatch) continue;
      if (n == cap) { cap *= 2; ix = (int64_t *)realloc(ix, (size_t)cap * 8); inf = (uint8_t *)realloc(inf, (size_t)cap * 2); }
      ix[n] = i; inf[2 * n] = (uint8_t)mm; inf[2 * n + 1] = (uint8_t)(type ? ((type == 1 ? 0x40 : 0x80) | pos) : 0);
      n++;
      total += (int)(int16_t)(targets[i] >> 48);
    }
    idx[g] = ix; info[g] = inf; cnt[g] = n;
    h->total_count[g] = total;
    h->overflowed[g] = total >= max_off_targets;
  }
  int64_t nh = 0;
  for (int64_t g = 0; g < n_guides; g++) { h->row_ptr[g] = nh; nh += cnt[g]; }
  h->row_ptr[n_guides] = nh;
  h->targets = (uint64_t *)malloc((size_t)(nh + 1) * 8);
  h->mismatches = (uint8_t *)malloc((size_t)nh + 1);
  h->bulge = (uint8_t *)malloc((size_t)nh + 1);
  for (int64_t g = 0, k = 0; g < n_guides; g++) {
    for (int64_t i = 0; i < cnt[g]; i++, k++) {
      h->targets[k] = targets[idx[g][i]];
      h->mismatches[k] = info[g][2 * i];
      h->bulge[k] = info[g][2 * i + 1];
    }
    free(idx[g]); free(info[g]);
  }
  free(idx); free(info); free(cnt);
  h->n_targets_scanned = (uint64_t)n_targets;
  *out = h;
  return 0;
}
