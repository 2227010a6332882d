/*
 * ff_oracle.h -- CPU restatement of FlashFry's off-target discovery + CFD/Hsu2013 scoring path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (flashfry_b200/, include/) may include, link or
 * call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, as the checker / CPU baseline.
 *
 * Parity pin status: PINNED.  The restatement reproduces the reference's own integration-test md5s
 * (test_data/integration_test.sh:81,84) and unit-test known answers; see tests/test_oracle_pins.py.
 *
 * Every function names the reference file:line (relative to the FlashFry checkout, v1.15) it follows.
 */
#ifndef FF_ORACLE_H
#define FF_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* standards/StandardScanParameters.scala:28-48 (ParameterPack), the fields the hot path reads */
typedef struct {
  int enzyme_index;   /* :61-80  1 Cpf1, 2 SpCas9, 3 SpCas9NGG, 4 SpCas9NAG, 5 SpCas9-19, 6 SpCas9NGG-19 */
  int scan_len;       /* totalScanLength */
  int pam_len;        /* pamLength */
  int five_prime;     /* fivePrimePam */
  uint64_t cmp_mask;  /* comparisonBitEncoding */
} ffo_pack;

/* StandardScanParameters.scala:61-70 indexToParameterPack; returns 0 or -1 for an unknown index */
int ffo_pack_from_index(int enzyme_index, ffo_pack *out);

/* bitcoding/BitEncoding.scala:46-67 bitEncodeString; returns 0 and sets *err=1 on a bad base/length/count */
uint64_t ffo_encode(const char *bases, int len, int count, int *err);
/* BitEncoding.scala:85-99 bitDecodeString (out must hold len+1 chars); returns the count */
int ffo_decode(uint64_t enc, int len, char *out);
/* BitEncoding.scala:127-132 mismatches(encoding1, encoding2, additionalMask) */
int ffo_mismatches(const ffo_pack *p, uint64_t a, uint64_t b, uint64_t additional_mask);
/* BitEncoding.scala:153-185 binToLongComparitor / compBitmaskForBin / binShift.
 * bin_code = base-4 value of the bin string (A=0..T=3, first base most significant). */
void ffo_bin_comparitor(const ffo_pack *p, uint64_t bin_code, int bin_size, int right_shift_bases,
                        uint64_t *bin_long, uint64_t *guide_mask);
/* BitEncoding.scala:142-144 mismatchBin */
int ffo_mismatch_bin(const ffo_pack *p, uint64_t bin_long, uint64_t guide_mask, uint64_t guide);

/* Result of a discover run, CSR over guides (guide order = caller's order = ResultsAggregator order). */
typedef struct {
  int64_t n_guides;
  int64_t *row_ptr;        /* [n_guides+1] */
  uint64_t *targets;       /* target long incl. count, in database order, overflow-truncated */
  uint8_t *mismatches;     /* BitEncoding.mismatches(guide, target) */
  int64_t *pos_ptr;        /* [n_hits+1] into positions (NULL when the DB image has no positions) */
  uint64_t *positions;
  int32_t *total_count;    /* CRISPRSiteOT.currentTotal */
  uint8_t *overflowed;     /* CRISPRSiteOT.full */
  uint64_t n_compares;     /* Traverser.allComparisons (target x guide) + sub-bin prefix compares + set-up loop */
  uint64_t n_target_compares; /* BlockManager.scala:246 only */
  uint64_t n_targets_scanned;
  int saturated;           /* OrderedBinTraversalFactory.saturated -> linear traversal taken */
  int bins_visited;
  uint8_t *bulge;          /* only from ffo_discover_bulge (extension, see below): 0 = no bulge, 0x40|q RNA, 0x80|q DNA */
} ffo_hits;

void ffo_hits_free(ffo_hits *h);

/*
 * discover over the reference's own block format (a5-a9 of SURVEY section 8):
 *   db          all inflated bin blocks, concatenated in bin order (native-endian longs)
 *   bin_off     [n_bins+1] offsets, in longs, of each bin's block inside db
 * Follows OrderedBinTraversalFactory.scala:146-177 (precompute + saturation), LinearTraversal.scala:82-97,
 * SeekTraverser.scala:78-102, BlockManager.scala:63-90,143-201,212-254, ResultsAggregator.scala:61-69,
 * CRISPRSiteOT.scala:39-46.
 */
int ffo_discover_blocks(const ffo_pack *p, int bin_width, const uint64_t *db, const int64_t *bin_off,
                        const uint64_t *guides, int64_t n_guides, int max_mismatch, int max_off_targets,
                        int force_linear, ffo_hits **out);

/*
 * Same loop order over a positions-free SoA image (targets only, database order, bin_off in targets),
 * for bench-scale databases whose block form would not fit: a bin with >500 targets is walked as an
 * indexed block (256 sub-bin filters, BlockManager.scala:143-201), otherwise linearly (:212-254).
 * n_threads > 1 farms bins over host threads and merges hits in bin order (same results).
 * Validated against ffo_discover_blocks in tests/test_oracle_pins.py.
 */
int ffo_discover_soa(const ffo_pack *p, int bin_width, const uint64_t *targets, const int64_t *bin_off,
                     const uint64_t *guides, int64_t n_guides, int max_mismatch, int max_off_targets,
                     int n_threads, ffo_hits **out);

/* scoring/Doench2016CFDScore.scala:132-151 scoreCFD over the 20 protospacer bases of two longs (23-mers) */
double ffo_cfd_pair(uint64_t guide, uint64_t off_target);
/* Doench2016CFDScore.scala:53-88 scoreGuide: returns max (thresholded, :83-87) and specificity;
 * per_ot (may be NULL) receives pam*score per hit, NaN for skipped on-target copies (:67) */
void ffo_cfd_guide(uint64_t guide, const uint64_t *ots, int64_t n, double *max_out, double *spec_out,
                   double *per_ot);
/* scoring/CrisprMitEduOffTarget.scala:107-148 scoreOffTarget */
double ffo_hsu_offtarget(uint64_t guide, uint64_t off_target);
/* CrisprMitEduOffTarget.scala:60,85-105 score_crispr */
double ffo_hsu_guide(const ffo_pack *p, uint64_t guide, const uint64_t *ots, int64_t n);

/* ---------------------------------------------------------------------------------------------------
 * EXTENSION -- 1-bp bulge mode (BASELINE.json configs[3]).  The reference has NO gap / bulge /
 * edit-distance mode: these functions restate nothing, they DEFINE the semantics the CUDA path is
 * tested against.  PARITY UNPINNED (there is no reference behaviour to pin to); never part of a parity claim.
 *
 * Definition (20-base protospacer packs with a 3' PAM only; g = guide protospacer, t = target protospacer, base 0
 * is PAM-distal, both PAM-anchored; q in 1..18):
 *   no bulge      : mm = #{ j in 0..19 : g[j] != t[j] }                               (= BitEncoding.mismatches)
 *   RNA bulge at q: guide base q is looped out; the other 19 guide bases pair with the 19 genomic bases next to the
 *                   PAM:  mm = hamming( g[0..q) + g(q..19] , t[1..19] );  t[0] lies outside the alignment
 *   DNA bulge at q: genomic base q is looped out; the other 19 stored genomic bases pair with g[1..19]:
 *                   mm = hamming( g[1..19] , t[0..q) + t(q..19] );  g[0] would pair with the genomic base upstream of
 *                   the stored 23-mer, which a FlashFry database does not hold, so it is not scored
 * A target is a hit when the best alignment allowed by `flags` (bit0 RNA, bit1 DNA; the no-bulge alignment always)
 * has mm <= max_mismatch; best = smallest (mm, type [none < RNA < DNA], q).  Hit order and the overflow rule are the
 * reference's (database order, ResultsAggregator.scala:61-69).
 */
#define FFO_BULGE_RNA 1
#define FFO_BULGE_DNA 2
/* best alignment of one pair; *type 0 none / 1 RNA / 2 DNA, *pos = q (0 when none) */
void ffo_bulge_align(uint64_t guide, uint64_t target, int flags, int *mm, int *type, int *pos);
/* brute force over all targets (database order); hits->bulge is filled */
int ffo_discover_bulge(const ffo_pack *p, const uint64_t *targets, int64_t n_targets, const uint64_t *guides,
                       int64_t n_guides, int max_mismatch, int max_off_targets, int flags, int n_threads,
                       ffo_hits **out);

#ifdef __cplusplus
}
#endif
#endif
