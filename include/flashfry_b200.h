/*
 * flashfry_b200.h -- C ABI of libflashfry_b200.so: FlashFry's off-target discovery + CFD / Hsu2013 scoring
 * hot path on NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary.  The reference (mckennalab/FlashFry, Scala/JVM) has no FFI; the two seams
 * this ABI replaces are
 *   - trait Traverser.scan            src/main/scala/reference/traverser/Traverser.scala:52-59
 *       (SeekTraverser.scan  reference/traverser/SeekTraverser.scala:58-104,
 *        LinearTraverser.scan reference/traverser/LinearTraverser.scala:59-112; called from
 *        modules/OffTargetDiscovery.scala:124,130)
 *   - trait ScoreModel.scoreGuides    src/main/scala/scoring/ScoreModel.scala:31-89,113-132
 *       (Doench2016CFDScore.scoreGuide scoring/Doench2016CFDScore.scala:53-88,
 *        CrisprMitEduOffTarget.score_crispr scoring/CrisprMitEduOffTarget.scala:60-148; selected in
 *        modules/ScoreResults.scala:159-226)
 * INTEGRATION.md shows the JNI stub + Scala glue a FlashFry maintainer would add on top of these symbols.
 *
 * Conventions: every call returns 0 on success or a negative FF_E* code; ff_last_error() gives the message of
 * the last failure on the calling thread.  No exceptions cross the boundary.  Plain pointers and sizes only.
 * All pointers are HOST pointers unless the parameter name starts with d_ (device).  One ff_ctx is bound to one
 * GPU and may be used by one host thread at a time; contexts are independent (one per GPU / per rank).
 * There is no CPU fallback: without a CUDA device ff_create fails with FF_ENODEVICE.
 */
#ifndef FLASHFRY_B200_H
#define FLASHFRY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FF_OK 0
#define FF_EINVAL (-1)    /* bad argument (the reference would fail a require/assert) */
#define FF_ENODEVICE (-2) /* no usable CUDA device */
#define FF_ECUDA (-3)     /* CUDA runtime error, see ff_last_error() */
#define FF_EIO (-4)       /* cannot read / parse the database or header */
#define FF_EFORMAT (-5)   /* bad magic/version (BinaryHeader.scala:121-124) or invalid bin type (BlockManager.scala:85-87) */
#define FF_ENODB (-6)     /* no database resident in the context */
#define FF_EUNSUPPORTED (-7)
#define FF_ENOMEM (-8)

#define FF_METRIC_CFD 1u      /* Doench2016CFDScore: DoenchCFD_maxOT, DoenchCFD_specificityscore */
#define FF_METRIC_HSU2013 2u  /* CrisprMitEduOffTarget: Hsu2013 */

#define FF_BULGE_RNA 1 /* ff_discover_bulge: allow one looped-out guide base */
#define FF_BULGE_DNA 2 /* ff_discover_bulge: allow one looped-out genomic base */

typedef struct ff_ctx ff_ctx;

/* ---- context ------------------------------------------------------------------------------------------- */
int ff_create(ff_ctx **out, int device_id);
void ff_destroy(ff_ctx *ctx);
const char *ff_last_error(void);
int ff_abi_version(void);
/* Run the context's work on a caller-owned CUDA stream (a cudaStream_t passed as void*; NULL = the context's own
 * stream).  Lets a host runtime (torch, JCuda ...) order/time the library's kernels with its own events. */
int ff_set_stream(ff_ctx *ctx, void *cuda_stream);

/* Tuning and test knobs of one context (the library reads no environment variables).  Keys:
 *   scan_kernel   0 = chosen from batch and database size (default), 1 = guide-major k_seed_scan, 2 = bin-major
 *                 k_bin_scan + k_pair_scan (falls back to 1 where the bin scan does not apply)
 *   force_general 1 = run the mismatch-only search through the windowed general path (early termination of full guides)
 *   window_cells  > 0 = fixed window size, in 1/64ths of the database, of the general path
 *   subbatch_min, subbatch_c1, subbatch_c2   how ff_discover cuts a guide set into sub-batches (D2H overlap)
 *   group_sort    0 = always order candidate hits with the radix sort
 *   b_spi         > 0 = seeds per work item of the part-two pass of the guide-major kernels
 *   trace         1 = ff_discover prints host-side timestamps of its sub-batches to stderr
 *   split_a       > 0 = bases in the part-one key, applied by the next database load
 *   compact_hits  1 = ff_discover ships database indices (ff_hits.target_index) and leaves ff_hits.targets NULL until
 *                 ff_hits_resolve fills it from the host mirror of the target array
 * Unknown key or out-of-range value: FF_EINVAL. */
int ff_set_option(ff_ctx *ctx, const char *key, long long value);

/* ---- database (replaces BinaryHeader.readHeader + the BGZF seek/inflate of SeekTraverser.fillBlock) ------ */
/* Reads FlashFry's own on-disk format unchanged: BGZF body (reference/binary/DatabaseWriter.scala:78-97, both
 * block types of reference/binary/blocks/BlockManager.scala:362-442) + text side-car
 * (reference/binary/BinaryHeader.scala:115-160); inflates on host threads and makes it resident in HBM. */
int ff_load_database(ff_ctx *ctx, const char *db_path, const char *header_path);
/* Database image side-car (not a FlashFry format; a cache of the decoded database): ff_save_image writes the resident
 * database (targets, positions, contig names) as flat little-endian arrays, ff_load_image makes it resident again with
 * one sequential read + H2D, skipping the BGZF inflate and block walk of ff_load_database.  The index build
 * re-validates the content (FF_EFORMAT on a damaged image). */
int ff_save_image(ff_ctx *ctx, const char *image_path);
int ff_load_image(ff_ctx *ctx, const char *image_path);
/* Same residency from caller-provided arrays in database order (targets carry their 16-bit count; positions may be
 * NULL).  contigs may be NULL.  Used by tests and by hosts that already hold the decoded blocks. */
int ff_load_database_arrays(ff_ctx *ctx, int enzyme_index, int bin_width, const uint64_t *targets,
                            uint64_t n_targets, const uint64_t *positions, uint64_t n_positions,
                            const char *const *contigs, int n_contigs);
/* Bench support: generate a synthetic spCas9-family database of ~n_targets distinct sorted targets directly in HBM
 * (uniform random protospacer+N, PAM per enzyme, count model of SURVEY.md 8(d) cfg 3).  No positions. */
int ff_synth_database(ff_ctx *ctx, int enzyme_index, uint64_t n_targets, uint64_t seed);
/* The same with sequence-level skew: n_families repeat neighbourhoods of family_size targets each, every member its
 * family's consensus with 0..family_subs random substitutions (Alu-like: hot index buckets, guides with thousands of
 * candidates).  Members replace uniform targets, duplicates collapse as usual. */
int ff_synth_database_skewed(ff_ctx *ctx, int enzyme_index, uint64_t n_targets, uint64_t seed, uint64_t n_families,
                             uint64_t family_size, int family_subs);

typedef struct {
  int enzyme_index;      /* standards/StandardScanParameters.scala:61-80 */
  int bin_width;         /* header bin count = 4^bin_width (BinaryHeader.scala:131-132) */
  int scan_len;          /* totalScanLength */
  int pam_len;
  int five_prime_pam;
  uint64_t cmp_mask;     /* comparisonBitEncoding */
  uint64_t n_targets;
  uint64_t n_positions;  /* 0 when the image has no positions */
  int n_contigs;
  int seed_split_a;      /* the device seed index splits the protospacer into first a | last P-a bases */
  uint64_t device_bytes; /* HBM held by the resident image */
} ff_db_info_t;
int ff_db_info(const ff_ctx *ctx, ff_db_info_t *out);
const char *ff_db_contig(const ff_ctx *ctx, int contig_id /* 1-based, BitPosition.scala:37-44 */);
/* Copy targets[first, first+n) of the resident image back to the host (test/bench cross-checks). */
int ff_db_copy_targets(ff_ctx *ctx, uint64_t first, uint64_t n, uint64_t *out);

/* ---- discover (replaces Traverser.scan) ----------------------------------------------------------------- */
/* CSR over guides, library-owned host memory (pinned), free with ff_hits_free.  Row g holds what the reference
 * leaves in aggregator.wrappedGuides(g).otSite.offTargets: CRISPRHit(target long, positions) in DATABASE ORDER,
 * truncated by the overflow rule "append while currentTotal < maximumOffTargets; currentTotal += count"
 * (crispr/ResultsAggregator.scala:61-69, crispr/CRISPRSiteOT.scala:39-46). */
typedef struct ff_hits {
  int64_t n_guides;
  int64_t n_hits;
  const int64_t *row_ptr;      /* [n_guides+1] */
  const uint64_t *targets;     /* [n_hits] target long incl. 16-bit count */
  const uint8_t *mismatches;   /* [n_hits] == BitEncoding.mismatches(guide, target) (bitcoding/BitEncoding.scala:127-132) */
  const int64_t *pos_ptr;      /* [n_hits+1] or NULL when positions were not requested / not resident */
  const uint64_t *positions;   /* BitPosition longs (bitcoding/BitPosition.scala:51-92) */
  const int32_t *total_count;  /* [n_guides] CRISPRSiteOT.currentTotal */
  const uint8_t *overflowed;   /* [n_guides] CRISPRSiteOT.full -> "OVERFLOW" in the TSV */
  uint64_t n_compares;         /* guide x target comparisons the kernels performed (for the log line of
                                  modules/OffTargetDiscovery.scala:137; NOT equal to the reference's count: pruning differs) */
  uint64_t n_candidate_hits;   /* hits before the overflow cut */
  void *opaque;
  const uint8_t *bulge;        /* [n_hits], ff_discover_bulge only (else NULL): 0 = no bulge, 0x40|q = RNA bulge at guide
                                  base q, 0x80|q = DNA bulge at genomic base q; mismatches[] then counts the aligned pairs */
  const uint32_t *target_index; /* [n_hits] with option compact_hits = 1 (else NULL): the hit's index in database order;
                                  targets[i] == ff_db_host_targets(ctx)[target_index[i]].  targets stays NULL until
                                  ff_hits_resolve fills it */
} ff_hits;

/* Compact hit lists (option compact_hits): a hit's target long is 8 bytes, its database index 4; the device-to-host copy
 * of the hit list is the slowest part of a large discover call, so a host that can look the long up itself (the JVM glue
 * does, when it builds a CRISPRHit) asks for indices.  ff_db_host_targets returns a host mirror of the resident target
 * array (database order, made on first use, owned by the context; NULL on failure); ff_hits_resolve fills
 * hits->targets from it on all host threads. */
const uint64_t *ff_db_host_targets(ff_ctx *ctx);
int ff_hits_resolve(ff_ctx *ctx, ff_hits *hits);

/* guides: BitEncoding.bitEncodeString(bases, count = 1) longs in ResultsAggregator order (any order is accepted;
 * rows come back in the same order).  max_mismatch >= 0, max_off_targets >= 0. */
int ff_discover(ff_ctx *ctx, const uint64_t *guides, int64_t n_guides, int max_mismatch, int max_off_targets,
                int want_positions, ff_hits **out);
void ff_hits_free(ff_hits *hits);

/* ---- discover TSV fast path (replaces TabDelimitedOutput.write for `discover`) --------------------------------------
 * Writes FlashFry's discover TSV (targetio/TabDelimitedHandler.scala:119-154: header line, one row per guide --
 * contig, start, stop, target, context, overflow, orientation, otCount, offTargets -- tokens SEQ_count_mm joined by
 * commas, crispr/CRISPRHit.scala:54-88; with write_positions the <contig:start^F|...> suffix of :75-81) straight from the
 * CSR, without building a CRISPRHit per hit (1e7 JVM objects for 100 000 guides) and without `score` re-parsing text it
 * could have had in memory.  guides[g] describes row g (what CRISPRSite holds); guide_longs[g] its encoding.  Works on
 * compact hit lists too (option compact_hits: target longs are looked up in the host mirror).  Bulge hit lists add the
 * _R<q> / _D<q> field of the CLI's --bulge extension. */
typedef struct {
  const char *contig;
  int32_t start;        /* CRISPRSite.position; stop = start + strlen(bases) */
  const char *bases;    /* the guide incl. PAM, as found in the FASTA */
  const char *context;  /* flanked sequence, or NULL for "NONE" */
  int32_t forward;      /* 1 = FWD, 0 = RVS */
} ff_tsv_guide;
int ff_hits_write_tsv(ff_ctx *ctx, const char *path, const ff_tsv_guide *guides, const ff_hits *hits, int write_positions);

/* ---- EXTENSION: 1-bp bulge mode (BASELINE.json configs[3]) --------------------------------------------------
 * NOT a replacement of anything: the reference has no gap / bulge / edit-distance search (no match for
 * bulge|gap|indel|levenshtein under src/main), so there is no reference behaviour to be bit-exact with.  Semantics
 * (the tests hold a brute-force, base-by-base statement of the same definition): for 23-bp Cas9 packs (20-base protospacer g / t,
 * base 0 PAM-distal, 3' PAM), with q in 1..18,
 *   no bulge       mm = #{ j : g[j] != t[j] }                                       (= BitEncoding.mismatches)
 *   RNA bulge at q guide base q looped out: mm = hamming(g[0..q) + g(q..19], t[1..19]); t[0] is outside the alignment
 *   DNA bulge at q genomic base q looped out: mm = hamming(g[1..19], t[0..q) + t(q..19]); g[0] would pair with the base
 *                  upstream of the stored 23-mer, which a FlashFry database does not hold, so it is not scored
 * A target is a hit when its best allowed alignment -- smallest (mm, type none < RNA < DNA, q) -- has
 * mm <= max_mismatch.  Rows are in database order and cut by the reference's overflow rule, exactly like ff_discover;
 * bulge_flags = 0 gives ff_discover's results (with hits->bulge all zero).  Other enzymes: FF_EUNSUPPORTED. */
int ff_discover_bulge(ff_ctx *ctx, const uint64_t *guides, int64_t n_guides, int max_mismatch, int max_off_targets,
                      int bulge_flags, int want_positions, ff_hits **out);

/* ---- score (replaces ScoreModel.scoreGuides for Doench2016CFDScore and CrisprMitEduOffTarget) ------------ */
/* hits: row g = the off-targets of guides[g] (from ff_discover or re-read from a discover TSV; only n_guides,
 * row_ptr and targets are read).  Outputs are [n_guides] (per_ot_cfd: [n_hits], NaN for on-target copies that
 * Doench2016CFDScore.scala:67 skips); any output pointer may be NULL.  cfd_max already carries the 0.023
 * threshold of :83-87.  Valid only for 23-bp Cas9 packs (validOverEnzyme); otherwise FF_EUNSUPPORTED and the
 * host prints "NA" (scoring/ScoreModel.scala:125-128). */
int ff_score(ff_ctx *ctx, const uint64_t *guides, const ff_hits *hits, uint32_t metrics, double *cfd_max,
             double *cfd_specificity, double *hsu2013, double *per_ot_cfd);

/* The same with the enzyme stated by the caller (FlashFry's standalone `score` knows it from the --database header):
 * ff_score uses the resident database's enzyme, or spCas9-NGG when no database is resident. */
int ff_score_enzyme(ff_ctx *ctx, int enzyme_index, const uint64_t *guides, const ff_hits *hits, uint32_t metrics,
                    double *cfd_max, double *cfd_specificity, double *hsu2013, double *per_ot_cfd);

/* Integer aggregates over the same hit lists: scoring/ClosestHit.scala:43-76 ("minot": smallest non-zero mismatch
 * count, the summed occurrence count at that distance, the 0..4-mismatch occurrence histogram) and the in-genome count
 * of scoring/DangerousSequences.scala:61-65 (occurrences with zero mismatches).  Outputs are [n_guides]
 * (hist: [n_guides][5]); closest is INT32_MAX where the reference prints "UNK".  Any output may be NULL. */
int ff_hit_aggregates(ff_ctx *ctx, int enzyme_index, const uint64_t *guides, const ff_hits *hits, int32_t *closest,
                      int32_t *closest_count, int32_t *hist, int32_t *in_genome);

/* Fused discover + score: the hit list is scored while still in HBM, then both come back in one D2H. */
int ff_discover_score(ff_ctx *ctx, const uint64_t *guides, int64_t n_guides, int max_mismatch, int max_off_targets,
                      int want_positions, uint32_t metrics, ff_hits **out, double *cfd_max,
                      double *cfd_specificity, double *hsu2013);

/* ---- device-resident path (kernel-only timing, multi-GPU plumbing) --------------------------------------- */
typedef struct {
  int64_t n_guides;
  int64_t n_hits;             /* after the overflow cut */
  uint64_t n_candidate_hits;
  uint64_t n_compares;
  const int64_t *d_row_ptr;   /* device pointers owned by the context, valid until its next discover call */
  const uint64_t *d_targets;
  const uint8_t *d_mismatches;
  const int32_t *d_total_count;
  const uint8_t *d_overflowed;
  const double *d_cfd_max, *d_cfd_specificity, *d_hsu2013; /* NULL unless metrics requested */
  const uint8_t *d_bulge;     /* ff_discover_bulge_device only, else NULL */
} ff_device_result;
/* d_guides already in HBM; results stay in HBM.  metrics may be 0. */
int ff_discover_device(ff_ctx *ctx, const uint64_t *d_guides, int64_t n_guides, int max_mismatch,
                       int max_off_targets, uint32_t metrics, ff_device_result *out);

/* The bulge extension on device-resident guides (no scoring: CFD / Hsu2013 are not defined for bulged alignments). */
int ff_discover_bulge_device(ff_ctx *ctx, const uint64_t *d_guides, int64_t n_guides, int max_mismatch,
                             int max_off_targets, int bulge_flags, ff_device_result *out);

/* ---- database-sharded discover over NVLink peer memory ------------------------------------------------------
 * Strong scaling of ONE guide set over the GPUs of a box (reference/traverser/Traverser.scala:52-59 hands a traverser all
 * guides and the whole database; the reference itself is single-process, OffTargetDiscovery.scala:117).  Every rank holds
 * an index replica, scans 1/world of the INDEX for ALL guides, and the scan kernels push each candidate straight into the
 * exchange block of the rank that owns the guide (ff_shard_range) -- fire-and-forget P2P stores into the region the owner
 * keeps for each source rank (positions from local atomics); the barriers (one remote atomic per peer) and the all-gather
 * of the per-guide totals go through the same blocks: no NCCL on this path.  Rows are identical to ff_discover's.
 *   1. every rank: ff_peer_export        -> its block exists; handle = CUDA IPC handle (other processes), *block_out = pointer
 *   2. every rank: ff_peer_attach        with the handles of all ranks (one process per GPU) or their pointers (one process)
 *   3. every rank, same arguments:        ff_discover_sharded[_device](all guides) -> rows of ITS guides
 * A call that fails on one rank makes the others return FF_ECUDA after a 4 s barrier time-out instead of hanging.  Every
 * ff_peer_attach needs a fresh ff_peer_export on all ranks (the barrier counters live in the blocks and start at 0). */
#define FF_PEER_HANDLE_BYTES 64
/* (hit_cap and guide_cap must be the same on every rank: they fix the layout of the blocks the peers write into) */
int ff_peer_export(ff_ctx *ctx, uint64_t hit_cap /* candidate keys per block; 0 = 2^24 */, int64_t guide_cap /* all guides; 0 = 2^20 */,
                   void *handle_out /* FF_PEER_HANDLE_BYTES */, void **block_out /* may be NULL */);
int ff_peer_attach(ff_ctx *ctx, int rank, int world, const void *handles /* world x FF_PEER_HANDLE_BYTES, or NULL */,
                   void *const *blocks /* world device pointers of this process, or NULL */);
int ff_peer_detach(ff_ctx *ctx);
/* d_guides_all: ALL guides (identical on every rank), already in HBM.  out: rows of guides [first, first + count) of this
 * rank (ff_shard_range(n_guides_all, world, rank)); d_total_count = this rank's slice. */
int ff_discover_sharded_device(ff_ctx *ctx, const uint64_t *d_guides_all, int64_t n_guides_all, int max_mismatch,
                               int max_off_targets, uint32_t metrics, ff_device_result *out);
/* the same with host buffers: H2D of all guides, D2H of this rank's rows (ff_hits of its guides) */
int ff_discover_sharded(ff_ctx *ctx, const uint64_t *guides_all, int64_t n_guides_all, int max_mismatch, int max_off_targets,
                        ff_hits **out);
/* after a sharded call: device pointer to CRISPRSiteOT.currentTotal of ALL guides (int32[n_guides_all]), on every rank */
const int32_t *ff_peer_totals_device(ff_ctx *ctx);

/* ---- several GPUs behind one host process ------------------------------------------------------------------
 * The reference is a single process (modules/OffTargetDiscovery.scala:117: "multithreaded (not supported currently)"); a
 * JVM host reaches every GPU of a box through one ff_multi: one ff_ctx and one persistent host thread per device, every
 * device holds a replica of the database, discover shards the guide array (rank r: ff_shard_range), and ONE
 * ncclAllGather of the per-guide occurrence totals at the end leaves the global vector on every device and on the host.
 * NCCL is loaded at run time (dlopen) only when n_devices > 1. */
typedef struct ff_multi ff_multi;
int ff_multi_create(ff_multi **out, const int *device_ids, int n_devices);
void ff_multi_destroy(ff_multi *m);
int ff_multi_size(const ff_multi *m);
ff_ctx *ff_multi_ctx(ff_multi *m, int rank);  /* the rank's context: ff_db_info, ff_last_timings ... (not while a multi call runs) */
int ff_multi_set_option(ff_multi *m, const char *key, long long value);
/* FlashFry's files are read and inflated once; the decoded arrays reach the other devices with ncclBroadcast (NVLink). */
int ff_multi_load_database(ff_multi *m, const char *db_path, const char *header_path);
int ff_multi_synth_database(ff_multi *m, int enzyme_index, uint64_t n_targets, uint64_t seed);
/* guides [first, first + count) belong to shard `shard` of n_shards: first = n_guides * shard / n_shards (pure function) */
void ff_shard_range(int64_t n_guides, int n_shards, int shard, int64_t *first, int64_t *count);
/* out: [n_devices] hit lists, out[r] = rows of shard r's guides (free each with ff_hits_free); total_count_all:
 * [n_guides] (may be NULL) = the all-gathered CRISPRSiteOT.currentTotal of every guide, in guide order. */
int ff_multi_discover(ff_multi *m, const uint64_t *guides, int64_t n_guides, int max_mismatch, int max_off_targets,
                      int want_positions, ff_hits **out, int32_t *total_count_all);
/* device pointer (on rank's GPU) to the gathered totals of the last ff_multi_discover: n_devices shards padded to
 * ceil(n_guides / n_devices) int32 each; NULL for a single device */
const int32_t *ff_multi_device_totals(ff_multi *m, int rank);

/* Device-time of the kernels of the last discover call on this context, from CUDA events recorded on the
 * context's stream (ms).  scan = the bin-scan kernel (dominant), prep = guide sort/bucketing, order = hit sort,
 * cut = overflow cut + gather, score = CFD/Hsu kernels, total = first to last event. */
typedef struct {
  float prep_ms, scan_ms, order_ms, cut_ms, score_ms, total_ms;
  int scan_launches;          /* >1 when the hit buffer had to grow and the scan was repeated */
  int kernel_launches;        /* kernels of this library launched by the call */
  uint64_t scan_bytes_read;   /* bytes the scan kernels REQUEST for this call (index entries + streamed planes + guides + hit
                                 keys) -- not the roofline's algorithmic bytes, which bench.py computes per SURVEY.md 8(d) */
  float scan_part1_ms, scan_part2_ms; /* bin scan: k_bin_scan (index A) / k_pair_scan (index B); else scan_ms / 0 */
  uint64_t entries_part1, entries_part2; /* index entries compared by each */
} ff_timings;
int ff_last_timings(const ff_ctx *ctx, ff_timings *out);

#ifdef __cplusplus
}
#endif
#endif /* FLASHFRY_B200_H */
