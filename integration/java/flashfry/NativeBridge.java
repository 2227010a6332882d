// JNI declarations for libflashfry_b200 (see INTEGRATION.md).  Not compiled in this repository's image (no JDK); the C side
// (integration/jni/flashfry_b200_jni.c) is executed against a functional mock JNIEnv by tests/test_gpu_jni.py.
package flashfry;

public final class NativeBridge {
  static { System.loadLibrary("flashfry_b200_jni"); }   // the shim, linked against libflashfry_b200.so

  /** ff_create / ff_destroy: returns the ff_ctx* as a long handle, throws IllegalStateException on failure. */
  public static native long create(int deviceId);
  public static native void destroy(long ctx);
  /** ff_set_option, e.g. ("compact_hits", 1) */
  public static native void setOption(long ctx, String key, long value);

  /** ff_load_database(ctx, dbPath, dbPath + ".header"); ff_load_image / ff_save_image: the decoded side-car */
  public static native void loadDatabase(long ctx, String dbPath);
  public static native void loadImage(long ctx, String imagePath);
  public static native void saveImage(long ctx, String imagePath);

  /** ff_discover.  guides = BitEncoding.bitEncodeString(bases, 1) in ResultsAggregator order.
   *  Returns the ff_hits* handle; read it with the accessors, then hitsFree. */
  public static native long discover(long ctx, long[] guides, int maxMismatch, int maxOffTargets, boolean wantPositions);
  /** ff_discover_score: the hit list is scored while still in HBM; cfdMax / cfdSpecificity / hsu2013 are filled in place. */
  public static native long discoverScore(long ctx, long[] guides, int maxMismatch, int maxOffTargets, boolean wantPositions,
                                          int metrics, double[] cfdMax, double[] cfdSpecificity, double[] hsu2013);
  public static native long[] hitsRowPtr(long hits);      // [nGuides + 1]
  public static native long[] hitsTargets(long hits);     // target longs incl. the 16-bit count, database order
  public static native byte[] hitsMismatches(long hits);  // == BitEncoding.mismatches(guide, target)
  public static native int[]  hitsTargetIndex(long hits); // option compact_hits: index into dbHostTargets (else null)
  public static native void   hitsResolve(long ctx, long hits);            // fill targets of a compact hit list
  public static native java.nio.ByteBuffer dbHostTargets(long ctx);        // .order(LITTLE_ENDIAN).asLongBuffer(): the target array
  public static native int[]  hitsTotalCount(long hits);  // CRISPRSiteOT.currentTotal per guide
  public static native byte[] hitsOverflowed(long hits);  // CRISPRSiteOT.full per guide
  public static native long[] hitsPosPtr(long hits);      // null unless wantPositions
  public static native long[] hitsPositions(long hits);   // BitPosition longs
  public static native long   hitsCompares(long hits);
  public static native void   hitsFree(long hits);
  /** ff_hits_write_tsv: FlashFry's discover TSV straight from the CSR (contexts[i] == null prints "NONE") */
  public static native void hitsWriteTsv(long ctx, long hits, String path, String[] contigs, int[] starts, String[] bases,
                                         String[] contexts, boolean[] forward, boolean positions);

  /** ff_score_enzyme over guides + CSR hit list; out = {cfdMax[n], cfdSpecificity[n], hsu2013[n], perOtCfd[nHits]} (NaN = "not scored") */
  public static native double[][] score(long ctx, int enzymeIndex, long[] guides, long[] rowPtr, long[] targets, int metrics);
  /** ff_hit_aggregates: int[n * 8] = {closest (Integer.MAX_VALUE = "UNK"), closestCount, hist0..hist4, inGenome} per guide */
  public static native int[] hitAggregates(long ctx, int enzymeIndex, long[] guides, long[] rowPtr, long[] targets);

  /** ff_multi: every GPU of the box behind this one process; multiDiscover returns one ff_hits* per device (shard r =
   *  guides [r n / d, (r + 1) n / d)) and fills totals[nGuides] with the NCCL-all-gathered per-guide totals. */
  public static native long multiCreate(int[] devices);
  public static native void multiDestroy(long multi);
  /** ff_multi_set_option: every context option, plus "shard_mode" (0 = shard the guides, 1 = shard the index work: the
   *  candidates reach the guide's owner over NVLink peer memory) and "peer_hit_cap". */
  public static native void multiSetOption(long multi, String key, long value);
  public static native void multiLoadDatabase(long multi, String dbPath);
  public static native long[] multiDiscover(long multi, long[] guides, int maxMismatch, int maxOffTargets, boolean wantPositions, int[] totals);
}
