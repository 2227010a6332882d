// JNI declarations for libflashfry_b200 (see INTEGRATION.md).  Not compiled in this repository's image (no JDK).
package flashfry;

public final class NativeBridge {
  static { System.loadLibrary("flashfry_b200_jni"); }   // the shim below, linked against libflashfry_b200.so

  /** ff_create / ff_destroy: returns the ff_ctx* as a long handle, throws IllegalStateException on failure. */
  public static native long create(int deviceId);
  public static native void destroy(long ctx);

  /** ff_load_database(ctx, dbPath, dbPath + ".header") */
  public static native void loadDatabase(long ctx, String dbPath);

  /** ff_discover.  guides = BitEncoding.bitEncodeString(bases, 1) in ResultsAggregator order.
   *  Returns the ff_hits* handle; read it with the accessors, then hitsFree. */
  public static native long discover(long ctx, long[] guides, int maxMismatch, int maxOffTargets, boolean wantPositions);
  public static native long[] hitsRowPtr(long hits);      // [nGuides + 1]
  public static native long[] hitsTargets(long hits);     // target longs incl. the 16-bit count, database order
  public static native long[] hitsPosPtr(long hits);      // null unless wantPositions
  public static native long[] hitsPositions(long hits);   // BitPosition longs
  public static native long   hitsCompares(long hits);
  public static native void   hitsFree(long hits);

  /** ff_score over guides + CSR hit list; out = {cfdMax[n], cfdSpecificity[n], hsu2013[n], perOtCfd[nHits]} (NaN = "not scored") */
  public static native double[][] score(long ctx, long[] guides, long[] rowPtr, long[] targets, int metrics);
}
