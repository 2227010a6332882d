// Scala glue for FlashFry (see INTEGRATION.md section 4).  Not compiled here (no JVM toolchain).
package reference.traverser

/** Drop-in for SeekTraverser / LinearTraverser (trait Traverser, Traverser.scala:52-59).  `traversal` is ignored (and
  * with --gpu OffTargetDiscovery must not even BUILD it: `new OrderedBinTraversalFactory(...)` at
  * OffTargetDiscovery.scala:109-110 tests 4^7 bins x G guides up front -- 1.6e9 JVM compares for 100 000 guides in front
  * of a 3 ms scan; pass LinearTraversal or null).  The contract on return is identical: hits appended in database order,
  * cut by the overflow rule, currentTotal updated.
  *
  * maximumOffTargets is a constructor argument of CRISPRSiteOT without an accessor (crispr/CRISPRSiteOT.scala:31), so it
  * is handed in by OffTargetDiscovery, which parsed it (OffTargetDiscovery.scala:57-59). */
class GpuTraverser(maximumOffTargets: Int, wantPositions: Boolean) extends Traverser with LazyLogging {
  lazy val ctx: Long = flashfry.NativeBridge.create(sys.env.getOrElse("FLASHFRY_GPU", "0").toInt)

  def scan(binaryFile: File, header: BinaryHeader, traversal: BinTraversal, aggregator: ResultsAggregator,
           maxMismatch: Int, configuration: ParameterPack, bitCoder: BitEncoding, posCoder: BitPosition) {
    flashfry.NativeBridge.loadDatabase(ctx, binaryFile.getAbsolutePath)
    val guides  = aggregator.indexedGuides.map(_.guide)                        // ResultsAggregator order
    val hits    = flashfry.NativeBridge.discover(ctx, guides, maxMismatch, maximumOffTargets, wantPositions)
    val rowPtr  = flashfry.NativeBridge.hitsRowPtr(hits);  val targets = flashfry.NativeBridge.hitsTargets(hits)
    val posPtr  = flashfry.NativeBridge.hitsPosPtr(hits);  val pos     = flashfry.NativeBridge.hitsPositions(hits)
    var g = 0
    while (g < guides.length) {
      var i = rowPtr(g).toInt
      while (i < rowPtr(g + 1)) {                                              // same objects the reference would have built
        val coords = if (posPtr != null) java.util.Arrays.copyOfRange(pos, posPtr(i).toInt, posPtr(i + 1).toInt)
                     else new Array[Long](bitCoder.getCount(targets(i)))       // count placeholders, as --positionOutput off prints no coordinates
        aggregator.updateOT(aggregator.indexedGuides(g), new CRISPRHit(targets(i), coords))
        i += 1
      }
      g += 1
    }
    Traverser.allComparisons += flashfry.NativeBridge.hitsCompares(hits)     // the log line at OffTargetDiscovery.scala:137
    flashfry.NativeBridge.hitsFree(hits)
  }

  /** The fast path for `discover` (no per-hit objects): the TSV of TabDelimitedHandler.scala:132-154 written natively.
    * OffTargetDiscovery calls this INSTEAD of scan + TabDelimitedOutput.write when no score needs the objects. */
  def scanToTsv(binaryFile: File, aggregator: ResultsAggregator, maxMismatch: Int, output: File) {
    flashfry.NativeBridge.loadDatabase(ctx, binaryFile.getAbsolutePath)
    val gs   = aggregator.wrappedGuides
    val hits = flashfry.NativeBridge.discover(ctx, aggregator.indexedGuides.map(_.guide), maxMismatch, maximumOffTargets, wantPositions)
    flashfry.NativeBridge.hitsWriteTsv(ctx, hits, output.getAbsolutePath,
      gs.map(_.otSite.target.contig), gs.map(_.otSite.target.start), gs.map(_.otSite.target.bases),
      gs.map(_.otSite.target.sequenceContext.orNull), gs.map(_.otSite.target.forwardStrand), wantPositions)
    flashfry.NativeBridge.hitsFree(hits)
  }
}
