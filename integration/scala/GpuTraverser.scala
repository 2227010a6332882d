// Scala glue for FlashFry (see INTEGRATION.md section 4).  Not compiled here (no JVM toolchain).
package reference.traverser

/** Drop-in for SeekTraverser / LinearTraverser (trait Traverser, Traverser.scala:52-59).  `traversal` is ignored:
  * the native side does its own pruning; the contract on return is identical (hits appended in database order, cut by
  * the overflow rule, currentTotal updated). */
object GpuTraverser extends Traverser with LazyLogging {
  lazy val ctx: Long = flashfry.NativeBridge.create(sys.env.getOrElse("FLASHFRY_GPU", "0").toInt)

  def scan(binaryFile: File, header: BinaryHeader, traversal: BinTraversal, aggregator: ResultsAggregator,
           maxMismatch: Int, configuration: ParameterPack, bitCoder: BitEncoding, posCoder: BitPosition) {
    flashfry.NativeBridge.loadDatabase(ctx, binaryFile.getAbsolutePath)
    val guides  = aggregator.indexedGuides.map(_.guide)                        // ResultsAggregator order
    val maxOT   = aggregator.wrappedGuides.headOption.map(_.otSite.overflowValue).getOrElse(2000)
    val hits    = flashfry.NativeBridge.discover(ctx, guides, maxMismatch, maxOT, true)
    val rowPtr  = flashfry.NativeBridge.hitsRowPtr(hits);  val targets = flashfry.NativeBridge.hitsTargets(hits)
    val posPtr  = flashfry.NativeBridge.hitsPosPtr(hits);  val pos     = flashfry.NativeBridge.hitsPositions(hits)
    var g = 0
    while (g < guides.length) {
      var i = rowPtr(g).toInt
      while (i < rowPtr(g + 1)) {                                              // same objects the reference would have built
        aggregator.updateOT(aggregator.indexedGuides(g),
          new CRISPRHit(targets(i), java.util.Arrays.copyOfRange(pos, posPtr(i).toInt, posPtr(i + 1).toInt)))
        i += 1
      }
      g += 1
    }
    Traverser.allComparisons += flashfry.NativeBridge.hitsCompares(hits)     // the log line at OffTargetDiscovery.scala:137
    flashfry.NativeBridge.hitsFree(hits)
  }
}
