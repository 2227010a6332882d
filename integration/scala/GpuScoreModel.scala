// Scala glue for FlashFry (see INTEGRATION.md section 4).  Not compiled here (no JVM toolchain).
package scoring

/** Doench2016CFDScore / CrisprMitEduOffTarget on the GPU: same names, columns and validity rules, so
  * `ScoreResults.getRegisteredScoringMetric` can hand these out for "doench2016cfd" / "hsu2013". */
class GpuScoreModel(metric: Int, ctx: Long) extends ScoreModel {
  private val cpu: ScoreModel = if (metric == 1) new Doench2016CFDScore() else new CrisprMitEduOffTarget()
  def scoreName() = cpu.scoreName(); def scoreDescription() = cpu.scoreDescription(); def headerColumns() = cpu.headerColumns()
  def validOverEnzyme(e: ParameterPack) = cpu.validOverEnzyme(e)
  def validOverTargetSequence(e: ParameterPack, g: CRISPRSiteOT) = cpu.validOverTargetSequence(e, g)
  def setup() {}; def bitEncoder(b: BitEncoding) {}

  def scoreGuides(guides: Array[CRISPRSiteOT], bitEnc: BitEncoding, posEnc: BitPosition, pack: ParameterPack) {
    if (!validOverEnzyme(pack)) { guides.foreach(g => headerColumns().foreach(c => g.namedAnnotations(c) = Array("NA"))); return }
    val rowPtr  = guides.scanLeft(0L)(_ + _.offTargets.size)
    val targets = guides.flatMap(_.offTargets.map(_.sequence))
    val out = flashfry.NativeBridge.score(ctx, pack.enzyme.index, guides.map(_.longEncoding), rowPtr, targets, metric)
    guides.zipWithIndex.foreach { case (g, i) =>
      if (metric == 1) {
        g.namedAnnotations("DoenchCFD_maxOT") = Array(out(0)(i).toString)                 // already thresholded at 0.023
        g.namedAnnotations("DoenchCFD_specificityscore") = Array(out(1)(i).toString)
        g.offTargets.zipWithIndex.foreach { case (ot, j) =>
          val v = out(3)(rowPtr(i).toInt + j); if (!v.isNaN) ot.addScore("Doench2016CFDScore", v.toString) }
      } else g.namedAnnotations("Hsu2013") = Array(out(2)(i).toString)
    }
  }
}
