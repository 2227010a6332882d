// CPU-only self-test helper of the host mirror: re-reads a discover/score TSV with TabDelimitedInput and writes it back
// with TabDelimitedOutput (the reference's own round-trip test, src/test/scala/targetio/TabDelimitedHanderTest.scala:40-51),
// or lists the guides SimpleSiteFinder extracts from a FASTA.  No GPU needed.
#include <cstdio>

#include "flashfry_host.hpp"

using namespace flashfry;

struct PassThrough : ScoreModel {  // replays already-present annotation columns
  std::vector<std::string> cols;
  std::string scoreName() const override { return "passthrough"; }
  std::vector<std::string> headerColumns() const override { return cols; }
  bool validOverEnzyme(const ParameterPack &) const override { return true; }
  void scoreGuides(std::vector<CRISPRSiteOT> &, const BitEncoding &, const ParameterPack &) override {}
};

int main(int argc, char **argv) {
  try {
    if (argc >= 5 && std::string(argv[1]) == "roundtrip") {
      const ParameterPack &pack = indexToParameterPack(atoi(argv[2]));
      BitEncoding be(pack);
      BitPosition bp;
      TabDelimitedInput in(argv[3], be, bp, INT32_MAX, false);
      PassThrough pt;
      pt.cols = in.annotations;
      const bool positions = argc >= 6 && std::string(argv[5]) == "positions";
      TabDelimitedOutput out(argv[4], be, bp, {&pt}, true, positions);
      for (auto &g : in.guides) out.write(g);
      out.close();
      return 0;
    }
    if (argc >= 5 && std::string(argv[1]) == "sites") {
      const ParameterPack &pack = indexToParameterPack(atoi(argv[2]));
      GuideMemoryStorage g;
      findTargetSites(argv[3], &g, pack, atoi(argv[4]));
      for (auto &s : g.guideHits)
        printf("%s\t%d\t%s\t%s\t%s\n", s.contig.c_str(), s.position, s.bases.c_str(), s.forwardStrand ? "FWD" : "RVS",
               s.sequenceContext.empty() ? "NONE" : s.sequenceContext.c_str());
      return 0;
    }
    if (argc >= 5 && std::string(argv[1]) == "traverse") {  // needs a GPU: the object-building Traverser drop-in
      // (GpuTraverser::scan -> CRISPRHit per hit -> TabDelimitedOutput), what FlashFry's own discover does after a
      // Traverser.scan; the CLI writes the same file through ff_hits_write_tsv without the objects
      const std::string db = argv[2];
      const bool positions = argc >= 6 && std::string(argv[5]) == "positions";
      const int maxMismatch = argc >= 7 ? atoi(argv[6]) : 4, maxOT = argc >= 8 ? atoi(argv[7]) : 2000;
      BinaryHeader header = BinaryHeader::readHeader(db + ".header");
      const ParameterPack &pack = *header.inputParameterPack;
      BitEncoding be(pack);
      GuideMemoryStorage found;
      findTargetSites(argv[3], &found, pack, 6);
      std::vector<CRISPRSiteOT> ots;
      for (auto &g : found.guideHits) ots.push_back({g, be.bitEncodeString(g.bases, 1), maxOT});
      ResultsAggregator agg(std::move(ots));
      NativeContext nc(0);
      GpuTraverser::scan(nc, db, agg, maxMismatch, positions);
      TabDelimitedOutput out(argv[4], be, header.bitPosition, {}, true, positions);
      for (auto &g : agg.wrappedGuides) out.write(g);
      out.close();
      return 0;
    }
    if (argc >= 3 && std::string(argv[1]) == "double") {
      for (int i = 2; i < argc; ++i) printf("%s\n", javaDoubleToString(strtod(argv[i], nullptr)).c_str());
      return 0;
    }
    fprintf(stderr, "usage: host_selftest roundtrip ENZYME_INDEX in.tsv out.tsv [positions] | sites ENZYME_INDEX fasta flank | traverse DB FASTA OUT [positions|nopos] [maxMismatch] [maxOT] | double x...\n");
    return 2;
  } catch (const std::exception &e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
}
