// flashfry_b200_cli -- FlashFry's `discover` and `score` command-line surface on top of libflashfry_b200.so.
//
// Mirrors modules/OffTargetDiscovery.scala:42-153 and modules/ScoreResults.scala:40-154 (same option names, both the
// `-x` and `--x` spellings, same defaults, same TSV grammar), with the database scan and the CFD / Hsu2013 scorers
// running on the GPU through the C ABI.  Out of scope here: `index`, `random`, `extract` and the
// on-target / annotation metrics -- asking for one of those metrics is an error, not a silent skip.
#include <chrono>
#include <cstdio>
#include <iostream>
#include <map>
#include <string>

#include "flashfry_host.hpp"

using namespace flashfry;

namespace {

struct Args {
  std::map<std::string, std::string> kv;
  std::map<std::string, bool> flags;
  static std::string norm(const std::string &s) {
    size_t i = 0;
    while (i < s.size() && s[i] == '-') ++i;
    return s.substr(i);
  }
  bool has(const std::string &k) const { return kv.count(k) > 0; }
  std::string get(const std::string &k, const std::string &dflt = "") const {
    auto it = kv.find(k);
    return it == kv.end() ? dflt : it->second;
  }
  bool flag(const std::string &k) const { return flags.count(k) > 0; }
};

Args parse(int argc, char **argv, int first, const std::vector<std::string> &flagNames) {
  Args a;
  for (int i = first; i < argc; ++i) {
    std::string k = Args::norm(argv[i]);
    if (argv[i][0] != '-') throw std::invalid_argument(std::string("unexpected argument: ") + argv[i]);
    if (std::find(flagNames.begin(), flagNames.end(), k) != flagNames.end()) { a.flags[k] = true; continue; }
    if (i + 1 >= argc) throw std::invalid_argument("missing value for option " + std::string(argv[i]));
    a.kv[k] = argv[++i];
  }
  return a;
}

void require(const Args &a, const char *k) {
  if (!a.has(k)) throw std::invalid_argument(std::string("Missing required option '--") + k + "'");
}

// modules/OffTargetDiscovery.scala:79-153
int runDiscover(int argc, char **argv) {
  Args a = parse(argc, argv, 2, {"positionOutput", "forceLinear"});
  require(a, "fasta"); require(a, "database"); require(a, "output");
  const int maxMismatch = atoi(a.get("maxMismatch", "4").c_str());
  const int flankingSequence = atoi(a.get("flankingSequence", "6").c_str());
  const int maximumOffTargets = atoi(a.get("maximumOffTargets", "2000").c_str());
  const double minGC = atof(a.get("minGC", "0.0").c_str()), maxGC = atof(a.get("maxGC", "1.0").c_str());
  if (!(minGC >= 0 && minGC <= 1.0) || !(maxGC >= 0 && maxGC <= 1.0)) throw IllegalStateException("assertion failed");  // :81-82
  const bool positions = a.flag("positionOutput");
  const int device = atoi(a.get("device", "0").c_str());
  int bulgeFlags = 0;  // extension, not a FlashFry option: --bulge rna,dna
  for (auto &tok : split(a.get("bulge", ""), ',')) {
    if (tok == "rna") bulgeFlags |= FF_BULGE_RNA;
    else if (tok == "dna") bulgeFlags |= FF_BULGE_DNA;
    else if (!tok.empty()) throw std::invalid_argument("--bulge takes rna, dna or rna,dna");
  }

  fprintf(stderr, "Reading the header....\n");
  BinaryHeader header = BinaryHeader::readHeader(a.get("database") + ".header");
  const ParameterPack &pack = *header.inputParameterPack;
  BitEncoding bitCoder(pack);

  GuideMemoryStorage guideHits;
  findTargetSites(a.get("fasta"), &guideHits, pack, flankingSequence);
  fprintf(stderr, "Setting up the guide recording for our %zu candidate guides....\n", guideHits.guideHits.size());
  GuideMemoryStorage filtered = GuideMemoryStorage::filter_by_GC(guideHits, minGC, maxGC);
  fprintf(stderr, "Filtered GC guide count %zu\n", filtered.guideHits.size());
  std::vector<CRISPRSiteOT> guideOTs;
  for (auto &g : filtered.guideHits) guideOTs.push_back({g, bitCoder.bitEncodeString(g.bases, 1), maximumOffTargets});
  ResultsAggregator guideStorage(std::move(guideOTs));

  fprintf(stderr, "scanning against the known targets from the genome with %zu guides\n", guideStorage.wrappedGuides.size());
  // The TSV fast path: the CSR hit list goes straight to FlashFry's discover grammar through
  // ff_hits_write_tsv -- no CRISPRHit object per hit (GpuTraverser::scan in flashfry_host.hpp is the object-building
  // variant a Traverser drop-in needs; host_selftest exercises it).
  NativeContext nc(device);
  const auto t0 = std::chrono::steady_clock::now();
  ffCheck(ff_load_database(nc.ctx, a.get("database").c_str(), (a.get("database") + ".header").c_str()));
  std::vector<uint64_t> guideLongs;
  std::vector<ff_tsv_guide> rows;
  for (auto &g : guideStorage.wrappedGuides) {
    guideLongs.push_back(g.longEncoding);
    rows.push_back({g.target.contig.c_str(), g.target.start(), g.target.bases.c_str(),
                    g.target.sequenceContext.empty() ? nullptr : g.target.sequenceContext.c_str(), g.target.forwardStrand ? 1 : 0});
  }
  ff_hits *h = nullptr;
  if (bulgeFlags) ffCheck(ff_discover_bulge(nc.ctx, guideLongs.data(), (int64_t)guideLongs.size(), maxMismatch, maximumOffTargets, bulgeFlags, positions ? 1 : 0, &h));
  else ffCheck(ff_discover(nc.ctx, guideLongs.data(), (int64_t)guideLongs.size(), maxMismatch, maximumOffTargets, positions ? 1 : 0, &h));
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  fprintf(stderr, "Performed a total of %llu guide to target comparisons (load + scan %.3f s)\n", (unsigned long long)h->n_compares, secs);
  fprintf(stderr, "Writing final output for %zu guides\n", guideStorage.wrappedGuides.size());
  const int rc = ff_hits_write_tsv(nc.ctx, a.get("output").c_str(), rows.data(), h, positions ? 1 : 0);
  ff_hits_free(h);
  ffCheck(rc);
  return 0;
}

// modules/ScoreResults.scala:90-226
int runScore(int argc, char **argv) {
  Args a = parse(argc, argv, 2, {"includeOTs", "numericOutput", "countOnTargetInScore"});
  require(a, "input"); require(a, "output"); require(a, "scoringMetrics"); require(a, "database");
  const int maxMismatch = a.has("maxMismatch") ? atoi(a.get("maxMismatch").c_str()) : INT32_MAX;
  const bool includeOTs = a.flag("includeOTs");
  const int device = atoi(a.get("device", "0").c_str());

  BinaryHeader header = BinaryHeader::readHeader(a.get("database") + ".header");
  const ParameterPack &pack = *header.inputParameterPack;
  BitEncoding bitEnc(pack);
  fprintf(stderr, "Loading CRISPR objects (filtering out overflow guides).. \n");
  TabDelimitedInput input(a.get("input"), bitEnc, header.bitPosition, maxMismatch, true);

  NativeContext nc(device);
  std::vector<std::unique_ptr<ScoreModel>> models;
  for (auto &nameRaw : split(a.get("scoringMetrics"), ',')) {
    std::string name = nameRaw;
    for (char &c : name) c = (char)tolower((unsigned char)c);
    std::unique_ptr<ScoreModel> m;
    if (name == "hsu2013") m.reset(new GpuScoreModel(nc, FF_METRIC_HSU2013));
    else if (name == "doench2016cfd") m.reset(new GpuScoreModel(nc, FF_METRIC_CFD));
    else if (name == "minot") { auto *ch = new ClosestHit(); ch->nc = &nc; m.reset(ch); }
    else if (name == "dangerous") { auto *d = new DangerousSequences(); d->nc = &nc; d->cleanOutput = a.flag("numericOutput"); m.reset(d); }
    else if (name == "doench2014ontarget" || name == "moreno2015" || name == "bedannotator" || name == "reciprocalofftargets" ||
             name == "rank" || name == "jostandsantos" || name == "folding")
      throw std::invalid_argument("scoring metric '" + nameRaw + "' is outside the GPU hot path of this build (SURVEY.md section 8); use FlashFry itself for it");
    else throw std::invalid_argument("Unknown scoring metric: " + nameRaw);  // ScoreResults.scala:221-223
    if (m->validOverEnzyme(pack)) {
      fprintf(stderr, "adding score: %s\n", m->scoreName().c_str());
      models.push_back(std::move(m));
    } else {
      fprintf(stderr, "DROPPING SCORING METHOD: %s; it's not valid over enzyme parameter pack: %s\n", m->scoreName().c_str(), pack.name.c_str());
    }
  }
  fprintf(stderr, "Scoring all guides...\n");
  std::vector<ScoreModel *> raw;
  for (auto &m : models) { m->scoreGuides(input.guides, bitEnc, pack); raw.push_back(m.get()); }
  ResultsAggregator results(std::move(input.guides));
  TabDelimitedOutput output(a.get("output"), bitEnc, header.bitPosition, raw, includeOTs, true);
  for (auto &g : results.wrappedGuides) output.write(g);
  output.close();
  return 0;
}

void usage() {
  fprintf(stderr,
          "flashfry_b200_cli <discover|score> [options]\n"
          "  discover --fasta FILE --database FILE --output FILE [--positionOutput] [--forceLinear] [--maxMismatch 4]\n"
          "           [--flankingSequence 6] [--maximumOffTargets 2000] [--minGC 0.0] [--maxGC 1.0] [--device 0]\n"
          "           [--bulge rna,dna]   (extension, not in FlashFry: one 1-bp bulge on top of the mismatches; tokens gain _R<q> / _D<q>)\n"
          "  score    --input FILE --output FILE --scoringMetrics hsu2013,doench2016cfd[,minot,dangerous] --database FILE\n"
          "           [--maxMismatch N] [--includeOTs] [--numericOutput] [--device 0]\n");
}

}  // namespace

int main(int argc, char **argv) {
  if (argc < 2) { usage(); return 2; }
  try {
    const std::string cmd = argv[1];
    if (cmd == "discover") return runDiscover(argc, argv);
    if (cmd == "score") return runScore(argc, argv);
    usage();
    fprintf(stderr, "unknown command: %s (index / random / extract are outside the GPU hot path of this build)\n", cmd.c_str());
    return 2;
  } catch (const NativeError &e) {
    fprintf(stderr, "error from libflashfry_b200 (%d): %s\n", e.code, e.what());
    return 1;
  } catch (const std::exception &e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
}
