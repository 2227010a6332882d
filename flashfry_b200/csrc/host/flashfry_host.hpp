// Host-side mirror of the FlashFry types that sit on either side of the GPU hot path.
//
// The reference is Scala/JVM and there is no JVM toolchain in this image, so the host side above the C ABI is
// written in C++ with the reference's own names, argument meaning and error behaviour (an exception with the
// reference's message where the reference asserts/throws).  Citations are file:line in the FlashFry checkout,
// src/main/scala/...
#pragma once

#include <zlib.h>

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "flashfry_b200.h"

namespace flashfry {

struct IllegalStateException : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// ---------------------------------------------------------------------------------------------------------------
// standards/StandardScanParameters.scala:28-215
struct ParameterPack {
  std::string name;
  int index;
  int totalScanLength;
  int pamLength;
  bool fivePrimePam;
  uint64_t comparisonBitEncoding;
  bool cas9Type;
  // fwdRegex / revRegex restated as one allowed-letter set per position (the regexes are a consumed character plus a
  // fixed-length look-ahead, so every offset is tested independently)
  std::vector<std::string> fwd, rev;
  std::pair<int, int> guideRange() const {  // :107,:213
    return fivePrimePam ? std::make_pair(pamLength, totalScanLength) : std::make_pair(0, totalScanLength - pamLength);
  }
};

inline std::vector<std::string> pattern(std::initializer_list<std::pair<const char *, int>> parts) {
  std::vector<std::string> out;
  for (auto &p : parts)
    for (int i = 0; i < p.second; ++i) out.emplace_back(p.first);
  return out;
}

inline const std::vector<ParameterPack> &allPacks() {
  static const char *N = "ACGT";
  static const std::vector<ParameterPack> packs = {
      {"CPF1", 1, 24, 4, true, 0x00FFFFFFFFFFull, false, pattern({{"T", 3}, {N, 21}}), pattern({{N, 21}, {"A", 3}})},                        // :199-215
      {"SPCAS9", 2, 23, 3, false, 0x3FFFFFFFFFC0ull, true, pattern({{N, 21}, {"AG", 1}, {"G", 1}}), pattern({{"C", 1}, {"CT", 1}, {N, 21}})},    // :90-109
      {"SPCAS9NGG", 3, 23, 3, false, 0x3FFFFFFFFFC0ull, true, pattern({{N, 21}, {"G", 2}}), pattern({{"C", 2}, {N, 21}})},                      // :134-153
      {"SPCAS9NAG", 4, 23, 3, false, 0x3FFFFFFFFFC0ull, true, pattern({{N, 21}, {"A", 1}, {"G", 1}}), pattern({{"C", 1}, {"T", 1}, {N, 21}})},   // :178-197
      {"SPCAS919", 5, 22, 3, false, 0x0FFFFFFFFFC0ull, true, pattern({{N, 20}, {"AG", 1}, {"G", 1}}), pattern({{"C", 1}, {"CT", 1}, {N, 20}})},  // :112-131
      {"SPCAS9NGG19", 6, 22, 3, false, 0x0FFFFFFFFFC0ull, true, pattern({{N, 20}, {"G", 2}}), pattern({{"C", 2}, {N, 20}})},                    // :156-175
  };
  return packs;
}

inline const ParameterPack &indexToParameterPack(int index) {  // :61-70
  for (auto &p : allPacks())
    if (p.index == index) return p;
  throw IllegalStateException("Unable to find the correct parameter pack for enzyme: " + std::to_string(index));
}

// ---------------------------------------------------------------------------------------------------------------
// bitcoding/BitEncoding.scala
struct BitEncoding {
  const ParameterPack &mParameterPack;
  explicit BitEncoding(const ParameterPack &p) : mParameterPack(p) {}
  static constexpr uint64_t stringMask = 0xFFFFFFFFFFFFull;

  uint64_t bitEncodeString(const std::string &str, int count = 1) const {  // :46-67
    if (str.size() > 24) throw std::invalid_argument("String " + str + " is too long to be encoded (" + std::to_string(str.size()) + " > 24)");
    if (count < 1) throw std::invalid_argument("String count " + str + " - " + std::to_string(count) + " has a count <= 0");
    uint64_t enc = 0;
    for (char ch : str) {
      enc <<= 2;
      switch (ch) {
        case 'A': case 'a': break;
        case 'C': case 'c': enc |= 1; break;
        case 'G': case 'g': enc |= 2; break;
        case 'T': case 't': enc |= 3; break;
        default: throw IllegalStateException(std::string("Unable to encode character ") + ch);
      }
    }
    return enc | ((uint64_t)count << 48);
  }
  std::string bitDecodeString(uint64_t enc, int *count = nullptr) const {  // :85-99
    const int n = mParameterPack.totalScanLength;
    std::string s(n, 'A');
    for (int i = 0; i < n; ++i) s[n - 1 - i] = "ACGT"[(enc >> (2 * i)) & 3];
    if (count) *count = (int)(int16_t)(enc >> 48);
    return s;
  }
  int getCount(uint64_t enc) const { return (int)(int16_t)(enc >> 48); }  // :114
  int mismatches(uint64_t a, uint64_t b, uint64_t additionalMask = stringMask) const {  // :127-132
    const uint64_t first = (a ^ b) & additionalMask & mParameterPack.comparisonBitEncoding;
    return __builtin_popcountll((first & 0xAAAAAAAAAAAAull) | ((first << 1) & 0xAAAAAAAAAAAAull));
  }
};

// bitcoding/BitPosition.scala:32-92
struct BitPosition {
  std::vector<std::string> indexToContig;  // 1-based ids
  void addReference(const std::string &name) { indexToContig.push_back(name); }
  uint64_t encode(const std::string &ref, int position, int targetLength, bool forwardStrand) const {
    auto it = std::find(indexToContig.begin(), indexToContig.end(), ref);
    if (it == indexToContig.end()) throw IllegalStateException("Unknown contig: " + ref);
    const uint64_t id = (uint64_t)(it - indexToContig.begin()) + 1;
    return (id << 32) | (uint64_t)(uint32_t)position | (forwardStrand ? 0ull : (1ull << 60)) | ((uint64_t)targetLength << 52);
  }
  struct Decoded { std::string contig; int start; int size; bool forwardStrand; };
  Decoded decode(uint64_t e) const {
    const size_t id = (size_t)((e >> 32) & 0xFFFFF);
    if (id == 0 || id > indexToContig.size()) throw IllegalStateException("position refers to unknown contig id " + std::to_string(id));
    return {indexToContig[id - 1], (int)(e & 0xFFFFFFFFull), (int)((e >> 52) & 0xFF), ((e >> 60) & 0xF) == 0};
  }
};

// ---------------------------------------------------------------------------------------------------------------
// crispr/CRISPRSite.scala, CRISPRHit.scala, CRISPRSiteOT.scala
struct CRISPRSite {
  std::string contig, bases;
  bool forwardStrand;
  int position;
  std::string sequenceContext;  // empty == None
  int start() const { return position; }
  int length() const { return (int)bases.size(); }
};

struct CRISPRHit {
  uint64_t sequence;
  std::vector<uint64_t> coordinates;
  bool validOffTargetCoordinates = true;
  std::vector<std::pair<std::string, std::string>> scores;  // addScore :103-109
  // bulge extension only (not in the reference): mismatches of the best alignment and its code (0 none, 0x40|q RNA
  // bulge at guide base q, 0x80|q DNA bulge at genomic base q); -1 = plain mismatch search
  int bulgeMismatches = -1;
  uint8_t bulge = 0;
  int getOffTargetCount() const { return (int)coordinates.size(); }
};

struct CRISPRSiteOT {
  CRISPRSite target;
  uint64_t longEncoding;
  int overflow;
  bool inheritedOverflow = false;
  std::vector<CRISPRHit> offTargets;
  int currentTotal = 0;
  std::map<std::string, std::vector<std::string>> namedAnnotations;
  bool full() const { return currentTotal >= overflow; }  // :39
  void addOT(CRISPRHit hit) {                              // :41-46
    currentTotal += hit.getOffTargetCount();
    offTargets.push_back(std::move(hit));
  }
};

// crispr/ResultsAggregator.scala:32-49 -- guides sorted by target.start (ties keep their discovery order here; the
// reference's quickSort leaves them unspecified)
struct ResultsAggregator {
  std::vector<CRISPRSiteOT> wrappedGuides;
  explicit ResultsAggregator(std::vector<CRISPRSiteOT> guides) : wrappedGuides(std::move(guides)) {
    std::stable_sort(wrappedGuides.begin(), wrappedGuides.end(),
                     [](const CRISPRSiteOT &a, const CRISPRSiteOT &b) { return a.target.start() < b.target.start(); });
  }
};

// ---------------------------------------------------------------------------------------------------------------
// utils/Utils.scala
inline double gcContent(const std::string &g) {  // :46
  int n = 0;
  for (char c : g) n += (c == 'C' || c == 'G' || c == 'c' || c == 'g');
  return (double)n / (double)g.size();
}
inline std::string reverseCompString(const std::string &s) {  // :88
  std::string out(s.rbegin(), s.rend());
  for (char &c : out) switch (c) {
      case 'A': c = 'T'; break; case 'C': c = 'G'; break; case 'G': c = 'C'; break; case 'T': c = 'A'; break;
      case 'a': c = 't'; break; case 'c': c = 'g'; break; case 'g': c = 'c'; break; case 't': c = 'a'; break;
      default: break;
    }
  return out;
}

// Java's Double.toString (shortest digits that round-trip, fixed notation in [1e-3, 1e7), otherwise d.dddE[-]n)
inline std::string javaDoubleToString(double x) {
  if (std::isnan(x)) return "NaN";
  if (std::isinf(x)) return x > 0 ? "Infinity" : "-Infinity";
  if (x == 0) return std::signbit(x) ? "-0.0" : "0.0";
  char buf[64];
  auto r = std::to_chars(buf, buf + sizeof buf, std::fabs(x), std::chars_format::scientific);
  std::string sci(buf, r.ptr);  // d[.ddd]e[+-]XX
  const size_t epos = sci.find('e');
  std::string digits;
  for (size_t i = 0; i < epos; ++i)
    if (sci[i] != '.') digits.push_back(sci[i]);
  const int e10 = atoi(sci.c_str() + epos + 1);
  while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
  std::string out = std::signbit(x) ? "-" : "";
  const double ax = std::fabs(x);
  if (ax >= 1e-3 && ax < 1e7) {
    if (e10 >= 0) {
      std::string whole = digits.substr(0, std::min(digits.size(), (size_t)e10 + 1));
      whole.append((size_t)e10 + 1 - whole.size(), '0');
      std::string frac = digits.size() > (size_t)e10 + 1 ? digits.substr(e10 + 1) : "0";
      out += whole + "." + frac;
    } else {
      out += "0." + std::string((size_t)(-e10 - 1), '0') + digits;
    }
  } else {
    out += digits.substr(0, 1) + "." + (digits.size() > 1 ? digits.substr(1) : "0") + "E" + std::to_string(e10);
  }
  return out;
}

// ---------------------------------------------------------------------------------------------------------------
// reference/ReferenceEncoder.scala:46-169
inline std::vector<std::string> readLines(const std::string &path) {
  std::vector<std::string> lines;
  gzFile f = gzopen(path.c_str(), "rb");  // transparently reads plain text too (ReferenceEncoder.fileToSource :76-82)
  if (!f) throw std::runtime_error("cannot open " + path);
  std::string cur;
  char buf[1 << 16];
  int n;
  while ((n = gzread(f, buf, sizeof buf)) > 0) {
    for (int i = 0; i < n; ++i) {
      if (buf[i] == '\n') { if (!cur.empty() && cur.back() == '\r') cur.pop_back(); lines.push_back(cur); cur.clear(); }
      else cur.push_back(buf[i]);
    }
  }
  gzclose(f);
  if (!cur.empty()) lines.push_back(cur);
  return lines;
}

struct GuideMemoryStorage {  // crispr/GuideMemoryStorage.scala
  std::vector<CRISPRSite> guideHits;
  void addHit(CRISPRSite s) { guideHits.push_back(std::move(s)); }
  static GuideMemoryStorage filter_by_GC(const GuideMemoryStorage &in, double lowGC, double highGC) {  // :41-50
    GuideMemoryStorage out;
    for (auto &g : in.guideHits) {
      const double gc = gcContent(g.bases);
      if (gc >= lowGC && gc <= highGC) out.addHit(g);
    }
    return out;
  }
};

inline bool matchesAt(const std::string &s, size_t i, const std::vector<std::string> &pat) {
  if (i + pat.size() > s.size()) return false;
  for (size_t j = 0; j < pat.size(); ++j)
    if (pat[j].find(s[i + j]) == std::string::npos) return false;
  return true;
}

// SimpleSiteFinder.reset :114-169: all forward matches of a contig, then all reverse matches
inline void findSitesInContig(const std::string &contig, const std::string &seq, const ParameterPack &params, int flank,
                              GuideMemoryStorage *out) {
  const int L = params.totalScanLength;
  for (int pass = 0; pass < 2; ++pass) {
    const auto &pat = pass == 0 ? params.fwd : params.rev;
    for (size_t start = 0; start + (size_t)L <= seq.size(); ++start) {
      if (!matchesAt(seq, start, pat)) continue;
      const size_t end = start + L;
      const size_t c0 = start >= (size_t)flank ? start - flank : 0, c1 = std::min(seq.size(), end + flank);
      std::string sub = seq.substr(start, L), ctx = seq.substr(c0, c1 - c0);
      if (pass == 1) { sub = reverseCompString(sub); ctx = reverseCompString(ctx); }
      out->addHit({contig, sub, pass == 0, (int)start, (int)ctx.size() == L + 2 * flank ? ctx : std::string()});
    }
  }
}

// ReferenceEncoder.findTargetSites :46-70
inline void findTargetSites(const std::string &fasta, GuideMemoryStorage *out, const ParameterPack &params, int flank,
                            BitPosition *posEncoder = nullptr) {
  std::string name, buf;
  bool have = false;
  for (auto &line : readLines(fasta)) {
    if (!line.empty() && line[0] == '>') {
      if (have) findSitesInContig(name, buf, params, flank, out);
      name = line.substr(1);
      for (char &c : name) if (c == ' ' || c == '\t') c = '_';
      if (posEncoder) posEncoder->addReference(name);
      buf.clear();
      have = true;
    } else {
      for (char c : line) buf.push_back((char)toupper((unsigned char)c));
    }
  }
  if (have) findSitesInContig(name, buf, params, flank, out);
}

// ---------------------------------------------------------------------------------------------------------------
// reference/binary/BinaryHeader.scala:115-160 -- only what the host needs (the library parses the bin table itself)
struct BinaryHeader {
  const ParameterPack *inputParameterPack = nullptr;
  int binWidth = 0;
  BitPosition bitPosition;
  static BinaryHeader readHeader(const std::string &filename) {
    std::ifstream in(filename);
    if (!in) throw std::runtime_error("cannot open header " + filename);
    std::string line;
    BinaryHeader h;
    if (!std::getline(in, line) || strtoull(line.c_str(), nullptr, 10) != 0x1234ABCDE123890ull)
      throw IllegalStateException("Binary file " + filename + " doesn't have the magic number expected at the top of the file");
    if (!std::getline(in, line) || atoll(line.c_str()) != 1)
      throw IllegalStateException("Binary file " + filename + " doesn't have the correct version, expecting 1");
    std::getline(in, line);
    h.inputParameterPack = &indexToParameterPack(atoi(line.c_str()));
    std::getline(in, line);
    const long long binCount = atoll(line.c_str());
    h.binWidth = (int)(std::log((double)binCount) / std::log(4.0));
    for (long long b = 0; b < binCount; ++b)
      if (!std::getline(in, line)) throw IllegalStateException("Missing line for bin " + std::to_string(b));
    while (std::getline(in, line))
      if (!line.empty()) h.bitPosition.addReference(line.substr(0, line.find('=')));
    return h;
  }
};

// ---------------------------------------------------------------------------------------------------------------
// the native library behind the Traverser / ScoreModel seams
struct NativeError : std::runtime_error {
  int code;
  NativeError(int c, const char *msg) : std::runtime_error(msg), code(c) {}
};
inline void ffCheck(int rc) {
  if (rc != FF_OK) throw NativeError(rc, ff_last_error());
}

struct NativeContext {
  ff_ctx *ctx = nullptr;
  explicit NativeContext(int device = 0) { ffCheck(ff_create(&ctx, device)); }
  ~NativeContext() { ff_destroy(ctx); }
  NativeContext(const NativeContext &) = delete;
};

// Replaces SeekTraverser.scan / LinearTraverser.scan (reference/traverser/Traverser.scala:52-59).  The traversal
// argument of the reference is not needed (the native side does its own pruning); the contract is the same: on return
// every aggregator.wrappedGuides(i) holds its CRISPRHit list in database order, cut by the overflow rule.
struct GpuTraverser {
  // bulgeFlags != 0 selects the 1-bp bulge extension (ff_discover_bulge); 0 is the reference's search
  static uint64_t scan(NativeContext &nc, const std::string &binaryFile, ResultsAggregator &aggregator, int maxMismatch,
                       bool wantPositions, int bulgeFlags = 0) {
    ffCheck(ff_load_database(nc.ctx, binaryFile.c_str(), (binaryFile + ".header").c_str()));
    std::vector<uint64_t> guides;
    int maxOT = 2000;
    for (auto &g : aggregator.wrappedGuides) { guides.push_back(g.longEncoding); maxOT = g.overflow; }
    ff_hits *h = nullptr;
    if (bulgeFlags) ffCheck(ff_discover_bulge(nc.ctx, guides.data(), (int64_t)guides.size(), maxMismatch, maxOT, bulgeFlags, wantPositions ? 1 : 0, &h));
    else ffCheck(ff_discover(nc.ctx, guides.data(), (int64_t)guides.size(), maxMismatch, maxOT, wantPositions ? 1 : 0, &h));
    for (int64_t g = 0; g < h->n_guides; ++g) {
      CRISPRSiteOT &ot = aggregator.wrappedGuides[g];
      for (int64_t i = h->row_ptr[g]; i < h->row_ptr[g + 1]; ++i) {
        CRISPRHit hit;
        hit.sequence = h->targets[i];
        const int count = (int)(int16_t)(h->targets[i] >> 48);
        if (h->pos_ptr) hit.coordinates.assign(h->positions + h->pos_ptr[i], h->positions + h->pos_ptr[i + 1]);
        else { hit.coordinates.assign((size_t)count, 0); hit.validOffTargetCoordinates = false; }
        if (bulgeFlags) { hit.bulgeMismatches = h->mismatches[i]; hit.bulge = h->bulge ? h->bulge[i] : 0; }
        ot.addOT(std::move(hit));
      }
      if (ot.currentTotal != h->total_count[g] || ot.full() != (h->overflowed[g] != 0))
        throw IllegalStateException("native overflow bookkeeping disagrees with CRISPRSiteOT");
    }
    const uint64_t compares = h->n_compares;
    ff_hits_free(h);
    return compares;
  }
};

// ---------------------------------------------------------------------------------------------------------------
// scoring/ScoreModel.scala:31-133
struct ScoreModel {
  virtual ~ScoreModel() = default;
  virtual std::string scoreName() const = 0;
  virtual std::vector<std::string> headerColumns() const = 0;
  virtual bool validOverEnzyme(const ParameterPack &p) const = 0;
  virtual void scoreGuides(std::vector<CRISPRSiteOT> &guides, const BitEncoding &bitEnc, const ParameterPack &pack) = 0;
  static constexpr const char *missingAnnotation = "NA";  // :136
};

// Doench2016CFDScore (scoring/Doench2016CFDScore.scala) and CrisprMitEduOffTarget (scoring/CrisprMitEduOffTarget.scala)
// served by one ff_score call each; per-off-target CFD values are attached like ot.addScore(...) at :72.
struct GpuScoreModel : ScoreModel {
  NativeContext &nc;
  uint32_t metric;
  GpuScoreModel(NativeContext &c, uint32_t m) : nc(c), metric(m) {}
  std::string scoreName() const override { return metric == FF_METRIC_CFD ? "Doench2016CFDScore" : "Hsu2013"; }
  std::vector<std::string> headerColumns() const override {
    if (metric == FF_METRIC_CFD) return {"DoenchCFD_maxOT", "DoenchCFD_specificityscore"};
    return {"Hsu2013"};
  }
  bool validOverEnzyme(const ParameterPack &p) const override { return p.cas9Type && p.totalScanLength == 23; }
  void scoreGuides(std::vector<CRISPRSiteOT> &guides, const BitEncoding &, const ParameterPack &pack) override {
    if (!validOverEnzyme(pack)) {  // SingleGuideScoreModel.scoreGuides :125-128
      for (auto &g : guides)
        for (auto &c : headerColumns()) g.namedAnnotations[c] = {missingAnnotation};
      return;
    }
    std::vector<uint64_t> enc, targets;
    std::vector<int64_t> rowPtr{0};
    for (auto &g : guides) {
      enc.push_back(g.longEncoding);
      for (auto &ot : g.offTargets) targets.push_back(ot.sequence);
      rowPtr.push_back((int64_t)targets.size());
    }
    ff_hits h{};
    h.n_guides = (int64_t)guides.size(); h.n_hits = (int64_t)targets.size();
    h.row_ptr = rowPtr.data(); h.targets = targets.data();
    std::vector<double> a(guides.size() + 1), b(guides.size() + 1), c(guides.size() + 1), per(targets.size() + 1);
    ffCheck(ff_score(nc.ctx, enc.data(), &h, metric, a.data(), b.data(), c.data(), per.data()));
    for (size_t g = 0; g < guides.size(); ++g) {
      if (metric == FF_METRIC_CFD) {
        guides[g].namedAnnotations["DoenchCFD_maxOT"] = {javaDoubleToString(a[g])};
        guides[g].namedAnnotations["DoenchCFD_specificityscore"] = {javaDoubleToString(b[g])};
        for (int64_t i = rowPtr[g]; i < rowPtr[g + 1]; ++i)
          if (!std::isnan(per[i])) guides[g].offTargets[i - rowPtr[g]].scores.emplace_back("Doench2016CFDScore", javaDoubleToString(per[i]));
      } else {
        guides[g].namedAnnotations["Hsu2013"] = {javaDoubleToString(c[g])};
      }
    }
  }
};

// ff_hit_aggregates over all guides at once: closest / count / 0..4 histogram / in-genome occurrences
struct HitAggregates {
  std::vector<int32_t> closest, count, hist, inGenome;
  HitAggregates(NativeContext &nc, const ParameterPack &pack, const std::vector<CRISPRSiteOT> &guides) {
    std::vector<uint64_t> enc, targets;
    std::vector<int64_t> rowPtr{0};
    for (auto &g : guides) {
      enc.push_back(g.longEncoding);
      for (auto &ot : g.offTargets) targets.push_back(ot.sequence);
      rowPtr.push_back((int64_t)targets.size());
    }
    ff_hits h{};
    h.n_guides = (int64_t)guides.size(); h.n_hits = (int64_t)targets.size();
    h.row_ptr = rowPtr.data(); h.targets = targets.data();
    const size_t n = guides.size() + 1;
    closest.assign(n, 0); count.assign(n, 0); hist.assign(n * 5, 0); inGenome.assign(n, 0);
    ffCheck(ff_hit_aggregates(nc.ctx, pack.index, enc.data(), &h, closest.data(), count.data(), hist.data(), inGenome.data()));
  }
};

// scoring/ClosestHit.scala:43-76 ("minot") -- integer aggregate over the same hit list, on the GPU (ff_hit_aggregates)
struct ClosestHit : ScoreModel {
  NativeContext *nc = nullptr;
  std::string scoreName() const override { return "closest"; }
  std::vector<std::string> headerColumns() const override { return {"basesDiffToClosestHit", "closestHitCount", "0-1-2-3-4_mismatch"}; }
  bool validOverEnzyme(const ParameterPack &) const override { return true; }
  void scoreGuides(std::vector<CRISPRSiteOT> &guides, const BitEncoding &, const ParameterPack &pack) override {
    if (!nc) throw IllegalStateException("ClosestHit needs a native context: the reduction runs on the GPU, there is no host path");
    HitAggregates agg(*nc, pack, guides);
    for (size_t i = 0; i < guides.size(); ++i) {
      std::string hs;
      for (int m = 0; m < 5; ++m) hs += (m ? "," : "") + std::to_string(agg.hist[i * 5 + m]);
      const bool none = agg.closest[i] == INT32_MAX;
      guides[i].namedAnnotations["basesDiffToClosestHit"] = {none ? "UNK" : std::to_string(agg.closest[i])};
      guides[i].namedAnnotations["closestHitCount"] = {none ? "0" : std::to_string(agg.count[i])};
      guides[i].namedAnnotations["0-1-2-3-4_mismatch"] = {hs};
    }
  }
};

// scoring/DangerousSequences.scala:49-68
struct DangerousSequences : ScoreModel {
  NativeContext *nc = nullptr;
  bool cleanOutput = false;
  std::string scoreName() const override { return "dangerous"; }
  std::vector<std::string> headerColumns() const override { return {"dangerous_GC", "dangerous_polyT", "dangerous_in_genome"}; }
  bool validOverEnzyme(const ParameterPack &) const override { return true; }
  void scoreGuides(std::vector<CRISPRSiteOT> &guides, const BitEncoding &, const ParameterPack &pack) override {
    if (!nc) throw IllegalStateException("DangerousSequences needs a native context: the in-genome count runs on the GPU");
    HitAggregates agg(*nc, pack, guides);
    size_t gi = 0;
    for (auto &g : guides) {
      std::string p0 = cleanOutput ? "0" : "NONE", p1 = p0, p2 = p0;
      const double gc = gcContent(g.target.bases);
      if (cleanOutput) p0 = javaDoubleToString(gc);
      else if (gc < .25 || gc > .75) p0 = "GC_" + javaDoubleToString(gc);
      auto r = pack.guideRange();
      if (g.target.bases.substr(r.first, r.second - r.first).find("TTTT") != std::string::npos) p1 = cleanOutput ? "1" : "PolyT";
      const int inGenome = agg.inGenome[gi];
      ++gi;
      if (inGenome > 0) p2 = cleanOutput ? std::to_string(inGenome) : "IN_GENOME=" + std::to_string(inGenome);
      g.namedAnnotations["dangerous_GC"] = {p0};
      g.namedAnnotations["dangerous_polyT"] = {p1};
      g.namedAnnotations["dangerous_in_genome"] = {p2};
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------
// targetio/TabDelimitedHandler.scala
struct TabDelimitedOutput {  // :104-160
  FILE *out;
  const BitEncoding &bitEncoding;
  const BitPosition &bitPosition;
  std::vector<ScoreModel *> models;
  bool writeOTs, writePositions;
  TabDelimitedOutput(const std::string &path, const BitEncoding &be, const BitPosition &bp, std::vector<ScoreModel *> m, bool ots, bool pos)
      : bitEncoding(be), bitPosition(bp), models(std::move(m)), writeOTs(ots), writePositions(pos) {
    out = fopen(path.c_str(), "w");
    if (!out) throw std::runtime_error("cannot write " + path);
    std::string hdr = "contig\tstart\tstop\ttarget\tcontext\toverflow\torientation";
    for (auto *mdl : models)
      for (auto &c : mdl->headerColumns()) hdr += "\t" + c;
    hdr += writeOTs ? "\totCount\toffTargets\n" : "\totCount\n";
    fputs(hdr.c_str(), out);
  }
  ~TabDelimitedOutput() { close(); }
  void close() { if (out) fclose(out); out = nullptr; }

  std::string hitToOutput(const CRISPRHit &hit, uint64_t guide) const {  // CRISPRHit.toOutput :54-101
    int count = 0;
    const std::string bases = bitEncoding.bitDecodeString(hit.sequence, &count);
    std::string s = bases + "_" + std::to_string(count) + "_" +
                    std::to_string(hit.bulgeMismatches >= 0 ? hit.bulgeMismatches : bitEncoding.mismatches(guide, hit.sequence));
    // bulge extension: a fourth field names the looped-out base (R = guide base, D = genomic base); FlashFry's own
    // `score` cannot re-read such tokens -- CFD / Hsu2013 are not defined for bulged alignments
    if (hit.bulge) s += std::string("_") + ((hit.bulge & 0xC0) == 0x40 ? "R" : "D") + std::to_string(hit.bulge & 0x3F);
    if (!writePositions) return s;
    if (hit.validOffTargetCoordinates && !hit.coordinates.empty()) {
      s += "<";
      for (size_t i = 0; i < hit.coordinates.size(); ++i) {
        auto d = bitPosition.decode(hit.coordinates[i]);
        s += (i ? "|" : "") + d.contig + ":" + std::to_string(d.start) + "^" + (d.forwardStrand ? "F" : "R");
      }
      s += ">";
    }
    if (!hit.scores.empty()) {
      s += "{";
      for (size_t i = 0; i < hit.scores.size(); ++i) s += (i ? "!" : "") + hit.scores[i].first + "=" + hit.scores[i].second;
      s += "}";
    }
    return s;
  }

  void write(const CRISPRSiteOT &g) {  // :132-154
    std::string row = g.target.contig + "\t" + std::to_string(g.target.start()) + "\t" + std::to_string(g.target.start() + g.target.length()) +
                      "\t" + g.target.bases + "\t" + (g.target.sequenceContext.empty() ? "NONE" : g.target.sequenceContext) + "\t" +
                      ((g.full() || g.inheritedOverflow) ? "OVERFLOW" : "OK") + "\t" + (g.target.forwardStrand ? "FWD" : "RVS") + "\t";
    for (auto *mdl : models) {
      for (auto &c : mdl->headerColumns()) {
        auto it = g.namedAnnotations.find(c);
        std::string v;
        if (it == g.namedAnnotations.end()) v = ScoreModel::missingAnnotation;
        else for (size_t i = 0; i < it->second.size(); ++i) v += (i ? "," : "") + it->second[i];
        row += v + "\t";
      }
    }
    long long total = 0;
    for (auto &ot : g.offTargets) total += ot.getOffTargetCount();
    row += std::to_string(total);
    if (writeOTs) {
      row += "\t";
      for (size_t i = 0; i < g.offTargets.size(); ++i) row += (i ? "," : "") + hitToOutput(g.offTargets[i], g.longEncoding);
    }
    row += "\n";
    fputs(row.c_str(), out);
  }
};

inline std::vector<std::string> split(const std::string &s, char sep) {
  std::vector<std::string> out;
  size_t a = 0;
  for (;;) {
    const size_t b = s.find(sep, a);
    out.push_back(s.substr(a, b == std::string::npos ? std::string::npos : b - a));
    if (b == std::string::npos) break;
    a = b + 1;
  }
  return out;
}

// TabDelimitedInput :169-334
struct TabDelimitedInput {
  std::vector<CRISPRSiteOT> guides;
  std::vector<std::string> annotations;
  TabDelimitedInput(const std::string &path, const BitEncoding &bitEnc, BitPosition &bitPos, int maximumMismatches, bool filterOutOverflowedGuides) {
    auto lines = readLines(path);
    if (lines.empty()) throw IllegalStateException("empty input " + path);
    auto header = split(lines[0], '\t');
    static const char *def[] = {"contig", "start", "stop", "target", "context", "overflow", "orientation"};
    if (header.size() < 8) throw IllegalStateException("Header line not long enough for file: " + path);
    for (int i = 0; i < 7; ++i)
      if (header[i] != def[i]) throw IllegalStateException("Mismatched line doesn't contain the standard header tokens: " + path);
    const bool withOTs = header.size() >= 9 && header[header.size() - 2] == "otCount" && header.back() == "offTargets";
    if (!withOTs && header.back() != "otCount") throw IllegalStateException("Unable to parse out the final columns in the header");
    annotations.assign(header.begin() + 7, header.end() - (withOTs ? 2 : 1));
    for (size_t li = 1; li < lines.size(); ++li) {
      if (lines[li].empty()) continue;
      auto sp = split(lines[li], '\t');
      CRISPRSite site{sp[0], sp[3], sp[6] == "FWD", atoi(sp[1].c_str()), sp[4] == "NONE" ? std::string() : sp[4]};
      const bool isOverflowed = sp[5] != "OK";
      const int otCount = atoi(sp[7 + annotations.size()].c_str());
      CRISPRSiteOT ot{site, bitEnc.bitEncodeString(sp[3], 1), isOverflowed ? otCount : otCount + 1, isOverflowed};  // :242-249
      for (size_t a = 0; a < annotations.size(); ++a) ot.namedAnnotations[annotations[a]] = {sp[7 + a]};
      if (withOTs && sp.size() == header.size() && !sp.back().empty()) {
        for (auto &tokFull : split(sp.back(), ',')) {
          const std::string tok = tokFull.substr(0, tokFull.find('{'));
          auto f = split(tok, '_');
          if (f.size() < 3) throw IllegalStateException("Unable to parse line: " + lines[li].substr(0, 100));
          const int cnt = atoi(f[1].c_str());
          const int mm = atoi(f[2].c_str());
          if (mm > maximumMismatches) continue;  // :293
          CRISPRHit hit;
          hit.sequence = bitEnc.bitEncodeString(f[0], cnt);
          const size_t lt = tok.find('<');
          if (lt != std::string::npos) {
            for (auto &pe : split(tok.substr(lt + 1, tok.find('>') - lt - 1), '|')) {
              const size_t colon = pe.find(':'), hat = pe.find('^', colon);
              const std::string ctg = pe.substr(0, colon);
              if (std::find(bitPos.indexToContig.begin(), bitPos.indexToContig.end(), ctg) == bitPos.indexToContig.end()) bitPos.addReference(ctg);
              hit.coordinates.push_back(bitPos.encode(ctg, atoi(pe.substr(colon + 1, hat - colon - 1).c_str()), (int)f[0].size(), pe.substr(hat + 1) == "F"));
            }
          } else {
            hit.coordinates.assign((size_t)cnt, 0);
            hit.validOffTargetCoordinates = false;
          }
          const size_t lb = tokFull.find('{');  // :322-331 score pairs ride along
          if (lb != std::string::npos)
            for (auto &pair : split(tokFull.substr(lb + 1, tokFull.find('}') - lb - 1), '!')) {
              const size_t eq = pair.find('=');
              if (eq == std::string::npos) throw IllegalStateException("Score pairs should be key=value");
              hit.scores.emplace_back(pair.substr(0, eq), pair.substr(eq + 1));
            }
          if (!ot.full()) ot.addOT(std::move(hit));  // :311,:317
        }
      }
      if (!filterOutOverflowedGuides || (!ot.inheritedOverflow && !ot.full())) guides.push_back(std::move(ot));  // :259
    }
  }
};

}  // namespace flashfry
