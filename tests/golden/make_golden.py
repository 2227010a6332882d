#!/usr/bin/env python3
"""Generate the committed golden fixtures from the reference checkout (run in the authoring container).

    python tests/golden/make_golden.py            # needs /root/reference (read-only)

What it does
  1. builds the chr22 spCas9-NGG database with the oracle's `index` restatement from
     test_data/quickstart_data.tar.gz, runs the oracle's `discover` on the EMX1 FASTA and checks
     md5(EMX1.output) == 895e282bf486c359667e2c3e0e0e0260      (test_data/integration_test.sh:81);
  2. runs the oracle's `score` restatement (CFD + Hsu2013 from ff_oracle.c; minot / dangerous from
     ff_oracle.py; Doench2014 on-target computed here from the coefficient table parsed out of the
     reference source at generation time -- it is out of scope and is NOT copied into the repo) and checks
     md5(EMX1.output.scored) == 804bf3c1ff38b077f31f12f51d733aa4  (integration_test.sh:84);
  3. writes both TSVs, the reference's unit-test vectors (sequences + expected values pulled out of the
     ScalaTest sources with regexes) and a gzip of test_data/fake.sites under tests/golden/;
  4. leaves the built chr22 database in tests/golden/_chr22/ (git-ignored; travels to the GPU box).

The fixtures are data (sequences, numbers, md5s); no reference source code is copied.
"""
import gzip
import json
import math
import os
import re
import shutil
import sys
import tarfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
from oracle import ff_oracle as o  # noqa: E402

REF = os.environ.get("FLASHFRY_REFERENCE", "/root/reference")
MD5_DISCOVER = "895e282bf486c359667e2c3e0e0e0260"
MD5_SCORED = "804bf3c1ff38b077f31f12f51d733aa4"


def doench2014_table():
    src = open(os.path.join(REF, "src/main/scala/scoring/Doench2014OnTarget.scala")).read()
    tab = {m.group(1): float(m.group(2).replace(" ", "")) for m in re.finditer(r'"([ACGT]{1,2}\d+)"\s*->\s*(-?\s*[0-9.]+)', src)}
    gc_low = float(re.search(r"gc_low\s*=\s*(-?[0-9.]+)", src).group(1))
    gc_high = float(re.search(r"gc_high\s*=\s*(-?[0-9.]+)", src).group(1))
    intercept = float(re.search(r"intercept\s*=\s*(-?[0-9.]+)", src).group(1))
    return tab, gc_low, gc_high, intercept


def doench2014(ctx30, tab, gc_low, gc_high, intercept):
    """scoring/Doench2014OnTarget.scala:115-150 (out of scope for the GPU path; needed only for the md5 pin)."""
    assert len(ctx30) == 30
    gc = sum(1 for b in ctx30[4:24] if b in "CG")
    score = intercept + abs(gc - 10) * gc_low
    if gc > 10:
        score = intercept + (gc - 10) * gc_high
    for i, b in enumerate(ctx30):
        score = score + tab.get(b + str(i), 0.0)
        if i < 29:
            d = b + ctx30[i + 1] + str(i)
            if d in tab:
                score += tab[d]
    return 1.0 / (1.0 + math.exp(-1.0 * score))


def scala_strings(block):
    return re.findall(r'"([ACGT]{20,24})"', block)


def unit_vectors():
    """Pull the known-answer vectors out of the reference's ScalaTest sources."""
    out = {}
    t = open(os.path.join(REF, "src/test/scala/scoring/Doench2016CFDScoreTest.scala")).read()
    pair_guide = re.search(r'val guide = "([ACGT]{20})"', t).group(1)
    pairs = re.findall(r'scoreCFD\(guide,"([ACGT]{20})"\)\) should be\(([0-9.]+)', t)
    out["cfd_pairs"] = {"guide20": pair_guide, "cases": [[s, float(v)] for s, v in pairs]}
    cases = []
    for m in re.finditer(r'val offTargetList\d* = Array\[String\]\((.*?)\)\s*\n\s*\n\s*val otListLong.*?CRISPRSite\("test", "([ACGT]{23})".*?should be\(([0-9.]+)',
                         t, re.S):
        cases.append({"guide": m.group(2), "off_targets": scala_strings(m.group(1)), "expected_max": float(m.group(3))})
    assert len(cases) == 3, len(cases)
    out["cfd_guides"] = cases
    t = open(os.path.join(REF, "src/test/scala/scoring/CrisprMitEduOffTargetTest.scala")).read()
    guide = re.search(r'new CRISPRSite\("1", "([ACGT]{23})"', t).group(1)
    head = t[:t.index('"CrisprMitEduOffTargetTest" should')]
    ots = re.findall(r'StringCount\("([ACGT]{23})",1\)\),Array', head)
    assert len(ots) == 30, len(ots)
    single = re.search(r'score (\w{23}) to correctly (\w{23})', t)
    out["hsu"] = {"guide": guide, "off_targets": ots, "expected": 96.0, "tol": 1.0,
                  "single": {"guide": single.group(1), "ot": single.group(2), "expected": 0.36403873, "tol": 0.1}}
    # BitEncodingTest known answers (src/test/scala/bitcoding/BitEncodingTest.scala:79-151,236-359)
    sp = lambda s: s.replace(" ", "")
    out["mismatch_cases"] = [
        {"pack": "SPCAS9", "a": sp("AAAAA CCCCC GGGGG TTTTA GGG"), "ca": 1000, "b": sp("AAAAA CCCCC GGGGG TTTTA GGG"), "cb": 1000, "mm": 0},
        {"pack": "SPCAS9", "a": sp("AAAAA CCCCC GGGGG TTTTA GGG"), "ca": 1000, "b": sp("AAAAA CCCCC GGGGG TTTTA GGG"), "cb": 1001, "mm": 0},
        {"pack": "SPCAS9", "a": sp("AAAAA CCCCC GGGGG TTTTA GGG"), "ca": 1000, "b": sp("TAAAA CCCCC GGGGG TTTTA GGG"), "cb": 1001, "mm": 1},
        {"pack": "SPCAS9", "a": sp("AAAAA CCCCC GGGGG TTTTA GGG"), "ca": 1000, "b": sp("TTTTT TTTTT AAAAA GGGGG GGG"), "cb": 1001, "mm": 20},
        {"pack": "SPCAS9", "a": sp("AAAAA CCCCC GGGGG AAAAT AGG"), "ca": 1000, "b": sp("AAAAA CCCCC GGGGG TTTTA GGG"), "cb": 1001, "mm": 5},
        {"pack": "SPCAS9", "a": sp("AAAAA CCCCC GGGGG AAAAT AAG"), "ca": 1000, "b": sp("AAAAA CCCCC GGGGG TTTTA GGG"), "cb": 1001, "mm": 5},
        {"pack": "SPCAS9NGG", "a": sp("GAGTC CGAGC AGAAG AAGAA GGG"), "ca": 1, "b": sp("GAATC ATAGC AGAAG ATGAA AGG"), "cb": 1001, "mm": 4},
    ]
    out["bin_cases"] = [
        {"pack": "SPCAS9", "guide": sp("AAAAA CCCCC GGGGG TTTTA GGG"), "bin": "AAAAA", "mm": 0},
        {"pack": "SPCAS9", "guide": sp("TTAAT CCCCC GGGGG TTTTA GGG"), "bin": "TTTTT", "mm": 2},
        {"pack": "SPCAS9", "guide": sp("AAAAA AAAAC GGGGG TTTTA GGG"), "bin": "AAAAAAAAA", "mm": 0},
        {"pack": "SPCAS9NGG", "guide": sp("GAGTC CGAGC AGAAG AAGAA GGG"), "bin": "GAGTCCG", "mm": 0},
        {"pack": "SPCAS9NGG", "guide": sp("GGCTC CGAGC AGAAG AAGAA GGG"), "bin": "GAGTCCG", "mm": 2},
        {"pack": "SPCAS9NGG", "guide": sp("GGCTC CGAGC AGAAG AAGAA GGG"), "bin": "AAAAAAA", "mm": 7},
        {"pack": "CPF1", "guide": sp("TTTT CGAGC AGAAG AAGAA GGGAC"), "bin": "CGAGCAG", "mm": 0},
        {"pack": "CPF1", "guide": sp("TTTT CGAGC AGAAG AAGAA GGGAC"), "bin": "CAAGCAG", "mm": 1},
        {"pack": "CPF1", "guide": sp("TTTT CGAGC AGAAG AAGAA GGGAC"), "bin": "AGAGCAA", "mm": 2},
    ]
    # SimpleSiteFinderTest known answers (src/test/scala/reference/SimpleSiteFinderTest.scala:14-175):
    # input string, pack, flank -> expected site bases (rc = reverse complement of the slice) / context presence
    out["site_finder_cases"] = [
        {"pack": "SPCAS9NGG", "flank": 0, "seq": sp("ATTTA AAAAA CCCCC AAAAA GGG"), "sites": [["fwd", 0, 23]], "context_defined": [True]},
        {"pack": "SPCAS9NGG", "flank": 8, "seq": sp("ATA ATATA ATTTA AAAAA TTTTT AAAAA AGG AATTA AAT"), "sites": [["fwd", 8, 31]], "context_is_whole": True},
        {"pack": "SPCAS9NGG", "flank": 0, "seq": sp("CCTTA AAAAA CCCCC AAAAA AAA"), "sites": [["rc", 0, 23]]},
        {"pack": "SPCAS9NGG", "flank": 0, "seq": sp("A ATTTA AAAAA CCCCC AAAAA GGG"), "sites": [["fwd", 0, 23], ["fwd", 1, 24]]},
        {"pack": "SPCAS9NAG", "flank": 0, "seq": sp("ATTTA AAAAA CCCCC AAAAA GAG"), "sites": [["fwd", 0, 23]]},
        {"pack": "SPCAS9NAG", "flank": 0, "seq": sp("CTTTA AAAAA CCCCC AAAAA AAA"), "sites": [["rc", 0, 23]]},
        {"pack": "SPCAS9", "flank": 0, "seq": sp("A ATTTA AAAAA CCCCC AAAAA AGG"), "sites": [["fwd", 0, 23], ["fwd", 1, 24]]},
        {"pack": "SPCAS9NGG", "flank": 0, "seq": sp("AAATA AAAAA CCCCC AAAAA GGG"), "sites": [["fwd", 0, 23]]},
        {"pack": "CPF1", "flank": 0, "seq": sp("TTTTA ATTTA AAAAA CCCCC AATTT"), "sites": [["fwd", 0, 24], ["fwd", 1, 25]]},
        {"pack": "CPF1", "flank": 0, "seq": sp("TAATA ATTTA AAAAA CCCCC AAAAA"), "sites": [["rc", 0, 24], ["rc", 1, 25]]},
        {"pack": "SPCAS9NGG", "flank": 1, "seq": sp("ATTTA AAAAA CCCCC AAAAA GGG"), "sites": [["fwd", 0, 23]], "context_defined": [False]},
    ]
    st = open(os.path.join(REF, "src/test/scala/reference/SimpleSiteFinderTest.scala")).read().replace(" ", "")
    for c in out["site_finder_cases"]:
        assert c["seq"] in st, c
    # BitPositionTest.scala:25-94
    out["bit_position_cases"] = [{"contigs": ["chr1", "chr2"], "contig": "chr2", "start": 1000, "len": 23, "fwd": True},
                                 {"contigs": ["chr1", "chr2"], "contig": "chr2", "start": 102200, "len": 23, "fwd": False}]
    # sanity: the hand-entered BitEncodingTest cases must literally occur in the reference test source
    bt = open(os.path.join(REF, "src/test/scala/bitcoding/BitEncodingTest.scala")).read().replace(" ", "")
    for c in out["mismatch_cases"]:
        assert c["a"] in bt and c["b"] in bt, c
    for c in out["bin_cases"]:
        assert c["guide"] in bt and '"%s"' % c["bin"] in bt, c
    return out


def main():
    work = os.path.join(HERE, "_chr22")
    os.makedirs(work, exist_ok=True)
    with tarfile.open(os.path.join(REF, "test_data/quickstart_data.tar.gz")) as tf:
        tf.extractall(work)
    fasta = os.path.join(work, "EMX1_GAGTCCGAGCAGAAGAAGAAGGG.fasta")
    dbp = os.path.join(work, "chr22_cas9ngg_database")
    stats = o.build_database(os.path.join(work, "chr22.fa.gz"), dbp, "spcas9ngg")
    print("index:", stats)
    os.remove(os.path.join(work, "chr22.fa.gz"))  # 11 MB of input we no longer need next to the built DB
    shutil.copy(fasta, os.path.join(HERE, os.path.basename(fasta)))

    db = o.read_database(dbp)
    pack = db.pack
    guides = o.guides_from_fasta(fasta, pack, 6)
    hits = o.discover_blocks(db, [g.encoding for g in guides], 4, 2000)
    disc = os.path.join(HERE, "EMX1.output")
    o.write_discover_tsv(disc, pack, guides, hits, db.contigs)
    assert o.md5_file(disc) == MD5_DISCOVER, o.md5_file(disc)
    print("discover md5 OK", MD5_DISCOVER)
    o.write_discover_tsv(os.path.join(HERE, "EMX1.output.with_positions"), pack, guides, hits, db.contigs, with_positions=True)

    # ---- score: doench2014ontarget,doench2016cfd,dangerous,hsu2013,minot (integration_test.sh:45-50)
    tab, gl, gh, ic = doench2014_table()
    tg = o.read_discover_tsv(disc, pack)
    tg.sort(key=lambda g: g.site.position)
    import numpy as np
    cols = ["Doench2014OnTarget", "DoenchCFD_maxOT", "DoenchCFD_specificityscore", "dangerous_GC", "dangerous_polyT",
            "dangerous_in_genome", "Hsu2013", "basesDiffToClosestHit", "closestHitCount", "0-1-2-3-4_mismatch"]
    vals, per_ot = [], []
    row_ptr, targets, mms = [0], [], []
    for g in tg:
        ots = np.asarray(g.targets, np.uint64)
        ctx = g.site.context
        gp = (len(ctx) - len(g.site.bases)) // 2
        d14 = doench2014(ctx[gp - 4: gp + len(g.site.bases) + 3], tab, gl, gh, ic)
        mx, spec, per = o.cfd_guide(g.encoding, ots)
        hs = o.hsu_guide(pack, g.encoding, ots)
        vals.append([o.java_double_str(d14), o.java_double_str(mx), o.java_double_str(spec),
                     *o.dangerous(pack, g.site.bases, g.encoding, ots), o.java_double_str(hs),
                     *o.minot(pack, g.encoding, ots)])
        per_ot.append(["{Doench2016CFDScore=%s}" % o.java_double_str(p) if p == p else None for p in per])
        targets += list(g.targets)
        mms += g.recorded_mm
        row_ptr.append(len(targets))
    h2 = o.Hits(np.asarray(row_ptr, np.int64), np.asarray(targets, np.uint64), np.asarray(mms, np.uint8),
                np.zeros(len(tg), np.int32), np.zeros(len(tg), np.uint8))
    gl2 = [o.Guide(g.site, g.encoding) for g in tg]
    scored = os.path.join(HERE, "EMX1.output.scored")
    o.write_discover_tsv(scored, pack, gl2, h2, db.contigs, score_columns=cols, score_values=vals, write_ots=False)
    assert o.md5_file(scored) == MD5_SCORED, o.md5_file(scored)
    print("score md5 OK", MD5_SCORED)
    # the with-OTs flavour: the reference's pinned md5 (integration_test.sh:87) predates the per-OT {..} suffix
    with_ots = os.path.join(HERE, "EMX1.output.scored_with_ots")
    o.write_discover_tsv(with_ots, pack, gl2, h2, db.contigs, score_columns=cols, score_values=vals, write_ots=True,
                         per_ot_scores=per_ot)
    stripped = re.sub(r"\{[^}]*\}", "", open(with_ots).read())
    import hashlib
    print("scored_with_ots md5 (per-OT suffix stripped):", hashlib.md5(stripped.encode()).hexdigest(),
          "reference pin a5a67351bc389b4b6d0c944588d2749b")

    with open(os.path.join(HERE, "reference_unit_vectors.json"), "w") as fh:
        json.dump(unit_vectors(), fh, indent=1)
    with open(os.path.join(REF, "test_data/fake.sites"), "rb") as src, gzip.GzipFile(os.path.join(HERE, "fake.sites.gz"), "wb", mtime=0) as dst:
        shutil.copyfileobj(src, dst)
    shutil.copy(os.path.join(REF, "test_data/test_blockAACCTTGG.binary"), os.path.join(HERE, "test_blockAACCTTGG.binary"))
    json.dump({"md5_discover": MD5_DISCOVER, "md5_scored": MD5_SCORED, "chr22_index_stats": stats,
               "emx1_compares": {"n_compares": hits.n_compares, "n_target_compares": hits.n_target_compares,
                                 "bins_visited": hits.bins_visited}},
              open(os.path.join(HERE, "pins.json"), "w"), indent=1)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
