"""Pins the CPU oracle against the reference's own golden vectors (CPU only, no GPU, no /root/reference).

Fixtures under tests/golden/ were produced by tests/golden/make_golden.py from the reference checkout:
EMX1.output (md5 895e282b..., test_data/integration_test.sh:81), EMX1.output.scored (md5 804bf3c1..., :84),
reference_unit_vectors.json (ScalaTest known answers), fake.sites.gz (TabDelimitedHanderTest.scala:40-51).
"""
import gzip
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN


@pytest.fixture(scope="module")
def vec():
    return json.load(open(os.path.join(GOLDEN, "reference_unit_vectors.json")))


def test_golden_files_keep_their_pinned_md5():
    pins = json.load(open(os.path.join(GOLDEN, "pins.json")))
    md5 = lambda p: hashlib.md5(open(os.path.join(GOLDEN, p), "rb").read()).hexdigest()
    assert md5("EMX1.output") == pins["md5_discover"] == "895e282bf486c359667e2c3e0e0e0260"
    assert md5("EMX1.output.scored") == pins["md5_scored"] == "804bf3c1ff38b077f31f12f51d733aa4"


def test_encode_decode_round_trip(oracle):
    # BitEncodingTest.scala:20-64
    rng = np.random.default_rng(7)
    for L in (20, 22, 23, 24):
        for _ in range(200):
            s = "".join("ACGT"[i] for i in rng.integers(0, 4, L))
            c = int(rng.integers(1, 32768))
            enc = oracle.encode(s, c)
            assert oracle.decode(enc, L) == (s, c)
    assert oracle.encode("ACGT", 1) == (1 << 48) | 0b00011011
    with pytest.raises(ValueError):
        oracle.encode("ACGN")
    with pytest.raises(ValueError):
        oracle.encode("A" * 25)


def test_mismatch_known_answers(oracle, vec):
    # BitEncodingTest.scala:79-151,296-307
    for c in vec["mismatch_cases"]:
        p = oracle.pack_by_name(c["pack"])
        assert oracle.mismatches(p, oracle.encode(c["a"], c["ca"]), oracle.encode(c["b"], c["cb"])) == c["mm"], c


def test_mismatch_random_vs_naive(oracle):
    # BitEncodingTest.scala:153-200 (10k random pairs, Cas9 bases 0..20 and Cpf1 bases 4..24)
    rng = np.random.default_rng(11)
    for name, lo, hi in (("SPCAS9", 0, 20), ("CPF1", 4, 24), ("SPCAS9NGG19", 0, 19)):
        p = oracle.pack_by_name(name)
        for _ in range(3000):
            a = "".join("ACGT"[i] for i in rng.integers(0, 4, p.scan_len))
            b = "".join("ACGT"[i] for i in rng.integers(0, 4, p.scan_len))
            naive = sum(1 for x, y in zip(a[lo:hi], b[lo:hi]) if x != y)
            assert oracle.mismatches(p, oracle.encode(a, int(rng.integers(1, 30000))), oracle.encode(b)) == naive


def test_bin_prefix_known_answers(oracle, vec):
    # BitEncodingTest.scala:236-359
    for c in vec["bin_cases"]:
        p = oracle.pack_by_name(c["pack"])
        assert oracle.mismatch_bin(p, c["bin"], oracle.encode(c["guide"])) == c["mm"], c


def test_cfd_known_answers(oracle, vec):
    # Doench2016CFDScoreTest.scala:32-84 ; tolerances are the reference's (1e-3); restated exact values from SURVEY 8c.4
    g20 = vec["cfd_pairs"]["guide20"]
    exact = [0.2492374732932462, 0.2445141064537618, 0.23518099545763574, 0.18765610861813575, 0.1423628278388532]
    for (ot20, expect), ex in zip(vec["cfd_pairs"]["cases"], exact):
        got = oracle.cfd_pair(oracle.encode(g20 + "AGG"), oracle.encode(ot20 + "AGG"))
        assert abs(got - expect) < 1e-3
        assert got == ex
    exact_max = [0.0, 0.5238095242619047, 0.302521008307563]
    for case, ex in zip(vec["cfd_guides"], exact_max):
        ots = np.asarray([oracle.encode(s) for s in case["off_targets"]], np.uint64)
        mx, spec, per = oracle.cfd_guide(oracle.encode(case["guide"]), ots)
        assert abs(mx - case["expected_max"]) < 1e-3
        assert mx == ex
        assert 0.0 < spec <= 1.0


def test_hsu_known_answers(oracle, vec):
    # CrisprMitEduOffTargetTest.scala:54-70
    p = oracle.pack_by_name("SPCAS9")
    h = vec["hsu"]
    ots = np.asarray([oracle.encode(s) for s in h["off_targets"]], np.uint64)
    got = oracle.hsu_guide(p, oracle.encode(h["guide"]), ots)
    assert abs(got - h["expected"]) <= h["tol"]
    assert got == 96.0618868577998
    s = h["single"]
    one = oracle.hsu_offtarget(oracle.encode(s["guide"]), oracle.encode(s["ot"]))
    assert abs(one - s["expected"]) <= s["tol"]
    assert one == 0.3640387298259494


def test_scored_golden_columns(oracle):
    """Re-score the pinned discover TSV and compare with the pinned scored TSV column by column (text-exact)."""
    pack = oracle.pack_by_name("SPCAS9NGG")
    guides = oracle.read_discover_tsv(os.path.join(GOLDEN, "EMX1.output"), pack)
    lines = open(os.path.join(GOLDEN, "EMX1.output.scored")).read().strip().split("\n")
    hdr = lines[0].split("\t")
    rows = {r.split("\t")[3]: dict(zip(hdr, r.split("\t"))) for r in lines[1:]}
    assert len(guides) == 3
    for g in guides:
        ots = np.asarray(g.targets, np.uint64)
        mx, spec, _ = oracle.cfd_guide(g.encoding, ots)
        hs = oracle.hsu_guide(pack, g.encoding, ots)
        row = rows[g.site.bases]
        assert oracle.java_double_str(mx) == row["DoenchCFD_maxOT"]
        assert oracle.java_double_str(spec) == row["DoenchCFD_specificityscore"]
        assert oracle.java_double_str(hs) == row["Hsu2013"]
        assert oracle.minot(pack, g.encoding, ots) == (row["basesDiffToClosestHit"], row["closestHitCount"], row["0-1-2-3-4_mismatch"])
        assert oracle.dangerous(pack, g.site.bases, g.encoding, ots) == (row["dangerous_GC"], row["dangerous_polyT"], row["dangerous_in_genome"])
    # values quoted in BASELINE.md section 2
    assert rows["GAGTCCGAGCAGAAGAAGAAGGG"]["DoenchCFD_maxOT"] == "0.30014429994302205"
    assert rows["AGAGTCCGAGCAGAAGAAGAAGG"]["Hsu2013"] == "98.41774847095192"


def test_fake_sites_tokens_are_mismatch_known_answers(oracle):
    """Every SEQ_count_mm token of test_data/fake.sites is a known answer for BitEncoding.mismatches."""
    pack = oracle.pack_by_name("SPCAS9")
    path = os.path.join(GOLDEN, "fake.sites.gz")
    guides = oracle.read_discover_tsv(path, pack, filter_overflow=False)
    assert len(guides) == 99
    n = 0
    for g in guides:
        for t, mm in zip(g.targets, g.recorded_mm):
            assert oracle.mismatches(pack, g.encoding, t) == mm
            n += 1
    assert n > 5000


def test_fake_sites_round_trip(oracle, tmp_path):
    """TabDelimitedHanderTest.scala:40-51: read fake.sites, write it back, byte-identical."""
    pack = oracle.pack_by_name("SPCAS9")
    raw = gzip.open(os.path.join(GOLDEN, "fake.sites.gz"), "rt").read()
    lines = raw.rstrip("\n").split("\n")
    hdr = lines[0].split("\t")
    guides = oracle.read_discover_tsv(os.path.join(GOLDEN, "fake.sites.gz"), pack, filter_overflow=False)
    contigs = []
    row_ptr, targets, mms, pos_ptr, positions = [0], [], [], [0], []
    for g in guides:
        for t, mm, pl in zip(g.targets, g.recorded_mm, g.positions):
            targets.append(t)
            mms.append(mm)
            for (ctg, st, fwd) in (pl or []):
                if ctg not in contigs:
                    contigs.append(ctg)
                positions.append(oracle.pos_encode(contigs.index(ctg) + 1, st, 23, fwd))
            pos_ptr.append(len(positions))
        row_ptr.append(len(targets))
    hits = oracle.Hits(np.asarray(row_ptr), np.asarray(targets, np.uint64), np.asarray(mms, np.uint8),
                       np.zeros(len(guides), np.int32),
                       np.asarray([g.inherited_overflow for g in guides], np.uint8),
                       np.asarray(pos_ptr), np.asarray(positions, np.uint64))
    out = tmp_path / "fake.out"
    ann = hdr[7:-2]
    oracle.write_discover_tsv(str(out), pack, [oracle.Guide(g.site, g.encoding) for g in guides], hits, contigs,
                              with_positions=True, score_columns=ann,
                              score_values=[[g.annotations[a] for a in ann] for g in guides])
    assert open(out).read() == raw


def test_block_manager_linear_vs_indexed(oracle):
    """BlockManagerTest.scala:29-63 on test_blockAACCTTGG.binary (10 130 targets x 1 position, big-endian):
    the same targets laid out as a linear and as an indexed block give identical hits (the reference test only
    compares sizes; here the full hit lists are compared)."""
    pack = oracle.pack_by_name("SPCAS9")
    raw = np.fromfile(os.path.join(GOLDEN, "test_blockAACCTTGG.binary"), dtype=">u8").astype(np.uint64)
    assert int(raw[0]) == 2 * 10130 == len(raw) - 1  # leading long = number of longs (BlockManagerTest.scala:118-131)
    targets, pos = raw[1::2], raw[2::2]
    assert ((targets >> np.uint64(48)) == 1).all()
    uniq, counts, p2 = oracle.collapse_sites(targets & np.uint64(oracle.STRING_MASK), pos)
    blocks_idx, nt = oracle.make_blocks(pack, 7, uniq, counts, p2)
    b = oracle.bin_code("AACCTTG")
    assert nt[b] == len(uniq) and int(blocks_idx[b][0]) == 2
    # the same bin as a linear block
    body = blocks_idx[b][257:]
    lin = np.concatenate([np.asarray([1], np.uint64), body])
    rng = np.random.default_rng(5)
    guides = [oracle.encode("AACCTTGG" + "".join("ACGT"[i] for i in rng.integers(0, 4, 12)) + "TGG") for _ in range(1000)]

    def run(block_for_bin):
        blocks = [np.asarray([1], np.uint64)] * (4 ** 7)
        blocks[b] = block_for_bin
        off = np.zeros(4 ** 7 + 1, np.int64)
        np.cumsum([len(x) for x in blocks], out=off[1:])
        db = oracle.Database(pack, 7, np.concatenate(blocks), off, np.zeros(4 ** 7, np.int32), ["c"])
        return oracle.discover_blocks(db, guides, 1, 2000)

    a, c = run(blocks_idx[b]), run(lin)
    assert (a.row_ptr == c.row_ptr).all() and (a.targets == c.targets).all() and (a.mismatches == c.mismatches).all()
    assert int(a.row_ptr[-1]) > 0


def test_discover_chr22_emx1_reproduces_pinned_tsv(oracle, chr22_db_path, tmp_path):
    """index -> discover on the quick-start data == md5 895e282b... (needs the locally built chr22 DB)."""
    db = oracle.read_database(chr22_db_path)
    guides = oracle.guides_from_fasta(os.path.join(GOLDEN, "EMX1_GAGTCCGAGCAGAAGAAGAAGGG.fasta"), db.pack, 6)
    for force_linear in (False, True):
        hits = oracle.discover_blocks(db, [g.encoding for g in guides], 4, 2000, force_linear=force_linear)
        out = tmp_path / "EMX1.output"
        oracle.write_discover_tsv(str(out), db.pack, guides, hits, db.contigs)
        assert oracle.md5_file(str(out)) == "895e282bf486c359667e2c3e0e0e0260"
    # and the SoA walk gives the same rows
    t, bo, _po, _pp = db.soa()
    soa = oracle.discover_soa(db.pack, 7, t, bo, [g.encoding for g in guides], 4, 2000, n_threads=2)
    assert (soa.row_ptr == hits.row_ptr).all() and (soa.targets == hits.targets).all()


def test_site_finder_known_answers(oracle, vec):
    # SimpleSiteFinderTest.scala:14-175
    for c in vec["site_finder_cases"]:
        pack = oracle.pack_by_name(c["pack"])
        sites = oracle.find_target_sites([("testContig", c["seq"])], pack, c["flank"])
        assert len(sites) == len(c["sites"]), c
        for s, (kind, a, b) in zip(sites, c["sites"]):
            want = c["seq"][a:b] if kind == "fwd" else oracle.revcomp(c["seq"][a:b])
            assert s.bases == want and s.position == a and s.forward == (kind == "fwd")
        if "context_defined" in c:
            assert [s.context is not None for s in sites] == c["context_defined"]
        if c.get("context_is_whole"):
            assert sites[0].context == c["seq"]


def test_bit_position_known_answers(oracle, vec):
    # BitPositionTest.scala:25-61 (round trip) and :63-94 (overlap is host-side annotation code, out of scope)
    for c in vec["bit_position_cases"]:
        cid = c["contigs"].index(c["contig"]) + 1
        enc = oracle.pos_encode(cid, c["start"], c["len"], c["fwd"])
        assert oracle.pos_decode(enc) == (cid, c["start"], c["len"], c["fwd"])
    assert oracle.pos_encode(2, 1000, 23, True) == (2 << 32) | 1000 | (23 << 52)
    assert oracle.pos_encode(1, 5, 23, False) >> 60 == 1


# ---- EXTENSION (no reference behaviour; parity unpinned): the bulge-mode definition the CUDA path is tested against
def test_bulge_alignment_definition_is_self_consistent(oracle):
    """ffo_bulge_align (C, base by base) == the string-surgery restatement (delete guide base q / delete target base q);
    with no bulge allowed it is BitEncoding.mismatches."""
    import random
    rnd = random.Random(5)
    pack = oracle.PACK_BY_INDEX[3]
    for _ in range(3000):
        g = "".join(rnd.choice("ACGT") for _ in range(20))
        t = list(g)
        r = rnd.random()
        q = rnd.randrange(1, 19)
        if r < 0.3:
            t = [rnd.choice("ACGT")] + list(g[:q] + g[q + 1:])
        elif r < 0.6:
            t = list(g[1:q + 1]) + [rnd.choice("ACGT")] + list(g[q + 1:])
        for _ in range(rnd.randrange(0, 5)):
            t[rnd.randrange(20)] = rnd.choice("ACGT")
        t = "".join(t)
        ge, te = oracle.encode(g + "AGG"), oracle.encode(t + "TGG", 7)
        for flags in (0, 1, 2, 3):
            assert oracle.bulge_align(ge, te, flags) == oracle.bulge_align_strings(g, t, flags)
        assert oracle.bulge_align(ge, te, 0) == (oracle.mismatches(pack, ge, te), 0, 0)


def test_bulge_discover_without_bulges_is_the_reference_search(oracle):
    """flags = 0 reduces the extension's brute force to the reference's discover (same rows, same overflow cut)."""
    import helpers
    pack = oracle.PACK_BY_INDEX[3]
    targets, seeds = helpers.family_database(oracle, seed=3, n_seeds=12, variants_per_seed=400)
    for max_ot in (5, 2000):
        ref = oracle.discover_soa(pack, 7, targets, oracle.bin_offsets_from_sorted(pack, 7, targets), seeds, 4, max_ot)
        got = oracle.discover_bulge(pack, targets, seeds, 4, max_ot, 0)
        helpers.assert_hits_equal(got, ref)
        assert not got.bulge.any()


def test_bulge_templates_are_equivalent_to_the_alignments(oracle):
    """The CUDA path searches bulges as 20-position TEMPLATES with one wildcard (csrc/ff_general.inl pattern_template):
    RNA bulge at q -> [*, g0..g(q-1), g(q+1)..g19], DNA bulge at q -> [g1..gq, *, g(q+1)..g19].  Restated here with the
    same bit formulas and checked against the string definition: the masked mismatch count of (template, target) must
    equal the mismatch count of the corresponding alignment, for every q."""
    import random
    rnd = random.Random(11)
    P = 20
    full = (1 << (2 * P)) - 1

    def enc(s):
        v = 0
        for ch in s:
            v = (v << 2) | "ACGT".index(ch)
        return v

    def template(proto, typ, q):
        lo_mask = (1 << (2 * (P - 1 - q))) - 1
        if typ == 1:
            return (((proto >> 2) & ~lo_mask & ((1 << (2 * (P - 1))) - 1)) | (proto & lo_mask)), 0
        hi_mask = full & ~((1 << (2 * (P - q))) - 1)
        return ((((proto << 2) & full) & hi_mask) | (proto & lo_mask)), q

    def masked_mm(a, b, wild):
        x = (a ^ b) & ~(3 << (2 * (P - 1 - wild)))
        return bin((x | (x >> 1)) & int("01" * P, 2)).count("1")

    ham = lambda a, b: sum(x != y for x, y in zip(a, b))
    for _ in range(400):
        g = "".join(rnd.choice("ACGT") for _ in range(P))
        t = "".join(rnd.choice("ACGT") if rnd.random() < 0.3 else c for c in g)
        for q in range(1, P - 1):
            tr, wr = template(enc(g), 1, q)
            assert masked_mm(tr, enc(t), wr) == ham(g[:q] + g[q + 1:], t[1:])
            td, wd = template(enc(g), 2, q)
            assert masked_mm(td, enc(t), wd) == ham(g[1:], t[:q] + t[q + 1:])


def test_streaming_database_writer_equals_the_object_writer(tmp_path):
    """oracle/big_db.py (the streaming writer bench.py's cold-start measurement uses for a 3e8-target database) against
    ff_oracle's per-bin writer, which the md5-pinned chr22 database comes from: same inflated blocks (linear AND indexed
    bins), same header semantics, read back by the oracle's reader."""
    import numpy as np
    from oracle import big_db, ff_oracle as o
    pack = o.PACK_BY_INDEX[3]
    rng = np.random.default_rng(4)
    # 40 crowded 7-mer bins (> 500 targets: indexed blocks) on a sparse background (linear blocks, many empty bins)
    crowded = [(np.uint64(int(b)) << np.uint64(28)) | rng.integers(0, 1 << 28, 900, dtype=np.uint64) for b in rng.integers(0, 4 ** 7, 40)]
    seq = np.unique(np.concatenate(crowded + [rng.integers(0, 1 << 42, 5000, dtype=np.uint64)]))
    seq = (seq << np.uint64(4)) | np.uint64(0xA)
    counts = rng.integers(1, 5, len(seq)).astype(np.int64)
    counts[::997] = 700
    targets = seq | (counts.astype(np.uint64) << np.uint64(48))
    st = big_db.write_big_database(str(tmp_path / "big"), pack, targets, threads=4, bins_per_run=100)
    assert st["indexed_bins"] >= 30 and st["positions"] == int(counts.sum())
    pos = big_db.synthetic_positions(0, int(counts.sum()))
    blocks, ntargets = o.make_blocks(pack, 7, seq, counts, pos)
    o.write_database(str(tmp_path / "small"), pack, 7, blocks, ntargets, ["chrSynth%d" % (i + 1) for i in range(24)])
    a, b = o.read_database(str(tmp_path / "big")), o.read_database(str(tmp_path / "small"))
    for x, y in zip(a.soa(), b.soa()):
        assert (x == y).all()
    pa, _ = o.bgzf_inflate_all(str(tmp_path / "big"))
    pb, _ = o.bgzf_inflate_all(str(tmp_path / "small"))
    assert pa == pb  # the same block stream, byte for byte (member boundaries and compression level may differ)
