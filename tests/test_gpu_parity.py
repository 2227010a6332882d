"""GPU parity: the CUDA path, called through the C ABI (ctypes), against the CPU oracle on the same inputs.
Bit-exact for hit sets, database order, mismatch integers, overflow cut, positions; bit-exact doubles for the scorers
(contract: 1e-6).  Run on the B200 box:  python -m pytest tests -m gpu -x -q
"""
import gzip
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
import helpers

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def small_ctx(ff, small_db):
    ctx = ff.Context(0)
    ctx.load_database(small_db[0])
    yield ctx
    ctx.close()


def test_loader_reads_flashfry_format(small_ctx, small_db, oracle):
    dbp, db, stats = small_db
    info = small_ctx.info()
    targets, bin_off, pos_off, positions = db.soa()
    assert info.enzyme_index == 3 and info.scan_len == 23 and info.cmp_mask == 0x3FFFFFFFFFC0
    assert info.n_targets == len(targets) == stats["targets"]
    assert info.n_positions == len(positions) == stats["sites"]
    assert (small_ctx.copy_targets() == targets).all()
    assert small_ctx.contigs() == db.contigs == ["ctg1_test", "ctg2_test"]
    assert stats["max_count"] > 100  # the repeat family made it in


@pytest.mark.parametrize("k", [0, 1, 2, 3, 4, 5])
def test_discover_matches_oracle_small(small_ctx, small_db, oracle, k):
    _, db, _ = small_db
    targets = db.soa()[0]
    guides = np.concatenate([helpers.random_guides(oracle, db.pack, 7 + k, 150),
                             helpers.planted_guides(db.pack, targets, 70 + k, 150, max_subs=5)])
    ref = oracle.discover_blocks(db, guides, k, 2000)
    got = small_ctx.discover(guides, k, 2000, positions=True)
    helpers.assert_hits_equal(got, ref, check_positions=True)
    if k >= 3:
        assert int(ref.row_ptr[-1]) > 100


@pytest.mark.parametrize("max_ot", [0, 1, 3, 50, 2000])
def test_overflow_cut_matches_reference_rule(small_ctx, small_db, oracle, max_ot):
    """Guides planted in the repeat family overflow; the kept list is the shortest database-order prefix whose
    summed occurrence count reaches maximumOffTargets (ResultsAggregator.scala:61-69, CRISPRSiteOT.scala:39-46)."""
    _, db, _ = small_db
    targets = db.soa()[0]
    heavy = targets[(targets >> np.uint64(48)) > 50]
    assert len(heavy) > 10
    guides = np.concatenate([helpers.planted_guides(db.pack, heavy, 5, 60, max_subs=2),
                             helpers.planted_guides(db.pack, targets, 6, 60, max_subs=3)])
    ref = oracle.discover_blocks(db, guides, 4, max_ot)
    got = small_ctx.discover(guides, 4, max_ot, positions=True)
    helpers.assert_hits_equal(got, ref, check_positions=True)
    if 0 < max_ot <= 50:
        assert ref.overflowed.sum() > 30
        # a guide whose first hit alone exceeds the limit keeps exactly that hit
        firsts = [g for g in range(len(guides)) if ref.row_ptr[g + 1] - ref.row_ptr[g] == 1 and ref.total_count[g] > max_ot]
        assert len(firsts) > 0
    if max_ot == 0:
        assert int(got.row_ptr[-1]) == 0 and got.overflowed.all()


def test_edge_cases(small_ctx, small_db, oracle):
    _, db, _ = small_db
    # empty guide list
    got = small_ctx.discover(np.zeros(0, np.uint64), 4, 2000)
    assert got.n_guides == 0 and len(got.targets) == 0
    # duplicated guides each get their own identical row; a far-away guide gets an empty row
    t = db.soa()[0]
    g0 = (int(t[1234]) & 0xFFFFFFFFFFFF) | (1 << 48)
    lonely = oracle.encode("ACGTACGTACGTACGTACGTAGG")
    guides = np.asarray([g0, lonely, g0, g0], np.uint64)
    ref = oracle.discover_blocks(db, guides, 2, 2000)
    got = small_ctx.discover(guides, 2, 2000)
    helpers.assert_hits_equal(got, ref)
    assert (got.row(0)[0] == got.row(2)[0]).all() and int(got.row(0)[1].min()) == 0
    # a single guide
    one = small_ctx.discover(guides[:1], 4, 2000)
    helpers.assert_hits_equal(one, oracle.discover_blocks(db, guides[:1], 4, 2000))


def test_errors_are_loud(ff, tmp_path):
    ctx = ff.Context(0)
    with pytest.raises(ff.FlashFryError) as e:
        ctx.discover(np.asarray([1 << 48], np.uint64))
    assert e.value.code == -6  # FF_ENODB
    bad = tmp_path / "bad.header"
    bad.write_text("12345\n1\n3\n16384\n")
    (tmp_path / "bad").write_bytes(b"")
    with pytest.raises(ff.FlashFryError) as e:
        ctx.load_database(str(tmp_path / "bad"), str(bad))
    assert e.value.code == -5 and "magic number" in str(e.value)
    with pytest.raises(ff.FlashFryError):
        ctx.load_database(str(tmp_path / "missing"))
    # unsorted arrays are rejected instead of silently mis-indexed
    with pytest.raises(ff.FlashFryError):
        ctx.load_database_arrays(3, np.asarray([(1 << 48) | 50, (1 << 48) | 40], np.uint64))
    ctx.close()


def test_arrays_loader_and_other_enzymes(ff, oracle, tmp_path):
    """spCas9 (NGG+NAG) and the 19-mer pack (scan length 22) through the same kernels."""
    for enzyme, seed in (("spcas9", 31), ("spcas9ngg19", 32), ("spcas9nag", 33)):
        contigs = helpers.random_genome(seed, 120_000, repeat_unit=50, n_repeats=100)
        fa = str(tmp_path / (enzyme + ".fa"))
        helpers.write_fasta(fa, contigs)
        dbp = str(tmp_path / (enzyme + "_db"))
        oracle.build_database(fa, dbp, enzyme)
        db = oracle.read_database(dbp)
        targets, _bo, _po, positions = db.soa()
        with ff.Context(0) as ctx:
            ctx.load_database_arrays(db.pack.index, targets, positions, db.contigs)
            guides = helpers.planted_guides(db.pack, targets, seed, 200, max_subs=4)
            for k in (2, 4):
                ref = oracle.discover_blocks(db, guides, k, 2000)
                got = ctx.discover(guides, k, 2000, positions=True)
                helpers.assert_hits_equal(got, ref, check_positions=True)


def test_chr22_emx1_golden_tsv(ff, oracle, chr22_db_path, tmp_path):
    """configs[0]: EMX1 vs the chr22 spCas9-NGG database -> the md5-pinned discover TSV, hits from the GPU."""
    with ff.Context(0) as ctx:
        ctx.load_database(chr22_db_path)
        info = ctx.info()
        assert info.n_targets == 4189398 and info.n_positions == 4966778
        pack = oracle.PACK_BY_INDEX[info.enzyme_index]
        guides = oracle.guides_from_fasta(os.path.join(GOLDEN, "EMX1_GAGTCCGAGCAGAAGAAGAAGGG.fasta"), pack, 6)
        got = ctx.discover([g.encoding for g in guides], 4, 2000, positions=True)
        out = tmp_path / "EMX1.output"
        oracle.write_discover_tsv(str(out), pack, guides, got, ctx.contigs())
        assert oracle.md5_file(str(out)) == "895e282bf486c359667e2c3e0e0e0260"
        assert open(out).read() == open(os.path.join(GOLDEN, "EMX1.output")).read()
        out2 = tmp_path / "EMX1.pos"
        oracle.write_discover_tsv(str(out2), pack, guides, got, ctx.contigs(), with_positions=True)
        assert open(out2).read() == open(os.path.join(GOLDEN, "EMX1.output.with_positions")).read()
        # fused scores == the md5-pinned scored TSV columns
        hits, cmax, cspec, hsu = ctx.discover_score([g.encoding for g in guides], 4, 2000)
        lines = open(os.path.join(GOLDEN, "EMX1.output.scored")).read().strip().split("\n")
        hdr = lines[0].split("\t")
        rows = {r.split("\t")[3]: dict(zip(hdr, r.split("\t"))) for r in lines[1:]}
        for i, g in enumerate(guides):
            row = rows[g.site.bases]
            assert oracle.java_double_str(cmax[i]) == row["DoenchCFD_maxOT"]
            assert oracle.java_double_str(cspec[i]) == row["DoenchCFD_specificityscore"]
            assert oracle.java_double_str(hsu[i]) == row["Hsu2013"]


def test_chr22_1000_guides_config(ff, oracle, chr22_db_path):
    """configs[1]: 1 000 synthetic (seed 1001) + 100 planted (seed 1002) guides vs chr22, <=4 mismatches."""
    db = oracle.read_database(chr22_db_path)
    targets, bin_off, _po, _pp = db.soa()
    guides = np.concatenate([helpers.random_guides(oracle, db.pack, 1001, 1000),
                             helpers.planted_guides(db.pack, targets, 1002, 100, max_subs=4)])
    ref = oracle.discover_soa(db.pack, 7, targets, bin_off, guides, 4, 2000, n_threads=os.cpu_count() or 1)
    with ff.Context(0) as ctx:
        ctx.load_database(chr22_db_path)
        got = ctx.discover(guides, 4, 2000)
        helpers.assert_hits_equal(got, ref)
        assert ref.overflowed.sum() >= 1 or int(ref.total_count.max()) > 100
        # the same batch through the cell-major kernel and through the windowed general path: a real genome's skewed
        # buckets (repeats, poly-N runs of the seed keys) instead of the uniform synthetic index
        for opts in ({"scan_kernel": 2}, {"force_general": 1, "window_cells": 3}):
            with ctx.options(**opts):
                helpers.assert_hits_equal(ctx.discover(guides, 4, 2000), ref)
        planted = helpers.planted_guides(db.pack, targets, 1003, 300, max_subs=2)
        ref_b = oracle.discover_bulge(db.pack, targets, planted[:6], 3, 2000, 3, n_threads=os.cpu_count() or 1)
        got_b = ctx.discover_bulge(planted, 3, 2000, 3)
        for g in range(6):
            lo, hi = int(got_b.row_ptr[g]), int(got_b.row_ptr[g + 1])
            rlo, rhi = int(ref_b.row_ptr[g]), int(ref_b.row_ptr[g + 1])
            assert (got_b.targets[lo:hi] == ref_b.targets[rlo:rhi]).all() and (got_b.bulge[lo:hi] == ref_b.bulge[rlo:rhi]).all()
            assert (got_b.mismatches[lo:hi] == ref_b.mismatches[rlo:rhi]).all()


def test_scorers_bit_exact_on_fake_sites(ff, oracle):
    """99 guides with real hg19 hit lists (test_data/fake.sites): CFD max / specificity / per-OT and Hsu2013 from the
    GPU equal the oracle's doubles bit for bit (contract: 1e-6)."""
    pack = oracle.pack_by_name("SPCAS9")
    guides = oracle.read_discover_tsv(os.path.join(GOLDEN, "fake.sites.gz"), pack, filter_overflow=False)
    enc = np.asarray([g.encoding for g in guides], np.uint64)
    row_ptr, targets = [0], []
    for g in guides:
        targets += g.targets
        row_ptr.append(len(targets))
    with ff.Context(0) as ctx:
        cmax, cspec, hsu, per = ctx.score(enc, row_ptr, targets)
    t = np.asarray(targets, np.uint64)
    for i, g in enumerate(guides):
        ots = t[row_ptr[i]:row_ptr[i + 1]]
        mx, sp, p = oracle.cfd_guide(g.encoding, ots)
        hs = oracle.hsu_guide(pack, g.encoding, ots)
        assert abs(cmax[i] - mx) <= 1e-6 and abs(cspec[i] - sp) <= 1e-6 and abs(hsu[i] - hs) <= 1e-6
        assert cmax[i] == mx and cspec[i] == sp and hsu[i] == hs
        got_p = per[row_ptr[i]:row_ptr[i + 1]]
        assert ((got_p == p) | (np.isnan(got_p) & np.isnan(p))).all()


def test_scorers_reference_unit_vectors(ff, oracle):
    vec = json.load(open(os.path.join(GOLDEN, "reference_unit_vectors.json")))
    with ff.Context(0) as ctx:
        for case, exact in zip(vec["cfd_guides"], [0.0, 0.5238095242619047, 0.302521008307563]):
            ots = [oracle.encode(s) for s in case["off_targets"]]
            cmax, _sp, _h, _p = ctx.score([oracle.encode(case["guide"])], [0, len(ots)], ots)
            assert abs(cmax[0] - case["expected_max"]) < 1e-3 and cmax[0] == exact
        h = vec["hsu"]
        ots = [oracle.encode(s) for s in h["off_targets"]]
        _a, _b, hsu, _p = ctx.score([oracle.encode(h["guide"])], [0, len(ots)], ots)
        assert abs(hsu[0] - 96.0) <= 1.0 and hsu[0] == 96.0618868577998
        # guides with no off-targets: specificity 1.0, max 0.0, Hsu 100.0
        cmax, cspec, hsu, _ = ctx.score([oracle.encode(h["guide"])], [0, 0], [])
        assert (cmax[0], cspec[0], hsu[0]) == (0.0, 1.0, 100.0)


def test_synthetic_database_properties(ff, oracle):
    """The on-device generator used by bench.py: sorted, distinct, NGG, counts in range; discover on it agrees with the
    oracle's reference-order walk over the same targets."""
    with ff.Context(0) as ctx:
        ctx.synth_database(3, 2_000_000, 3001)
        info = ctx.info()
        t = ctx.copy_targets()
        assert len(t) == info.n_targets and 1_990_000 < len(t) <= 2_000_000
        seq = t & np.uint64(0xFFFFFFFFFFFF)
        assert (np.diff(seq.astype(np.int64)) > 0).all()
        assert ((seq & np.uint64(0xF)) == 0xA).all()
        cnt = (t >> np.uint64(48)).astype(np.int64)
        assert cnt.min() >= 1 and cnt.max() <= 32767 and 0.92 < (cnt == 1).mean() < 0.96 and (cnt >= 500).sum() > 500
        pack = oracle.PACK_BY_INDEX[3]
        guides = np.concatenate([helpers.random_guides(oracle, pack, 3002, 300), helpers.planted_guides(pack, t, 3003, 100)])
        bin_off = oracle.bin_offsets_from_sorted(pack, 7, t)
        ref = oracle.discover_soa(pack, 7, t, bin_off, guides, 4, 2000, n_threads=os.cpu_count() or 1)
        got = ctx.discover(guides, 4, 2000)
        helpers.assert_hits_equal(got, ref)
        assert ref.overflowed.sum() > 0


CLI = os.path.join(os.path.dirname(GOLDEN), "..", "flashfry_b200", "flashfry_b200_cli")


def test_cli_discover_and_score_reproduce_the_pinned_files(oracle, chr22_db_path, tmp_path):
    """The drop-in CLI end to end on the quick-start data: `discover` output is byte-identical to the md5-pinned
    EMX1.output (integration_test.sh:81); `score` reproduces every CFD / Hsu2013 / minot / dangerous column of the
    md5-pinned EMX1.output.scored (:84) and the per-off-target CFD annotations of the --includeOTs flavour."""
    import subprocess
    fasta = os.path.join(GOLDEN, "EMX1_GAGTCCGAGCAGAAGAAGAAGGG.fasta")
    out = str(tmp_path / "EMX1.output")
    r = subprocess.run([CLI, "discover", "--database", chr22_db_path, "--fasta", fasta, "--output", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert oracle.md5_file(out) == "895e282bf486c359667e2c3e0e0e0260"
    outp = str(tmp_path / "EMX1.output.with_positions")
    r = subprocess.run([CLI, "discover", "-database", chr22_db_path, "-fasta", fasta, "-output", outp, "-positionOutput"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(outp).read() == open(os.path.join(GOLDEN, "EMX1.output.with_positions")).read()

    scored = str(tmp_path / "EMX1.output.scored")
    r = subprocess.run([CLI, "score", "--input", out, "--output", scored, "--scoringMetrics", "doench2016cfd,DanGerous,hsu2013,minot",
                        "--database", chr22_db_path], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr

    def table(path):
        lines = open(path).read().rstrip("\n").split("\n")
        hdr = lines[0].split("\t")
        return hdr, [dict(zip(hdr, ln.split("\t"))) for ln in lines[1:]]
    ghdr, grows = table(os.path.join(GOLDEN, "EMX1.output.scored"))
    hdr, rows = table(scored)
    assert [c for c in ghdr if c != "Doench2014OnTarget"] == hdr  # same columns, same order, minus the out-of-scope metric
    assert len(rows) == len(grows) == 3
    for a, b in zip(rows, grows):
        for c in hdr:
            assert a[c] == b[c], c
    with_ots = str(tmp_path / "EMX1.output.scored_with_ots")
    r = subprocess.run([CLI, "score", "--input", out, "--output", with_ots, "--scoringMetrics", "doench2016cfd,dangerous,hsu2013,minot",
                        "--database", chr22_db_path, "--includeOTs"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    _, wrows = table(with_ots)
    _, gwrows = table(os.path.join(GOLDEN, "EMX1.output.scored_with_ots"))
    for a, b in zip(wrows, gwrows):
        assert a["offTargets"] == b["offTargets"] and a["otCount"] == b["otCount"]
    # an out-of-scope metric is an error, not a silent skip
    r = subprocess.run([CLI, "score", "--input", out, "--output", scored, "--scoringMetrics", "doench2014ontarget", "--database", chr22_db_path],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "outside the GPU hot path" in r.stderr


def test_cpf1_five_prime_pam(ff, oracle, tmp_path):
    """Cpf1 (TTTN 5' PAM, 24-base scan): database order is bin-major, not lexicographic, and blocks are always linear
    (DatabaseWriter.scala:84-85); hits must still come back in database order."""
    contigs = helpers.random_genome(77, 300_000, repeat_unit=70, n_repeats=150)
    fa = str(tmp_path / "cpf1.fa")
    helpers.write_fasta(fa, contigs)
    dbp = str(tmp_path / "cpf1_db")
    stats = oracle.build_database(fa, dbp, "cpf1")
    db = oracle.read_database(dbp)
    assert stats["indexed_bins"] == 0 and stats["targets"] > 3000
    targets = db.soa()[0]
    rng = np.random.default_rng(9)
    planted = []
    for _ in range(150):
        t = int(targets[int(rng.integers(0, len(targets)))]) & 0xFFFFFFFFFFFF
        for _ in range(int(rng.integers(0, 5))):
            t ^= int(rng.integers(1, 4)) << (2 * int(rng.integers(0, 20)))
        planted.append(t | (1 << 48))
    guides = np.asarray(planted, np.uint64)
    with ff.Context(0) as ctx:
        ctx.load_database(dbp)
        assert ctx.info().five_prime_pam == 1 and ctx.info().scan_len == 24
        for k in (0, 2, 4):
            ref = oracle.discover_blocks(db, guides, k, 2000)
            got = ctx.discover(guides, k, 2000, positions=True)
            helpers.assert_hits_equal(got, ref, check_positions=True)
        assert int(ref.row_ptr[-1]) >= 100
        with pytest.raises(ff.FlashFryError) as e:   # CFD / Hsu2013 are Cas9-23 only -> host prints "NA"
            ctx.score(guides[:1], [0, 0], [])
        assert e.value.code == -7


def test_large_k_and_many_guides_property(ff, oracle):
    """Size-independent properties on a larger synthetic index: every reported hit really is within k mismatches, rows are
    in strictly increasing database order, totals equal the summed counts, and the k=3 hit set is the subset of the
    k=5 hit set with mm <= 3 (before the overflow cut bites)."""
    with ff.Context(0) as ctx:
        ctx.synth_database(3, 5_000_000, 77)
        t = ctx.copy_targets()
        pack = oracle.PACK_BY_INDEX[3]
        guides = np.concatenate([helpers.random_guides(oracle, pack, 55, 3000), helpers.planted_guides(pack, t, 56, 1000)])
        h5 = ctx.discover(guides, 5, 1_000_000)
        h3 = ctx.discover(guides, 3, 1_000_000)
        seq = t & np.uint64(0xFFFFFFFFFFFF)
        mask = np.uint64(0x3FFFFFFFFFC0)
        for h, k in ((h5, 5), (h3, 3)):
            g_of_hit = np.repeat(np.arange(len(guides)), np.diff(h.row_ptr))
            x = (h.targets ^ guides[g_of_hit]) & mask
            y = (x | (x >> np.uint64(1))) & np.uint64(0x555555555555)
            mm = np.asarray([bin(int(v)).count("1") for v in y[:20000]])
            assert (mm == h.mismatches[:20000]).all() and int(h.mismatches.max()) <= k
            idx = np.searchsorted(seq, h.targets & np.uint64(0xFFFFFFFFFFFF))
            assert (t[idx] == h.targets).all()
            same_row = g_of_hit[1:] == g_of_hit[:-1]
            assert (np.diff(idx)[same_row] > 0).all()
            tot = np.add.reduceat((h.targets >> np.uint64(48)).astype(np.int64), h.row_ptr[:-1][np.diff(h.row_ptr) > 0])
            assert (tot == h.total_count[np.diff(h.row_ptr) > 0]).all()
        keep = h5.mismatches <= 3
        assert (h5.targets[keep] == h3.targets).all()


def test_dense_neighbourhood_regrows_hit_buffer_and_streams_long_buckets(ff, oracle):
    """Adversarial repeat-like index: every protospacer within two substitutions of one centre (x 4 N bases).  Guides
    near the centre hit thousands of targets each, so the candidate-hit buffer overflows and the scan is repeated with a
    larger one, part-one buckets exceed one 128-entry chunk, the per-warp staging spills to the direct path, and every
    guide overflows maximumOffTargets -- the kept rows must still be the reference's database-order prefix."""
    rng = np.random.default_rng(2024)
    centre = [int(x) for x in rng.integers(0, 4, 20)]

    def enc(bases20, n):
        v = 0
        for b in bases20:
            v = v * 4 + b
        return (v << 6) | (n << 4) | 0xA

    protos = {tuple(centre)}
    for i in range(20):
        for x in range(1, 4):
            a = list(centre); a[i] ^= x; protos.add(tuple(a))
            for j in range(i + 1, 20):
                for y in range(1, 4):
                    b2 = list(a); b2[j] ^= y; protos.add(tuple(b2))
    seqs = sorted(enc(p, n) for p in protos for n in range(4))
    counts = rng.integers(1, 40, len(seqs))
    targets = np.asarray(seqs, np.uint64) | (counts.astype(np.uint64) << np.uint64(48))
    assert len(targets) == 4 * 1771
    pack = oracle.PACK_BY_INDEX[3]
    guides = helpers.planted_guides(pack, targets, 9, 1500, max_subs=2)
    bin_off = oracle.bin_offsets_from_sorted(pack, 7, targets)
    with ff.Context(0) as ctx:
        ctx.load_database_arrays(3, targets)
        for max_ot in (2000, 10 ** 9):
            ref = oracle.discover_soa(pack, 7, targets, bin_off, guides, 4, max_ot, n_threads=os.cpu_count() or 1)
            got = ctx.discover(guides, 4, max_ot)
            helpers.assert_hits_equal(got, ref)
            assert got.n_candidate_hits > (1 << 22)          # more than the initial hit buffer
        assert ctx.timings().scan_launches >= 1
        assert ref.row_ptr[-1] > 5_000_000


def test_randomised_small_configurations(ff, oracle):
    """Fuzz: random spCas9-family indexes of 1..20 000 targets (clustered so that hits exist), random guides, k in 0..7
    (and one k larger than the protospacer), random maximumOffTargets; every result must equal the oracle's."""
    rng = np.random.default_rng(20261017)
    for trial in range(40):
        enzyme = int(rng.choice([2, 3, 3, 4, 5, 6]))
        pack = oracle.PACK_BY_INDEX[enzyme]
        proto = pack.scan_len - pack.pam_len
        n = int(rng.choice([1, 2, 17, 300, 5000, 20000]))
        centres = rng.integers(0, 1 << (2 * proto), max(1, n // 50), dtype=np.uint64)
        base = centres[rng.integers(0, len(centres), n)]
        for _ in range(3):  # up to three random substitutions around a centre
            pos = rng.integers(0, proto, n).astype(np.uint64)
            sub = rng.integers(0, 4, n).astype(np.uint64)
            base = base ^ (sub << (np.uint64(2) * pos))
        nbase = rng.integers(0, 4, n).astype(np.uint64)
        pam2 = {2: [0xA, 0x2], 3: [0xA], 4: [0x2], 5: [0xA, 0x2], 6: [0xA]}[enzyme]
        seq = (base << np.uint64(6)) | (nbase << np.uint64(4)) | np.asarray(rng.choice(pam2, n), np.uint64)
        seq = np.unique(seq)
        counts = np.where(rng.random(len(seq)) < 0.8, 1, rng.integers(1, 3000, len(seq))).astype(np.uint64)
        targets = seq | (counts << np.uint64(48))
        g = int(rng.choice([1, 3, 40, 400]))
        gp = centres[rng.integers(0, len(centres), g)]
        for _ in range(int(rng.integers(0, 4))):
            gp = gp ^ (rng.integers(0, 4, g).astype(np.uint64) << (np.uint64(2) * rng.integers(0, proto, g).astype(np.uint64)))
        guides = (gp << np.uint64(6)) | np.uint64(0x2A) | (np.uint64(1) << np.uint64(48))
        k = int(rng.choice([0, 1, 2, 3, 4, 4, 5, 6, 7, 25]))
        max_ot = int(rng.choice([1, 10, 2000, 2000, 10 ** 6]))
        bin_off = oracle.bin_offsets_from_sorted(pack, 7, targets)
        ref = oracle.discover_soa(pack, 7, targets, bin_off, guides, k, max_ot, n_threads=4)
        with ff.Context(0) as ctx:
            ctx.load_database_arrays(enzyme, targets)
            got = ctx.discover(guides, k, max_ot)
        try:
            helpers.assert_hits_equal(got, ref)
        except AssertionError as e:
            raise AssertionError("trial %d enzyme %d n %d g %d k %d maxOT %d: %s" % (trial, enzyme, len(targets), g, k, max_ot, e))


def test_pipelined_sub_batches_equal_single_batch(ff, oracle, small_db):
    """ff_discover cuts large guide sets into sub-batches whose D2H overlaps the next scan; the stitched CSR (row
    pointers, per-guide totals, fused scores) must equal the single-batch result and the oracle's."""
    _, db, _ = small_db
    targets = db.soa()[0]
    guides = np.concatenate([helpers.planted_guides(db.pack, targets, 41, 333, max_subs=3), helpers.random_guides(oracle, db.pack, 42, 70)])
    ref = oracle.discover_blocks(db, guides, 4, 2000)
    with ff.Context(0) as ctx:
        ctx.load_database(small_db[0])
        ctx.set_option("subbatch_min", 1000000)
        one, c1, s1, h1 = ctx.discover_score(guides, 4, 2000)
        for min_batch in (200, 100, 7):  # 2 sub-batches (60 / 40 %), 3 sub-batches (65 / 25 / 10 %)
            ctx.set_option("subbatch_min", min_batch)
            many, c4, s4, h4 = ctx.discover_score(guides, 4, 2000)
            helpers.assert_hits_equal(many, ref)
            helpers.assert_hits_equal(many, one)
            assert (c1 == c4).all() and (s1 == s4).all() and (h1 == h4).all()
            plain = ctx.discover(guides, 4, 2000)
            helpers.assert_hits_equal(plain, ref)
        withpos = ctx.discover(guides, 4, 2000, positions=True)   # positions force a single batch
        helpers.assert_hits_equal(withpos, ref, check_positions=True)


def test_hit_aggregates_minot_and_in_genome(ff, oracle):
    """ff_hit_aggregates (scoring/ClosestHit.scala:43-76 + the in-genome count of DangerousSequences.scala:61-65) against the
    oracle on fake.sites, plus the ClosestHitTest.scala:23-59 known answers ("1","70","0,70,20,0,10" ...)."""
    pack = oracle.pack_by_name("SPCAS9")
    guides = oracle.read_discover_tsv(os.path.join(GOLDEN, "fake.sites.gz"), pack, filter_overflow=False)
    enc = np.asarray([g.encoding for g in guides], np.uint64)
    row_ptr, targets = [0], []
    for g in guides:
        targets += g.targets
        row_ptr.append(len(targets))
    seq = "GACTTGCATCCGAAGCCGGTGGG"

    def mutate(n, count, salt):
        s = list(seq)
        for j in range(n):
            pos = (3 * j + salt) % 20
            s[pos] = "ACGT"[("ACGT".index(s[pos]) + 1 + salt % 3) % 4]
        return oracle.encode("".join(s), count)
    cases = [([(1, 1)], ("1", "1", "0,1,0,0,0")), ([(1, 40)], ("1", "40", "0,40,0,0,0")),
             ([(1, 40), (1, 30), (2, 20), (4, 10)], ("1", "70", "0,70,20,0,10")), ([], ("UNK", "0", "0,0,0,0,0")),
             ([(0, 7), (6, 3)], ("6", "3", "7,0,0,0,0"))]
    extra_g, extra_rp, extra_t = [], [0], []
    for ots, _ in cases:
        extra_g.append(oracle.encode(seq))
        extra_t += [mutate(n, c, i) for i, (n, c) in enumerate(ots)]
        extra_rp.append(len(extra_t))
    with ff.Context(0) as ctx:
        closest, cnt, hist, ing = ctx.hit_aggregates(pack.index, enc, row_ptr, targets)
        c2, n2, h2, i2 = ctx.hit_aggregates(pack.index, extra_g, extra_rp, extra_t)
    t = np.asarray(targets, np.uint64)
    for i, g in enumerate(guides):
        ots = t[row_ptr[i]:row_ptr[i + 1]]
        want = oracle.minot(pack, g.encoding, ots)
        got = ("UNK" if closest[i] == 2 ** 31 - 1 else str(closest[i]), str(cnt[i]), ",".join(str(x) for x in hist[i]))
        assert got == want
        dang = oracle.dangerous(pack, g.site.bases, g.encoding, ots)[2]
        assert dang == ("IN_GENOME=%d" % ing[i] if ing[i] > 0 else "NONE")
    for i, (_ots, want) in enumerate(cases):
        got = ("UNK" if c2[i] == 2 ** 31 - 1 else str(c2[i]), str(n2[i]), ",".join(str(x) for x in h2[i]))
        assert got == want, (i, got, want)
    assert i2[4] == 7


@pytest.mark.parametrize("k", [0, 1, 3, 4, 5, 6])
def test_bin_major_scan_equals_guide_major_and_oracle(small_ctx, small_db, oracle, k):
    """The bin-major kernels (k_bin_scan + k_pair_scan: bit-sliced compare, bins staged in shared memory, part-two pairs
    sorted by bucket) are selected for large batches on large indexes; forced here on the small FlashFry-format database."""
    _, db, _ = small_db
    targets = db.soa()[0]
    guides = np.concatenate([helpers.random_guides(oracle, db.pack, 17 + k, 333),
                             helpers.planted_guides(db.pack, targets, 71 + k, 400, max_subs=5)])
    ref = oracle.discover_blocks(db, guides, k, 2000)
    for pair_kernel in (1, 2):  # part two: lanes reading global memory / B-bins staged by the TMA ring
        with small_ctx.options(scan_kernel=2, pair_kernel=pair_kernel):
            got = small_ctx.discover(guides, k, 2000, positions=True)
            helpers.assert_hits_equal(got, ref, check_positions=True)
    with small_ctx.options(scan_kernel=2, pair_kernel=2, pair_segs=8):
        helpers.assert_hits_equal(small_ctx.discover(guides, k, 2000), ref)
    with small_ctx.options(scan_kernel=1):
        helpers.assert_hits_equal(small_ctx.discover(guides, k, 2000, positions=True), ref, check_positions=True)


def test_bin_major_scan_other_enzymes_and_edge_cases(ff, oracle, tmp_path):
    contigs = helpers.random_genome(78, 250_000, repeat_unit=70, n_repeats=150)
    fa = str(tmp_path / "g.fa")
    helpers.write_fasta(fa, contigs)
    for enzyme in ("cpf1", "spcas9ngg19", "spcas9"):
        dbp = str(tmp_path / ("db_" + enzyme))
        oracle.build_database(fa, dbp, enzyme)
        db = oracle.read_database(dbp)
        targets = db.soa()[0]
        guides = helpers.planted_guides(db.pack, targets, 5, 300, max_subs=4)
        with ff.Context(0) as ctx:
            ctx.set_option("scan_kernel", 2)
            ctx.load_database(dbp)
            for k, max_ot in ((4, 2000), (3, 2), (0, 2000)):
                ref = oracle.discover_blocks(db, guides, k, max_ot)
                for pair_kernel in (1, 2):
                    ctx.set_option("pair_kernel", pair_kernel)
                    helpers.assert_hits_equal(ctx.discover(guides, k, max_ot, positions=True), ref, check_positions=True)
            one = ctx.discover(guides[:1], 4, 2000)
            helpers.assert_hits_equal(one, oracle.discover_blocks(db, guides[:1], 4, 2000))
            assert ctx.discover(guides[:0], 4, 2000).n_guides == 0


def test_bin_major_scan_skewed_batches(ff, oracle):
    """Guide batches that break the bin-major kernel's balance assumptions: thousands of copies of one guide (one guide
    class, one hot bucket per seed, more visits than the shared-memory list holds, hit buffer regrown and the scan
    repeated), all guides in two classes, one guide."""
    pack = oracle.PACK_BY_INDEX[3]
    targets, seeds = helpers.family_database(oracle, seed=19, n_seeds=24, variants_per_seed=900)
    bin_off = oracle.bin_offsets_from_sorted(pack, 7, targets)
    with ff.Context(0) as ctx:
        ctx.set_option("scan_kernel", 2)
        ctx.load_database_arrays(3, targets)
        same = np.repeat(seeds[:1], 3000)
        two = np.concatenate([np.repeat(seeds[1:2], 500), helpers.planted_guides(pack, targets[:2000], 3, 500, max_subs=2)])
        for guides, max_ot in ((same, 2000), (same, 10 ** 6), (two, 50), (seeds[:1], 2000)):
            ref = oracle.discover_soa(pack, 7, targets, bin_off, guides, 4, max_ot)
            for pair_kernel in (1, 2):
                ctx.set_option("pair_kernel", pair_kernel)
                got = ctx.discover(guides, 4, max_ot)
                helpers.assert_hits_equal(got, ref)
        assert int(ref.row_ptr[-1]) > 0


def test_database_image_round_trip(ff, oracle, small_db, tmp_path):
    """ff_save_image / ff_load_image: the side-car brings back the same resident database (targets, positions, contigs),
    so discover gives the same rows and positions; damaged images are refused."""
    dbp, db, _ = small_db
    targets = db.soa()[0]
    guides = helpers.planted_guides(db.pack, targets, 12, 200, max_subs=4)
    img = str(tmp_path / "small.ffimg")
    with ff.Context(0) as ctx:
        ctx.load_database(dbp)
        ref = ctx.discover(guides, 4, 2000, positions=True)
        ctx.save_image(img)
    with ff.Context(0) as ctx:
        ctx.load_image(img)
        info = ctx.info()
        assert info.enzyme_index == 3 and info.n_targets == len(targets) and ctx.contigs() == db.contigs
        assert (ctx.copy_targets() == targets).all()
        helpers.assert_hits_equal(ctx.discover(guides, 4, 2000, positions=True), ref, check_positions=True)
        raw = open(img, "rb").read()
        bad = str(tmp_path / "bad.ffimg")
        open(bad, "wb").write(b"XX" + raw[2:])
        with pytest.raises(ff.FlashFryError) as e:
            ctx.load_image(bad)
        assert e.value.code == -5
        open(bad, "wb").write(raw[:len(raw) // 2])
        with pytest.raises(ff.FlashFryError) as e:
            ctx.load_image(bad)
        assert e.value.code == -5
        swapped = bytearray(raw)
        off = 64 + ((sum(len(c) + 1 for c in db.contigs) + 7) // 8) * 8
        swapped[off:off + 8], swapped[off + 8:off + 16] = raw[off + 8:off + 16], raw[off:off + 8]  # break the target order
        open(bad, "wb").write(bytes(swapped))
        with pytest.raises(ff.FlashFryError) as e:
            ctx.load_image(bad)
        assert e.value.code == -5


def test_bin_major_scan_oversized_bins_and_buckets(ff, oracle):
    """k_bin_scan stages a bin's slice of the bit-sliced index in shared memory (736 groups = 23 552 entries).  A bin
    that is larger is processed in runs of buckets, a single bucket that is larger is streamed from global memory:
    30 000 targets sharing their first 11 bases, 40 000 sharing their first 7, on a random background."""
    pack = oracle.PACK_BY_INDEX[3]
    rng = np.random.default_rng(77)

    def block(n_fixed, n):  # n distinct targets whose first n_fixed protospacer bases are one random value
        fixed = int(rng.integers(0, 1 << (2 * n_fixed)))
        free = 2 * (21 - n_fixed)  # remaining protospacer bases + N
        tail = np.unique(rng.integers(0, 1 << free, size=int(n * 1.3), dtype=np.uint64))[:n]
        return ((np.uint64(fixed) << np.uint64(free)) | tail) << np.uint64(4) | np.uint64(0xA)

    def tail_block(n):  # n distinct targets sharing their LAST 9 protospacer bases: one long bucket of index B
        fixed = np.uint64(int(rng.integers(0, 1 << 18)))
        head = np.unique(rng.integers(0, 1 << 22, size=int(n * 1.3), dtype=np.uint64))[:n]
        proto = (head << np.uint64(18)) | fixed
        return ((proto << np.uint64(2)) | rng.integers(0, 4, len(head)).astype(np.uint64)) << np.uint64(4) | np.uint64(0xA)

    seqs = np.unique(np.concatenate([block(11, 30_000), block(7, 40_000), block(0, 50_000), tail_block(20_000)]))
    counts = rng.integers(1, 4, len(seqs)).astype(np.uint64)
    targets = seqs | (counts << np.uint64(48))
    bin_off = oracle.bin_offsets_from_sorted(pack, 7, targets)
    guides = np.concatenate([helpers.planted_guides(pack, targets, 5, 160, max_subs=4), helpers.random_guides(oracle, pack, 6, 40)])
    with ff.Context(0) as ctx:
        ctx.load_database_arrays(3, targets)
        for max_ot in (2000, 10 ** 7):
            ref = oracle.discover_soa(pack, 7, targets, bin_off, guides, 4, max_ot, n_threads=os.cpu_count() or 1)
            for pair_kernel in (1, 2):  # (2: a B-bin whose run of groups exceeds the ring's buffer streams from global memory)
                with ctx.options(scan_kernel=2, pair_kernel=pair_kernel):
                    helpers.assert_hits_equal(ctx.discover(guides, 4, max_ot), ref)
            with ctx.options(scan_kernel=1):
                helpers.assert_hits_equal(ctx.discover(guides, 4, max_ot), ref)
        assert int(ref.row_ptr[-1]) > 10_000  # the planted guides really sit in the big bucket / bin


def test_compact_hit_lists_resolve_to_the_same_rows(small_ctx, small_db, oracle):
    """Option compact_hits: ff_discover ships 32-bit database indices instead of target longs (5 instead of 9 bytes per hit
    over PCIe); ff_hits_resolve rebuilds the longs from the host mirror of the database.  Also across sub-batches."""
    _, db, _ = small_db
    targets = db.soa()[0]
    guides = np.concatenate([helpers.random_guides(oracle, db.pack, 91, 200), helpers.planted_guides(db.pack, targets, 92, 300, max_subs=4)])
    ref = oracle.discover_blocks(db, guides, 4, 2000)
    for min_batch in (1_000_000, 100):
        with small_ctx.options(compact_hits=1, subbatch_min=min_batch):
            raw = small_ctx.discover(guides, 4, 2000)
            full = small_ctx.discover(guides, 4, 2000, resolve=True)
        assert raw.target_index is not None and len(raw.targets) == 0
        assert (targets[raw.target_index] == ref.targets).all()
        helpers.assert_hits_equal(full, ref)
    helpers.assert_hits_equal(small_ctx.discover(guides, 4, 2000), ref)  # option off again: longs as before


def test_score_rejects_malformed_hit_lists(ff, oracle):
    """ff_score / ff_hit_aggregates read row_ptr[n_guides] targets from a caller-built CSR: a row_ptr that does not start at 0,
    is not monotone or does not end at n_hits is refused (FF_EINVAL), as is an enzyme the scorers are not valid for."""
    g = [oracle.encode("GAGTCCGAGCAGAAGAAGAAGGG")]
    t = np.asarray([oracle.encode("GAGTCCGAGCAGAAGAAGAATGG")], np.uint64)
    with ff.Context(0) as ctx:
        ctx.score(g, [0, 1], t)
        for bad in ([1, 1], [0, 2], [0, 0]):
            with pytest.raises(ff.FlashFryError) as e:
                ctx.score(g, bad, t)
            assert e.value.code == -1
        with pytest.raises(ff.FlashFryError):
            ctx.score(g + g, [0, 1, 0], t)
