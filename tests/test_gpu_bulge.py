"""GPU tests of the general discover path: the 1-bp bulge EXTENSION and database-order windows.

The bulge mode does not exist in the reference, so these are NOT reference-parity tests: the CUDA
path is compared with the brute-force definition in oracle/ff_oracle.c (ffo_discover_bulge), which tests/
test_oracle_pins.py cross-checks against an independent string-surgery restatement.  The windowed scan IS
reference behaviour (overflowed guides leave the traversal, OrderedBinTraversalFactory.scala:107-119) and must give
exactly the rows of the whole-database scan.
"""
import os

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ff():
    import flashfry_b200.api as api
    return api


@pytest.fixture(scope="module")
def family_db(oracle):
    """~60 k targets built from 48 seed protospacers with substitutions and single-base indels, so that guides have many
    mismatch-only, RNA-bulge and DNA-bulge neighbours; occurrence counts are mixed (a few above 1000)."""
    return helpers.family_database(oracle, seed=11, n_seeds=48, variants_per_seed=1400)


@pytest.fixture(scope="module")
def family_ctx(ff, family_db):
    ctx = ff.Context(0)
    ctx.load_database_arrays(3, family_db[0])
    yield ctx
    ctx.close()


def assert_bulge_equal(got, ref):
    helpers.assert_hits_equal(got, ref)
    assert got.bulge is not None and (np.asarray(got.bulge) == np.asarray(ref.bulge)).all(), "bulge codes differ"


@pytest.mark.parametrize("flags", [1, 2, 3])
@pytest.mark.parametrize("k", [0, 2, 4])
def test_bulge_matches_definition(family_ctx, family_db, oracle, flags, k):
    targets, seeds = family_db
    pack = oracle.PACK_BY_INDEX[3]
    guides = np.concatenate([seeds[:16], helpers.planted_guides(pack, targets, 100 + k, 16, max_subs=3),
                             helpers.random_guides(oracle, pack, 5, 4)])
    ref = oracle.discover_bulge(pack, targets, guides, k, 10 ** 6, flags, n_threads=os.cpu_count() or 1)
    got = family_ctx.discover_bulge(guides, k, 10 ** 6, flags)
    assert_bulge_equal(got, ref)
    assert not np.asarray(ref.overflowed).any()
    if k >= 2:
        b = np.asarray(ref.bulge)
        assert int((b != 0).sum()) > 50, "the inputs must exercise bulged alignments"
        if flags & 1:
            assert ((b & 0xC0) == 0x40).any()
        if flags & 2:
            assert ((b & 0xC0) == 0x80).any()


@pytest.mark.parametrize("max_ot", [0, 1, 7, 300])
def test_bulge_overflow_cut(family_ctx, family_db, oracle, max_ot):
    targets, seeds = family_db
    pack = oracle.PACK_BY_INDEX[3]
    guides = seeds[16:40]
    ref = oracle.discover_bulge(pack, targets, guides, 3, max_ot, 3, n_threads=os.cpu_count() or 1)
    got = family_ctx.discover_bulge(guides, 3, max_ot, 3)
    assert_bulge_equal(got, ref)
    assert np.asarray(ref.overflowed).any()


@pytest.mark.parametrize("cells", [1, 5, 16])
def test_bulge_windows_equal_whole_database_scan(family_ctx, family_db, oracle, cells):
    """Database-order windows + dropping full guides give exactly the rows of the one-window scan."""
    targets, seeds = family_db
    pack = oracle.PACK_BY_INDEX[3]
    guides = np.concatenate([seeds[:24], helpers.planted_guides(pack, targets, 9, 24, max_subs=2)])
    with family_ctx.options(window_cells=64):
        whole = family_ctx.discover_bulge(guides, 3, 150, 3)
    with family_ctx.options(window_cells=cells):
        win = family_ctx.discover_bulge(guides, 3, 150, 3)
    assert_bulge_equal(win, whole)
    ref = oracle.discover_bulge(pack, targets, guides, 3, 150, 3, n_threads=os.cpu_count() or 1)
    assert_bulge_equal(win, ref)


@pytest.mark.parametrize("cells", [1, 7, 64])
@pytest.mark.parametrize("max_ot", [5, 2000])
def test_windowed_mismatch_search_equals_plain(family_ctx, family_db, oracle, cells, max_ot):
    """The reference's own (mismatch-only) search through the general path -- windows, active-guide compaction --
    equals the plain kernel's rows and the oracle's restatement of the reference loops."""
    targets, seeds = family_db
    pack = oracle.PACK_BY_INDEX[3]
    guides = np.concatenate([seeds, helpers.planted_guides(pack, targets, 3, 64, max_subs=4), helpers.random_guides(oracle, pack, 8, 32)])
    plain = family_ctx.discover(guides, 4, max_ot)
    with family_ctx.options(force_general=1, window_cells=cells):
        gen = family_ctx.discover(guides, 4, max_ot)
    helpers.assert_hits_equal(gen, plain)
    ref = oracle.discover_soa(pack, 7, targets, oracle.bin_offsets_from_sorted(pack, 7, targets), guides, 4, max_ot)
    helpers.assert_hits_equal(gen, ref)


def test_bulge_flags_zero_is_plain_discover(family_ctx, family_db, oracle):
    targets, seeds = family_db
    plain = family_ctx.discover(seeds, 4, 2000)
    got = family_ctx.discover_bulge(seeds, 4, 2000, 0)
    helpers.assert_hits_equal(got, plain)
    assert got.bulge is not None and not np.asarray(got.bulge).any()


def test_bulge_with_positions(ff, oracle, tmp_path):
    """Positions of bulge hits are the positions of the stored targets (gathered by database index)."""
    contigs = helpers.random_genome(77, 120_000, repeat_unit=50, n_repeats=200)
    fa = str(tmp_path / "g.fa")
    helpers.write_fasta(fa, contigs)
    dbp = str(tmp_path / "db")
    oracle.build_database(fa, dbp, "spcas9ngg")
    db = oracle.read_database(dbp)
    targets, _, pos_off, positions = db.soa()
    guides = helpers.planted_guides(db.pack, targets, 4, 40, max_subs=2)
    with ff.Context(0) as ctx:
        ctx.load_database(dbp)
        got = ctx.discover_bulge(guides, 3, 2000, 3, positions=True)
    ref = oracle.discover_bulge(db.pack, targets, guides, 3, 2000, 3)
    assert_bulge_equal(got, ref)
    key = targets & np.uint64(0xFFFFFFFFFFFF)
    for h in range(len(got.targets)):
        t = int(np.searchsorted(key, got.targets[h] & np.uint64(0xFFFFFFFFFFFF)))
        want = positions[int(pos_off[t]):int(pos_off[t + 1])]
        assert (got.positions[int(got.pos_ptr[h]):int(got.pos_ptr[h + 1])] == want).all()


def test_bulge_unsupported_enzyme(ff, oracle):
    pack = oracle.PACK_BY_INDEX[1]  # Cpf1: 5' PAM, 24-mer
    t = np.sort(np.random.default_rng(1).integers(0, 1 << 48, 1000, dtype=np.uint64)) | (np.uint64(1) << np.uint64(48))
    with ff.Context(0) as ctx:
        ctx.load_database_arrays(5, np.unique(t & np.uint64((1 << 44) - 1)) | (np.uint64(1) << np.uint64(48)))
        with pytest.raises(ff.FlashFryError) as e:
            ctx.discover_bulge(t[:4], 3, 2000, 3)
        assert e.value.code == -7


def test_bulge_properties_on_a_large_synthetic_index(ff, oracle):
    """5e6-target synthetic index, k = 5 + both bulges (the shape of BASELINE configs[3], scaled down): every reported
    hit re-verifies against the definition, rows are strictly increasing in database order, totals follow the overflow
    rule, and four guides are checked against the brute force completely."""
    pack = oracle.PACK_BY_INDEX[3]
    with ff.Context(0) as ctx:
        ctx.synth_database(3, 5_000_000, 42)
        targets = ctx.copy_targets()
        guides = np.concatenate([helpers.random_guides(oracle, pack, 21, 200), helpers.planted_guides(pack, targets, 22, 56, max_subs=3)])
        got = ctx.discover_bulge(guides, 5, 300, 3)
        with ctx.options(window_cells=64):
            whole = ctx.discover_bulge(guides, 5, 300, 3)
    assert_bulge_equal(got, whole)
    assert np.asarray(got.overflowed).mean() > 0.5  # ~2000 expected hits per guide: windows and early exit are exercised
    seq = np.uint64(0xFFFFFFFFFFFF)
    for g in range(len(guides)):
        lo, hi = int(got.row_ptr[g]), int(got.row_ptr[g + 1])
        t = got.targets[lo:hi]
        assert (np.diff((t & seq).astype(np.int64)) > 0).all()
        cnt = (t >> np.uint64(48)).astype(np.int64)
        assert int(cnt.sum()) == int(got.total_count[g])
        assert bool(got.overflowed[g]) == (int(cnt.sum()) >= 300)
        if hi > lo:
            assert int(cnt[:-1].sum()) < 300  # the shortest prefix that reaches the limit
        for i in range(lo, min(hi, lo + 40)):
            mm, ty, q = oracle.bulge_align(int(guides[g]), int(got.targets[i]), 3)
            assert mm == int(got.mismatches[i]) <= 5
            assert int(got.bulge[i]) == (0 if ty == 0 else ((0x40 if ty == 1 else 0x80) | q))
    sub = np.asarray([0, 1, 200, 201])
    ref = oracle.discover_bulge(pack, targets, guides[sub], 5, 300, 3, n_threads=os.cpu_count() or 1)
    for j, g in enumerate(sub):
        lo, hi = int(got.row_ptr[g]), int(got.row_ptr[g + 1])
        rlo, rhi = int(ref.row_ptr[j]), int(ref.row_ptr[j + 1])
        assert (got.targets[lo:hi] == ref.targets[rlo:rhi]).all() and (got.bulge[lo:hi] == ref.bulge[rlo:rhi]).all()


def test_bulge_sub_batches_equal_single_batch(family_ctx, family_db, oracle):
    """The host entry point pipelines guide sub-batches (D2H of one behind the scan of the next) in bulge mode too."""
    targets, seeds = family_db
    pack = oracle.PACK_BY_INDEX[3]
    guides = np.concatenate([seeds, helpers.planted_guides(pack, targets, 31, 50, max_subs=3)])
    with family_ctx.options(subbatch_min=1000000):
        one = family_ctx.discover_bulge(guides, 3, 400, 3)
    for min_batch in (40, 20):
        with family_ctx.options(subbatch_min=min_batch):
            many = family_ctx.discover_bulge(guides, 3, 400, 3)
        assert_bulge_equal(many, one)


def test_cli_discover_with_bulge_option(ff, oracle, tmp_path):
    """`flashfry_b200_cli discover --bulge rna,dna` (an extension of FlashFry's command line): every row's tokens
    SEQ_count_mm[_R<q>|_D<q>] equal ff_discover_bulge's rows for the same guides."""
    import subprocess
    cli = os.path.join(os.path.dirname(os.path.abspath(ff.__file__)), "flashfry_b200_cli")
    contigs = helpers.random_genome(91, 150_000, repeat_unit=50, n_repeats=100)
    fa = str(tmp_path / "g.fa")
    helpers.write_fasta(fa, contigs)
    dbp = str(tmp_path / "db")
    oracle.build_database(fa, dbp, "spcas9ngg")
    db = oracle.read_database(dbp)
    targets = db.soa()[0]
    rng = np.random.default_rng(4)
    with open(str(tmp_path / "guides.fa"), "w") as fh:
        for i in range(12):
            seq, _ = oracle.decode(int(targets[int(rng.integers(0, len(targets)))]), 23)
            s = list(seq)
            for _ in range(int(rng.integers(0, 3))):
                s[int(rng.integers(0, 20))] = "ACGT"[int(rng.integers(0, 4))]
            fh.write(">g%d\n%s\n" % (i, "ATATATATAT" + "".join(s) + "ATATATATAT"))
    out = str(tmp_path / "out.tsv")
    r = subprocess.run([cli, "discover", "--fasta", str(tmp_path / "guides.fa"), "--database", dbp, "--output", out, "--maxMismatch", "3",
                        "--bulge", "rna,dna"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rows = [l.rstrip("\n").split("\t") for l in open(out)][1:]
    assert len(rows) >= 12
    guides = np.asarray([oracle.encode(row[3], 1) for row in rows], np.uint64)
    with ff.Context(0) as ctx:
        ctx.load_database(dbp)
        ref = ctx.discover_bulge(guides, 3, 2000, 3)
    n_bulged = 0
    for g, row in enumerate(rows):
        toks = [t for t in row[-1].split(",") if t]
        lo, hi = int(ref.row_ptr[g]), int(ref.row_ptr[g + 1])
        assert len(toks) == hi - lo and int(row[-2]) == int(ref.total_count[g])
        for t, i in zip(toks, range(lo, hi)):
            f = t.split("_")
            assert oracle.encode(f[0], int(f[1])) == int(ref.targets[i]) and int(f[2]) == int(ref.mismatches[i])
            code = int(ref.bulge[i])
            if code:
                n_bulged += 1
                assert f[3] == ("R" if (code & 0xC0) == 0x40 else "D") + str(code & 0x3F)
            else:
                assert len(f) == 3
    assert n_bulged > 0
