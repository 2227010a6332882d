"""Parity at the benchmark's own size (BASELINE.json configs[2]): a 3e8-target synthetic spCas9-NGG index resident in HBM,
the oracle on a 2 048-guide sample with all host threads, both scan kernels, the 100 000-guide batch of bench.py, scores;
and the same with sequence-level skew (repeat neighbourhoods).  ~1 minute on a B200 box."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import bench  # noqa: E402  (guide generator and seeds of the bench line)
import helpers  # noqa: E402

pytestmark = pytest.mark.gpu
N_TARGETS = 300_000_000


def _hits_equal(got, ref):
    helpers.assert_hits_equal(got, ref)


@pytest.fixture(scope="module")
def big(ff, oracle):
    ctx = ff.Context(0)
    ctx.synth_database(bench.ENZYME, N_TARGETS, bench.SEED_DB)
    targets = ctx.copy_targets()
    pack = oracle.PACK_BY_INDEX[bench.ENZYME]
    bin_off = oracle.bin_offsets_from_sorted(pack, 7, targets)
    rng = np.random.default_rng(17)
    pool = np.concatenate([targets[int(s):int(s) + 2048] for s in rng.integers(0, len(targets) - 2048, 16)])
    guides = bench.make_guides(100_000, bench.SEED_GUIDES, pool, bench.SEED_PLANTED)
    yield ctx, targets, bin_off, pack, guides
    ctx.close()


def test_benchmark_size_both_scan_kernels_equal_the_oracle(big, oracle):
    ctx, targets, bin_off, pack, guides = big
    assert len(targets) > 0.99 * N_TARGETS and ctx.info().seed_split_a == 11
    sample = guides[:2048]
    ref = oracle.discover_soa(pack, 7, targets, bin_off, sample, 4, 2000, n_threads=os.cpu_count() or 1)
    assert int(ref.row_ptr[-1]) > 100 * len(sample)  # ~116 hits per guide at k = 4 on a human-sized index
    for kernel in (1, 2):
        with ctx.options(scan_kernel=kernel):
            _hits_equal(ctx.discover(sample, 4, 2000), ref)
    with ctx.options(scan_kernel=2, pair_kernel=2):  # part two through the two-buffer TMA ring (default only beyond ~110 000 guides)
        _hits_equal(ctx.discover(sample, 4, 2000), ref)
    # the sample inside the full batch, through the host entry point with its sub-batches (bin-major kernels by default)
    full = ctx.discover(guides, 4, 2000)
    n = int(ref.row_ptr[-1])
    assert (full.row_ptr[:2049] == ref.row_ptr).all() and (full.targets[:n] == ref.targets).all() and (full.mismatches[:n] == ref.mismatches).all()
    assert (full.total_count[:2048] == ref.total_count).all() and (full.overflowed[:2048] == ref.overflowed).all()
    for opts in ({"scan_kernel": 1}, {"pair_kernel": 2}):
        with ctx.options(**opts):
            other = ctx.discover(guides, 4, 2000)
        assert (other.row_ptr == full.row_ptr).all() and (other.targets == full.targets).all() and (other.mismatches == full.mismatches).all()
    # compact hit lists at this size
    with ctx.options(compact_hits=1):
        comp = ctx.discover(guides[:20000], 4, 2000, resolve=True)
    m = int(full.row_ptr[20000])
    assert (comp.targets == full.targets[:m]).all() and (targets[comp.target_index] == comp.targets).all()


def test_benchmark_size_scores_equal_the_oracle(big, oracle):
    ctx, targets, bin_off, pack, guides = big
    sample = guides[5000:5064]
    ref = oracle.discover_soa(pack, 7, targets, bin_off, sample, 4, 2000, n_threads=os.cpu_count() or 1)
    hits, cmax, cspec, hsu = ctx.discover_score(sample, 4, 2000)
    _hits_equal(hits, ref)
    for g in range(len(sample)):
        ots = ref.targets[ref.row_ptr[g]:ref.row_ptr[g + 1]]
        mx, sp, _ = oracle.cfd_guide(int(sample[g]), ots)
        assert cmax[g] == mx and cspec[g] == sp and hsu[g] == oracle.hsu_guide(pack, int(sample[g]), ots)


@pytest.mark.parametrize("k", [3, 5])
def test_benchmark_size_other_mismatch_budgets(big, oracle, k):
    """k = 3 (plan: hA = 1 / hB = 1) and k = 5 (~1 100 candidates per guide: the radix-sort ordering) on a 256-guide sample."""
    ctx, targets, bin_off, pack, guides = big
    sample = guides[3000:3256]
    ref = oracle.discover_soa(pack, 7, targets, bin_off, sample, k, 2000, n_threads=os.cpu_count() or 1)
    for kernel in (1, 2):
        with ctx.options(scan_kernel=kernel):
            _hits_equal(ctx.discover(sample, k, 2000), ref)


def test_skewed_index_at_benchmark_size(ff, oracle):
    """Sequence-level skew (SURVEY 8(d) / VERDICT: Alu-like neighbourhoods): 64 families of 16 384 targets within 0..3
    substitutions of a consensus; guides planted in them collect thousands of candidates (segments beyond the in-register
    sort: k_sort_long; hot buckets in both index halves).  2 048-guide sample against the oracle, both kernels."""
    pack = oracle.PACK_BY_INDEX[bench.ENZYME]
    with ff.Context(0) as ctx:
        ctx.synth_database_skewed(bench.ENZYME, N_TARGETS, bench.SEED_DB, 64, 16384, 3)
        targets = ctx.copy_targets()
        bin_off = oracle.bin_offsets_from_sorted(pack, 7, targets)
        rng = np.random.default_rng(23)
        pool = np.concatenate([targets[int(s):int(s) + 2048] for s in rng.integers(0, len(targets) - 2048, 16)])
        guides = bench.make_guides(2048, 77, pool, 78)
        cons = bench.family_consensus(bench.SEED_DB, rng.integers(0, 64, 40))
        for _ in range(2):
            pos = rng.integers(0, 20, 40).astype(np.uint64)
            cons = cons ^ (rng.integers(0, 4, 40).astype(np.uint64) << (np.uint64(2) * (np.uint64(20) - pos)))
        guides[:40] = (cons << np.uint64(4)) | np.uint64(0xA) | (np.uint64(1) << np.uint64(48))
        for max_ot in (2000, 10 ** 7):
            ref = oracle.discover_soa(pack, 7, targets, bin_off, guides, 4, max_ot, n_threads=os.cpu_count() or 1)
            for kernel in (1, 2):
                with ctx.options(scan_kernel=kernel):
                    _hits_equal(ctx.discover(guides, 4, max_ot), ref)
        per_guide = np.diff(ref.row_ptr)
        assert int(per_guide[:40].max()) > 1000, "the planted guides must sit inside a repeat neighbourhood"
        assert int(ref.overflowed.sum()) == 0
