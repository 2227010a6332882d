"""ff_multi: several GPUs behind one host process -- guide-sharded discover through the C ABI with one NCCL all-gather of
the per-guide totals.  The one-device case runs everywhere; the sharded cases need >= 2 GPUs."""
import os

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("n_dev", [1, 2, 4])
def test_multi_discover_equals_single_context_and_oracle(ff, oracle, n_dev):
    if _n_gpus() < n_dev:
        pytest.skip("needs %d GPUs" % n_dev)
    pack = oracle.PACK_BY_INDEX[3]
    with ff.MultiContext(list(range(n_dev))) as mc, ff.Context(0) as one:
        mc.synth_database(3, 300_000, 7)
        one.synth_database(3, 300_000, 7)
        targets = one.copy_targets()
        guides = np.concatenate([helpers.planted_guides(pack, targets, 3, 1001, max_subs=4), helpers.random_guides(oracle, pack, 4, 200)])
        ref = oracle.discover_soa(pack, 7, targets, oracle.bin_offsets_from_sorted(pack, 7, targets), guides, 4, 50)
        single = one.discover(guides, 4, 50)
        helpers.assert_hits_equal(single, ref)
        for scan_kernel in (0, 2):
            mc.set_option("scan_kernel", scan_kernel)
            shards, totals = mc.discover(guides, 4, 50)
            assert len(shards) == n_dev and (totals == ref.total_count).all()
            for r, h in enumerate(shards):
                first, count = ff.shard_range(len(guides), n_dev, r)
                assert h.n_guides == count
                lo, hi = int(ref.row_ptr[first]), int(ref.row_ptr[first + count])
                assert (h.row_ptr == ref.row_ptr[first:first + count + 1] - lo).all()
                assert (h.targets == ref.targets[lo:hi]).all() and (h.mismatches == ref.mismatches[lo:hi]).all()
                assert (h.overflowed == ref.overflowed[first:first + count]).all()
        # fewer guides than devices, and none at all
        shards, totals = mc.discover(guides[:1], 4, 50)
        assert sum(h.n_guides for h in shards) == 1 and (totals == ref.total_count[:1]).all()
        shards, totals = mc.discover(guides[:0], 4, 50)
        assert sum(h.n_guides for h in shards) == 0 and len(totals) == 0


def test_multi_load_database_broadcasts_the_decoded_arrays(ff, oracle, small_db):
    """ff_multi_load_database reads and inflates FlashFry's files once; the other devices receive the decoded arrays with
    ncclBroadcast and build their own index: every shard must answer like a context that loaded the files itself."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    path, db, _ = small_db
    targets = db.soa()[0]
    guides = helpers.planted_guides(db.pack, targets, 12, 400, max_subs=3)
    ref = oracle.discover_blocks(db, guides, 4, 2000)
    with ff.MultiContext([0, 1]) as mc:
        mc.load_database(path)
        shards, totals = mc.discover(guides, 4, 2000, positions=True)
        assert (totals == ref.total_count).all()
        got_t = np.concatenate([h.targets for h in shards])
        got_p = np.concatenate([h.positions for h in shards])
        assert (got_t == ref.targets).all() and len(got_p) == len(ref.positions)


def _assert_shards_equal(ff, shards, totals, ref, n_guides, n_ranks):
    assert len(shards) == n_ranks and (totals == ref.total_count[:n_guides]).all()
    for r, h in enumerate(shards):
        first, count = ff.shard_range(n_guides, n_ranks, r)
        assert h.n_guides == count
        lo, hi = int(ref.row_ptr[first]), int(ref.row_ptr[first + count])
        assert (h.row_ptr == ref.row_ptr[first:first + count + 1] - lo).all()
        assert (h.targets == ref.targets[lo:hi]).all() and (h.mismatches == ref.mismatches[lo:hi]).all()
        assert (h.overflowed == ref.overflowed[first:first + count]).all()
        assert (h.total_count == ref.total_count[first:first + count]).all()


@pytest.mark.parametrize("devices", [[0, 0], [0, 0, 0], [0, 1], [0, 1, 2, 3]])
def test_database_sharded_discover_over_peer_memory(ff, oracle, devices):
    """shard_mode = 1 (ff_shard.inl): every rank scans 1/n of the INDEX for all guides; the scan kernels push each candidate
    into the exchange block of the guide's owner (peer stores into one region per source rank), barriers and the all-gather
    of the totals go through the same blocks.  Rows must equal the oracle's.  Several ranks on ONE device exercise the same kernels
    (the "peers" are then blocks in the same HBM), so the path is covered on a single-GPU box too."""
    if _n_gpus() <= max(devices):
        pytest.skip("needs %d GPUs" % (max(devices) + 1))
    pack = oracle.PACK_BY_INDEX[3]
    n = len(devices)
    with ff.MultiContext(devices) as mc, ff.Context(0) as one:
        mc.synth_database(3, 300_000, 7)
        one.synth_database(3, 300_000, 7)
        targets = one.copy_targets()
        guides = np.concatenate([helpers.planted_guides(pack, targets, 3, 1001, max_subs=4), helpers.random_guides(oracle, pack, 4, 200)])
        bin_off = oracle.bin_offsets_from_sorted(pack, 7, targets)
        mc.set_option("shard_mode", 1)
        for k, max_ot in ((4, 50), (4, 2000), (3, 2000), (0, 2000)):
            ref = oracle.discover_soa(pack, 7, targets, bin_off, guides, k, max_ot)
            for pair_kernel in (1, 2):
                mc.set_option("pair_kernel", pair_kernel)
                shards, totals = mc.discover(guides, k, max_ot)
                _assert_shards_equal(ff, shards, totals, ref, len(guides), n)
            assert mc.rank_timings(0).kernel_launches > 10  # the sharded path ran (not a fallback)
        ref = oracle.discover_soa(pack, 7, targets, bin_off, guides, 4, 2000)
        for m in (1, 2, 0):  # fewer guides than ranks, none at all; and again the full set (blocks are reused)
            shards, totals = mc.discover(guides[:m], 4, 2000)
            _assert_shards_equal(ff, shards, totals, ref, m, n) if m else None
            assert sum(h.n_guides for h in shards) == m
        shards, totals = mc.discover(guides, 4, 2000)
        _assert_shards_equal(ff, shards, totals, ref, len(guides), n)
        # an exchange block that is too small is an error, not a truncated row
        mc.set_option("peer_hit_cap", 64)
        with pytest.raises(ff.FlashFryError):
            mc.discover(guides, 4, 2000)
