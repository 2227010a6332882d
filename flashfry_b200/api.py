"""Thin Python harness over the C ABI, used by the tests and bench.py.  The product is the C ABI + CUDA library;
this module adds no computation of its own (and never touches oracle/)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _native as N
from ._native import FF_BULGE_DNA, FF_BULGE_RNA, FF_METRIC_CFD, FF_METRIC_HSU2013, FlashFryError  # noqa: F401


@dataclass
class Hits:
    """Copy of an ff_hits (CSR over guides; see include/flashfry_b200.h)."""
    row_ptr: np.ndarray
    targets: np.ndarray
    mismatches: np.ndarray
    total_count: np.ndarray
    overflowed: np.ndarray
    pos_ptr: Optional[np.ndarray]
    positions: Optional[np.ndarray]
    n_compares: int
    n_candidate_hits: int
    bulge: Optional[np.ndarray] = None  # ff_discover_bulge only: 0 none, 0x40|q RNA bulge, 0x80|q DNA bulge
    target_index: Optional[np.ndarray] = None  # option compact_hits: database index of every hit

    @property
    def n_guides(self) -> int:
        return len(self.row_ptr) - 1

    def row(self, g: int):
        lo, hi = int(self.row_ptr[g]), int(self.row_ptr[g + 1])
        return self.targets[lo:hi], self.mismatches[lo:hi]


def _arr(ptr, n, dtype):
    if n <= 0 or not ptr:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


def _take_hits(hp) -> Hits:
    h = hp.contents
    G, H = int(h.n_guides), int(h.n_hits)
    row_ptr = _arr(h.row_ptr, G + 1, np.int64)
    pos_ptr = positions = None
    if h.pos_ptr:
        pos_ptr = _arr(h.pos_ptr, H + 1, np.int64)
        positions = _arr(h.positions, int(pos_ptr[-1]), np.uint64)
    out = Hits(row_ptr, _arr(h.targets, H, np.uint64), _arr(h.mismatches, H, np.uint8), _arr(h.total_count, G, np.int32),
               _arr(h.overflowed, G, np.uint8), pos_ptr, positions, int(h.n_compares), int(h.n_candidate_hits))
    if h.bulge:
        out.bulge = _arr(h.bulge, H, np.uint8)
    if h.target_index:
        out.target_index = _arr(h.target_index, H, np.uint32)
    N.lib().ff_hits_free(hp)
    return out


def _u64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint64))


class Context:
    """One GPU, one resident database (ff_ctx)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        N.check(N.lib().ff_create(C.byref(self._h), device))

    def close(self):
        if self._h:
            N.lib().ff_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- database
    def set_stream(self, cuda_stream: int):
        N.check(N.lib().ff_set_stream(self._h, C.c_void_p(cuda_stream)))

    def set_option(self, key: str, value: int):
        """ff_set_option: tuning / test knobs of this context (see include/flashfry_b200.h)."""
        N.check(N.lib().ff_set_option(self._h, key.encode(), int(value)))

    def options(self, **kw):
        """Context manager: set options for a block, restore the defaults afterwards."""
        ctx = self
        defaults = {"scan_kernel": 0, "force_general": 0, "window_cells": 0, "subbatch_min": 45000, "subbatch_c1": 65,
                    "subbatch_c2": 90, "group_sort": 1, "trace": 0, "b_spi": 0, "split_a": 0, "compact_hits": 0, "pair_kernel": 0,
                    "pair_segs": 0, "subbatch_two": 70, "peer_local_only": 0, "debug_bin_div": 0}

        class _O:
            def __enter__(self_o):
                for k, v in kw.items():
                    ctx.set_option(k, v)

            def __exit__(self_o, *a):
                for k in kw:
                    ctx.set_option(k, defaults[k])
        return _O()

    def load_database(self, db_path: str, header_path: Optional[str] = None):
        N.check(N.lib().ff_load_database(self._h, db_path.encode(), header_path.encode() if header_path else None))

    def save_image(self, path: str):
        N.check(N.lib().ff_save_image(self._h, path.encode()))

    def load_image(self, path: str):
        N.check(N.lib().ff_load_image(self._h, path.encode()))

    def load_database_arrays(self, enzyme_index: int, targets, positions=None, contigs: Sequence[str] = (), bin_width: int = 7):
        t = _u64(targets)
        p = _u64(positions) if positions is not None else None
        names = (C.c_char_p * max(len(contigs), 1))(*[c.encode() for c in contigs])
        N.check(N.lib().ff_load_database_arrays(
            self._h, enzyme_index, bin_width, t.ctypes.data_as(C.POINTER(C.c_uint64)), len(t),
            p.ctypes.data_as(C.POINTER(C.c_uint64)) if p is not None else None, len(p) if p is not None else 0,
            names, len(contigs)))

    def synth_database(self, enzyme_index: int, n_targets: int, seed: int):
        N.check(N.lib().ff_synth_database(self._h, enzyme_index, n_targets, seed))

    def synth_database_skewed(self, enzyme_index: int, n_targets: int, seed: int, n_families: int, family_size: int, family_subs: int = 3):
        N.check(N.lib().ff_synth_database_skewed(self._h, enzyme_index, n_targets, seed, n_families, family_size, family_subs))

    def info(self) -> N.FFDbInfo:
        i = N.FFDbInfo()
        N.check(N.lib().ff_db_info(self._h, C.byref(i)))
        return i

    def contigs(self) -> List[str]:
        return [N.lib().ff_db_contig(self._h, k + 1).decode() for k in range(self.info().n_contigs)]

    def copy_targets(self, first: int = 0, n: Optional[int] = None) -> np.ndarray:
        if n is None:
            n = int(self.info().n_targets) - first
        out = np.empty(n, np.uint64)
        N.check(N.lib().ff_db_copy_targets(self._h, first, n, out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return out

    # ---- discover / score
    def discover(self, guides, max_mismatch: int = 4, maximum_off_targets: int = 2000, positions: bool = False,
                 resolve: bool = False) -> Hits:
        """ff_discover.  With option compact_hits the hit list carries target_index instead of targets; resolve=True
        then calls ff_hits_resolve (targets filled from the host mirror of the database)."""
        g = _u64(guides)
        hp = C.POINTER(N.FFHits)()
        N.check(N.lib().ff_discover(self._h, g.ctypes.data_as(C.POINTER(C.c_uint64)), len(g), max_mismatch,
                                    maximum_off_targets, int(positions), C.byref(hp)))
        if resolve:
            N.check(N.lib().ff_hits_resolve(self._h, hp))
        return _take_hits(hp)

    def discover_bulge(self, guides, max_mismatch: int = 4, maximum_off_targets: int = 2000,
                       bulge_flags: int = N.FF_BULGE_RNA | N.FF_BULGE_DNA, positions: bool = False) -> Hits:
        """ff_discover_bulge (extension: 1-bp RNA / DNA bulges; see include/flashfry_b200.h)."""
        g = _u64(guides)
        hp = C.POINTER(N.FFHits)()
        N.check(N.lib().ff_discover_bulge(self._h, g.ctypes.data_as(C.POINTER(C.c_uint64)), len(g), max_mismatch,
                                          maximum_off_targets, bulge_flags, int(positions), C.byref(hp)))
        return _take_hits(hp)

    def discover_bulge_device(self, d_guides_ptr: int, n_guides: int, max_mismatch: int = 4, maximum_off_targets: int = 2000,
                              bulge_flags: int = N.FF_BULGE_RNA | N.FF_BULGE_DNA) -> N.FFDeviceResult:
        r = N.FFDeviceResult()
        N.check(N.lib().ff_discover_bulge_device(self._h, C.c_void_p(d_guides_ptr), n_guides, max_mismatch, maximum_off_targets,
                                                 bulge_flags, C.byref(r)))
        return r

    def discover_score(self, guides, max_mismatch: int = 4, maximum_off_targets: int = 2000, positions: bool = False,
                       metrics: int = FF_METRIC_CFD | FF_METRIC_HSU2013):
        g = _u64(guides)
        n = max(len(g), 1)
        cmax, cspec, hsu = np.full(n, np.nan), np.full(n, np.nan), np.full(n, np.nan)
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        hp = C.POINTER(N.FFHits)()
        N.check(N.lib().ff_discover_score(self._h, g.ctypes.data_as(C.POINTER(C.c_uint64)), len(g), max_mismatch,
                                          maximum_off_targets, int(positions), metrics, C.byref(hp), dp(cmax), dp(cspec), dp(hsu)))
        return _take_hits(hp), cmax[:len(g)], cspec[:len(g)], hsu[:len(g)]

    def score(self, guides, row_ptr, targets, metrics: int = FF_METRIC_CFD | FF_METRIC_HSU2013):
        """ff_score over a host CSR hit list -> (cfd_max, cfd_specificity, hsu2013, per_ot_cfd)."""
        g = _u64(guides)
        rp = np.ascontiguousarray(np.asarray(row_ptr, dtype=np.int64))
        t = _u64(targets)
        h = N.FFHits()
        h.n_guides, h.n_hits = len(g), len(t)
        h.row_ptr = rp.ctypes.data_as(C.POINTER(C.c_int64))
        h.targets = t.ctypes.data_as(C.POINTER(C.c_uint64))
        n = max(len(g), 1)
        cmax, cspec, hsu, per = np.full(n, np.nan), np.full(n, np.nan), np.full(n, np.nan), np.full(max(len(t), 1), np.nan)
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        N.check(N.lib().ff_score(self._h, g.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(h), metrics, dp(cmax), dp(cspec),
                                 dp(hsu), dp(per)))
        return cmax[:len(g)], cspec[:len(g)], hsu[:len(g)], per[:len(t)]

    def hit_aggregates(self, enzyme_index: int, guides, row_ptr, targets):
        """ff_hit_aggregates -> (closest, closest_count, hist[n][5], in_genome) int32 arrays."""
        g = _u64(guides)
        rp = np.ascontiguousarray(np.asarray(row_ptr, dtype=np.int64))
        t = _u64(targets)
        h = N.FFHits()
        h.n_guides, h.n_hits = len(g), len(t)
        h.row_ptr = rp.ctypes.data_as(C.POINTER(C.c_int64))
        h.targets = t.ctypes.data_as(C.POINTER(C.c_uint64))
        n = max(len(g), 1)
        closest, cnt, hist, ing = (np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros((n, 5), np.int32), np.zeros(n, np.int32))
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        N.check(N.lib().ff_hit_aggregates(self._h, enzyme_index, g.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(h), ip(closest), ip(cnt),
                                          ip(hist), ip(ing)))
        return closest[:len(g)], cnt[:len(g)], hist[:len(g)], ing[:len(g)]

    def discover_device(self, d_guides_ptr: int, n_guides: int, max_mismatch: int = 4, maximum_off_targets: int = 2000,
                        metrics: int = 0) -> N.FFDeviceResult:
        r = N.FFDeviceResult()
        N.check(N.lib().ff_discover_device(self._h, C.c_void_p(d_guides_ptr), n_guides, max_mismatch, maximum_off_targets,
                                           metrics, C.byref(r)))
        return r

    # ---- database-sharded discover over NVLink peer memory (one process per GPU: exchange the handles, e.g. with
    # torch.distributed.all_gather; include/flashfry_b200.h, ff_peer_*) ----
    def peer_export(self, hit_cap: int = 0, guide_cap: int = 0) -> bytes:
        """Make this rank's exchange block; returns its CUDA IPC handle (64 bytes) for the other ranks."""
        h = (C.c_ubyte * N.FF_PEER_HANDLE_BYTES)()
        N.check(N.lib().ff_peer_export(self._h, hit_cap, guide_cap, C.cast(h, C.c_void_p), None))
        return bytes(h)

    def peer_attach(self, rank: int, world: int, handles: Sequence[bytes]):
        """Map the exchange blocks of all ranks (handles[r] from rank r's peer_export)."""
        buf = b"".join(handles)
        assert len(buf) == world * N.FF_PEER_HANDLE_BYTES
        arr = (C.c_ubyte * len(buf)).from_buffer_copy(buf)
        N.check(N.lib().ff_peer_attach(self._h, rank, world, C.cast(arr, C.c_void_p), None))

    def peer_detach(self):
        N.check(N.lib().ff_peer_detach(self._h))

    def discover_sharded_device(self, d_guides_all_ptr: int, n_guides_all: int, max_mismatch: int = 4,
                                maximum_off_targets: int = 2000, metrics: int = 0) -> N.FFDeviceResult:
        """All guides in (identical on every rank), the rows of this rank's guides out (ff_shard_range)."""
        r = N.FFDeviceResult()
        N.check(N.lib().ff_discover_sharded_device(self._h, C.c_void_p(d_guides_all_ptr), n_guides_all, max_mismatch,
                                                   maximum_off_targets, metrics, C.byref(r)))
        return r

    def discover_sharded(self, guides_all, max_mismatch: int = 4, maximum_off_targets: int = 2000) -> "Hits":
        g = _u64(guides_all)
        hp = C.POINTER(N.FFHits)()
        N.check(N.lib().ff_discover_sharded(self._h, g.ctypes.data_as(C.POINTER(C.c_uint64)), len(g), max_mismatch,
                                            maximum_off_targets, C.byref(hp)))
        return _take_hits(hp)

    def peer_totals_ptr(self) -> int:
        """Device pointer to the all-gathered per-guide totals (int32[n_guides_all]) after a sharded call."""
        return int(N.lib().ff_peer_totals_device(self._h) or 0)

    def timings(self) -> N.FFTimings:
        t = N.FFTimings()
        N.check(N.lib().ff_last_timings(self._h, C.byref(t)))
        return t


def shard_range(n_guides: int, n_shards: int, shard: int):
    """ff_shard_range: (first, count) of a shard of the guide array (pure function of the C ABI)."""
    first, count = C.c_int64(), C.c_int64()
    N.lib().ff_shard_range(n_guides, n_shards, shard, C.byref(first), C.byref(count))
    return int(first.value), int(count.value)


class MultiContext:
    """Several GPUs behind one process (ff_multi): guide-sharded discover + one NCCL all-gather of the totals."""

    def __init__(self, devices: Sequence[int]):
        self._h = C.c_void_p()
        arr = (C.c_int * len(devices))(*devices)
        N.check(N.lib().ff_multi_create(C.byref(self._h), arr, len(devices)))
        self.n = len(devices)

    def close(self):
        if self._h:
            N.lib().ff_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_option(self, key: str, value: int):
        N.check(N.lib().ff_multi_set_option(self._h, key.encode(), int(value)))

    def synth_database(self, enzyme_index: int, n_targets: int, seed: int):
        N.check(N.lib().ff_multi_synth_database(self._h, enzyme_index, n_targets, seed))

    def load_database(self, db_path: str, header_path: Optional[str] = None):
        N.check(N.lib().ff_multi_load_database(self._h, db_path.encode(), header_path.encode() if header_path else None))

    def rank_timings(self, rank: int) -> N.FFTimings:
        t = N.FFTimings()
        N.check(N.lib().ff_last_timings(N.lib().ff_multi_ctx(self._h, rank), C.byref(t)))
        return t

    def discover_raw(self, guides_ptr, n_guides: int, max_mismatch: int, maximum_off_targets: int, totals: np.ndarray):
        """The bare call (bench timing): hit lists are freed immediately, totals land in `totals` (int32[n_guides])."""
        hp = (C.POINTER(N.FFHits) * self.n)()
        N.check(N.lib().ff_multi_discover(self._h, guides_ptr, n_guides, max_mismatch, maximum_off_targets, 0, hp,
                                          totals.ctypes.data_as(C.POINTER(C.c_int32))))
        hits = sum(int(hp[r].contents.n_hits) for r in range(self.n))
        for r in range(self.n):
            N.lib().ff_hits_free(hp[r])
        return hits

    def discover(self, guides, max_mismatch: int = 4, maximum_off_targets: int = 2000, positions: bool = False):
        """-> (list of per-shard Hits, all-gathered total_count int32[n_guides])."""
        g = _u64(guides)
        hp = (C.POINTER(N.FFHits) * self.n)()
        totals = np.zeros(max(len(g), 1), np.int32)
        N.check(N.lib().ff_multi_discover(self._h, g.ctypes.data_as(C.POINTER(C.c_uint64)), len(g), max_mismatch, maximum_off_targets,
                                          int(positions), hp, totals.ctypes.data_as(C.POINTER(C.c_int32))))
        return [_take_hits(hp[r]) for r in range(self.n)], totals[:len(g)]
