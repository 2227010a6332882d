"""ctypes binding of libflashfry_b200.so (include/flashfry_b200.h).  No fallback: if the CUDA library is missing
or no GPU is present, loading / ff_create fails loudly."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FLASHFRY_B200_LIB") or os.path.join(HERE, "libflashfry_b200.so")

FF_METRIC_CFD = 1
FF_METRIC_HSU2013 = 2
FF_BULGE_RNA = 1
FF_BULGE_DNA = 2

ERRORS = {0: "FF_OK", -1: "FF_EINVAL", -2: "FF_ENODEVICE", -3: "FF_ECUDA", -4: "FF_EIO", -5: "FF_EFORMAT",
          -6: "FF_ENODB", -7: "FF_EUNSUPPORTED", -8: "FF_ENOMEM"}


class FlashFryError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__("%s: %s" % (ERRORS.get(code, str(code)), message))
        self.code = code


class FFHits(C.Structure):
    _fields_ = [("n_guides", C.c_int64), ("n_hits", C.c_int64), ("row_ptr", C.POINTER(C.c_int64)),
                ("targets", C.POINTER(C.c_uint64)), ("mismatches", C.POINTER(C.c_uint8)),
                ("pos_ptr", C.POINTER(C.c_int64)), ("positions", C.POINTER(C.c_uint64)),
                ("total_count", C.POINTER(C.c_int32)), ("overflowed", C.POINTER(C.c_uint8)),
                ("n_compares", C.c_uint64), ("n_candidate_hits", C.c_uint64), ("opaque", C.c_void_p),
                ("bulge", C.POINTER(C.c_uint8)), ("target_index", C.POINTER(C.c_uint32))]


class FFTsvGuide(C.Structure):
    _fields_ = [("contig", C.c_char_p), ("start", C.c_int32), ("bases", C.c_char_p), ("context", C.c_char_p), ("forward", C.c_int32)]


class FFDbInfo(C.Structure):
    _fields_ = [("enzyme_index", C.c_int), ("bin_width", C.c_int), ("scan_len", C.c_int), ("pam_len", C.c_int),
                ("five_prime_pam", C.c_int), ("cmp_mask", C.c_uint64), ("n_targets", C.c_uint64),
                ("n_positions", C.c_uint64), ("n_contigs", C.c_int), ("seed_split_a", C.c_int),
                ("device_bytes", C.c_uint64)]


class FFDeviceResult(C.Structure):
    _fields_ = [("n_guides", C.c_int64), ("n_hits", C.c_int64), ("n_candidate_hits", C.c_uint64),
                ("n_compares", C.c_uint64), ("d_row_ptr", C.c_void_p), ("d_targets", C.c_void_p),
                ("d_mismatches", C.c_void_p), ("d_total_count", C.c_void_p), ("d_overflowed", C.c_void_p),
                ("d_cfd_max", C.c_void_p), ("d_cfd_specificity", C.c_void_p), ("d_hsu2013", C.c_void_p),
                ("d_bulge", C.c_void_p)]


class FFTimings(C.Structure):
    _fields_ = [("prep_ms", C.c_float), ("scan_ms", C.c_float), ("order_ms", C.c_float), ("cut_ms", C.c_float),
                ("score_ms", C.c_float), ("total_ms", C.c_float), ("scan_launches", C.c_int),
                ("kernel_launches", C.c_int), ("scan_bytes_read", C.c_uint64), ("scan_part1_ms", C.c_float),
                ("scan_part2_ms", C.c_float), ("entries_part1", C.c_uint64), ("entries_part2", C.c_uint64)]


# every symbol include/flashfry_b200.h declares (tests/test_abi.py checks the list against the header)
SYMBOLS = ["ff_create", "ff_destroy", "ff_last_error", "ff_abi_version", "ff_set_stream", "ff_set_option", "ff_load_database",
           "ff_save_image", "ff_load_image", "ff_load_database_arrays", "ff_synth_database", "ff_synth_database_skewed", "ff_db_info", "ff_db_contig", "ff_db_copy_targets",
           "ff_discover", "ff_hits_write_tsv", "ff_discover_bulge", "ff_discover_bulge_device", "ff_hits_free", "ff_db_host_targets", "ff_hits_resolve", "ff_score", "ff_score_enzyme", "ff_hit_aggregates", "ff_discover_score", "ff_discover_device", "ff_last_timings",
           "ff_multi_create", "ff_multi_destroy", "ff_multi_size", "ff_multi_ctx", "ff_multi_set_option", "ff_multi_load_database",
           "ff_multi_synth_database", "ff_shard_range", "ff_multi_discover", "ff_multi_device_totals",
           "ff_peer_export", "ff_peer_attach", "ff_peer_detach", "ff_discover_sharded_device", "ff_discover_sharded",
           "ff_peer_totals_device"]
FF_PEER_HANDLE_BYTES = 64

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FlashFryError(-2, "libflashfry_b200.so is not built (%s); run `python -m flashfry_b200.build` -- "
                                "there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, u64p, i64p, dp = C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_int64), C.POINTER(C.c_double)
    L.ff_create.argtypes = [C.POINTER(vp), C.c_int]
    L.ff_destroy.argtypes = [vp]
    L.ff_destroy.restype = None
    L.ff_last_error.restype = C.c_char_p
    L.ff_set_stream.argtypes = [vp, vp]
    L.ff_set_option.argtypes = [vp, C.c_char_p, C.c_longlong]
    L.ff_load_database.argtypes = [vp, C.c_char_p, C.c_char_p]
    L.ff_save_image.argtypes = [vp, C.c_char_p]
    L.ff_load_image.argtypes = [vp, C.c_char_p]
    L.ff_load_database_arrays.argtypes = [vp, C.c_int, C.c_int, u64p, C.c_uint64, u64p, C.c_uint64,
                                          C.POINTER(C.c_char_p), C.c_int]
    L.ff_synth_database.argtypes = [vp, C.c_int, C.c_uint64, C.c_uint64]
    L.ff_synth_database_skewed.argtypes = [vp, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int]
    L.ff_db_info.argtypes = [vp, C.POINTER(FFDbInfo)]
    L.ff_db_contig.argtypes = [vp, C.c_int]
    L.ff_db_contig.restype = C.c_char_p
    L.ff_db_copy_targets.argtypes = [vp, C.c_uint64, C.c_uint64, u64p]
    L.ff_discover.argtypes = [vp, u64p, C.c_int64, C.c_int, C.c_int, C.c_int, C.POINTER(C.POINTER(FFHits))]
    L.ff_discover_bulge.argtypes = [vp, u64p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.POINTER(FFHits))]
    L.ff_discover_bulge_device.argtypes = [vp, vp, C.c_int64, C.c_int, C.c_int, C.c_int, C.POINTER(FFDeviceResult)]
    L.ff_hits_write_tsv.argtypes = [vp, C.c_char_p, C.POINTER(FFTsvGuide), C.POINTER(FFHits), C.c_int]
    L.ff_hits_free.argtypes = [C.POINTER(FFHits)]
    L.ff_hits_free.restype = None
    L.ff_db_host_targets.argtypes = [vp]
    L.ff_db_host_targets.restype = u64p
    L.ff_hits_resolve.argtypes = [vp, C.POINTER(FFHits)]
    L.ff_score.argtypes = [vp, u64p, C.POINTER(FFHits), C.c_uint32, dp, dp, dp, dp]
    L.ff_score_enzyme.argtypes = [vp, C.c_int, u64p, C.POINTER(FFHits), C.c_uint32, dp, dp, dp, dp]
    i32p = C.POINTER(C.c_int32)
    L.ff_hit_aggregates.argtypes = [vp, C.c_int, u64p, C.POINTER(FFHits), i32p, i32p, i32p, i32p]
    L.ff_discover_score.argtypes = [vp, u64p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_uint32,
                                    C.POINTER(C.POINTER(FFHits)), dp, dp, dp]
    L.ff_discover_device.argtypes = [vp, vp, C.c_int64, C.c_int, C.c_int, C.c_uint32, C.POINTER(FFDeviceResult)]
    L.ff_last_timings.argtypes = [vp, C.POINTER(FFTimings)]
    L.ff_multi_create.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), C.c_int]
    L.ff_multi_destroy.argtypes = [vp]
    L.ff_multi_destroy.restype = None
    L.ff_multi_size.argtypes = [vp]
    L.ff_multi_ctx.argtypes = [vp, C.c_int]
    L.ff_multi_ctx.restype = vp
    L.ff_multi_set_option.argtypes = [vp, C.c_char_p, C.c_longlong]
    L.ff_multi_load_database.argtypes = [vp, C.c_char_p, C.c_char_p]
    L.ff_multi_synth_database.argtypes = [vp, C.c_int, C.c_uint64, C.c_uint64]
    L.ff_shard_range.argtypes = [C.c_int64, C.c_int, C.c_int, i64p, i64p]
    L.ff_shard_range.restype = None
    L.ff_multi_discover.argtypes = [vp, u64p, C.c_int64, C.c_int, C.c_int, C.c_int, C.POINTER(C.POINTER(FFHits)), i32p]
    L.ff_multi_device_totals.argtypes = [vp, C.c_int]
    L.ff_multi_device_totals.restype = vp
    L.ff_peer_export.argtypes = [vp, C.c_uint64, C.c_int64, vp, C.POINTER(vp)]
    L.ff_peer_attach.argtypes = [vp, C.c_int, C.c_int, vp, C.POINTER(vp)]
    L.ff_peer_detach.argtypes = [vp]
    L.ff_discover_sharded_device.argtypes = [vp, vp, C.c_int64, C.c_int, C.c_int, C.c_uint32, C.POINTER(FFDeviceResult)]
    L.ff_discover_sharded.argtypes = [vp, u64p, C.c_int64, C.c_int, C.c_int, C.POINTER(C.POINTER(FFHits))]
    L.ff_peer_totals_device.argtypes = [vp]
    L.ff_peer_totals_device.restype = vp
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise FlashFryError(rc, lib().ff_last_error().decode("utf-8", "replace"))
