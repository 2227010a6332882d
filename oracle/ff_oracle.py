"""Python side of the CPU oracle: ctypes bindings to ``ff_oracle.c`` plus numpy restatements of the
reference's data-format code (site finder, ``index``, BGZF database + ``.header``, discover/score TSV).

TEST INFRASTRUCTURE ONLY -- never imported by the product package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs use it.

Parity pin status: PINNED (integration-test md5s ``895e282b...`` / ``804bf3c1...`` and the unit-test
known answers of the reference, see ``tests/test_oracle_pins.py`` and ``tests/golden/make_golden.py``).

Citations are ``file:line`` inside the FlashFry checkout (``src/main/scala/...``).
"""
from __future__ import annotations

import ctypes as C
import gzip
import hashlib
import io
import os
import struct
import subprocess
import zlib
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(HERE, "_build", "libff_oracle.so")

STRING_MASK = 0xFFFFFFFFFFFF  # bitcoding/BitEncoding.scala:206
MAGIC = 0x1234ABCDE123890     # reference/binary/BinaryConstants.scala:26
VERSION = 1


# ------------------------------------------------------------------------------------------------
# enzyme parameter packs -- standards/StandardScanParameters.scala:61-80,90-215
@dataclass(frozen=True)
class ParameterPack:
    name: str
    index: int
    scan_len: int
    pam_len: int
    five_prime: bool
    cmp_mask: int
    # per-position allowed letters restating fwdRegex / revRegex (one consumed char + look-ahead)
    fwd: Tuple[str, ...]
    rev: Tuple[str, ...]

    @property
    def is_cas9_23(self) -> bool:  # validOverEnzyme of both scorers (Doench2016CFDScore.scala:96-98)
        return (not self.five_prime) and self.scan_len == 23


def _pat(*parts):
    out = []
    for letters, n in parts:
        out.extend([letters] * n)
    return tuple(out)


N = "ACGT"
PACKS: Dict[str, ParameterPack] = {
    # :199-215  fwd (T)(?=(TT[ACGT]{21}))  rev ([ACGT])(?=([ACGT]{20}AAA))
    "CPF1": ParameterPack("CPF1", 1, 24, 4, True, 0x00FFFFFFFFFF, _pat(("T", 3), (N, 21)), _pat((N, 21), ("A", 3))),
    # :90-109   fwd [ACGT]{21}[AG]G   rev C[CT][ACGT]{21}
    "SPCAS9": ParameterPack("SPCAS9", 2, 23, 3, False, 0x3FFFFFFFFFC0, _pat((N, 21), ("AG", 1), ("G", 1)), _pat(("C", 1), ("CT", 1), (N, 21))),
    # :134-153
    "SPCAS9NGG": ParameterPack("SPCAS9NGG", 3, 23, 3, False, 0x3FFFFFFFFFC0, _pat((N, 21), ("G", 2)), _pat(("C", 2), (N, 21))),
    # :178-197
    "SPCAS9NAG": ParameterPack("SPCAS9NAG", 4, 23, 3, False, 0x3FFFFFFFFFC0, _pat((N, 21), ("A", 1), ("G", 1)), _pat(("C", 1), ("T", 1), (N, 21))),
    # :112-131
    "SPCAS919": ParameterPack("SPCAS919", 5, 22, 3, False, 0x0FFFFFFFFFC0, _pat((N, 20), ("AG", 1), ("G", 1)), _pat(("C", 1), ("CT", 1), (N, 20))),
    # :156-175
    "SPCAS9NGG19": ParameterPack("SPCAS9NGG19", 6, 22, 3, False, 0x0FFFFFFFFFC0, _pat((N, 20), ("G", 2)), _pat(("C", 2), (N, 20))),
}
PACK_BY_INDEX = {p.index: p for p in PACKS.values()}


def pack_by_name(name: str) -> ParameterPack:  # ParameterPack.nameToParameterPack :50-59
    try:
        return PACKS[name.upper()]
    except KeyError:
        raise ValueError("Unable to find the correct parameter pack for enzyme: " + name)


# ------------------------------------------------------------------------------------------------
# ctypes binding of ff_oracle.c
class _CPack(C.Structure):
    _fields_ = [("enzyme_index", C.c_int), ("scan_len", C.c_int), ("pam_len", C.c_int),
                ("five_prime", C.c_int), ("cmp_mask", C.c_uint64)]


class _CHits(C.Structure):
    _fields_ = [("n_guides", C.c_int64), ("row_ptr", C.POINTER(C.c_int64)), ("targets", C.POINTER(C.c_uint64)),
                ("mismatches", C.POINTER(C.c_uint8)), ("pos_ptr", C.POINTER(C.c_int64)),
                ("positions", C.POINTER(C.c_uint64)), ("total_count", C.POINTER(C.c_int32)),
                ("overflowed", C.POINTER(C.c_uint8)), ("n_compares", C.c_uint64),
                ("n_target_compares", C.c_uint64), ("n_targets_scanned", C.c_uint64),
                ("saturated", C.c_int), ("bins_visited", C.c_int), ("bulge", C.POINTER(C.c_uint8))]


_lib = None


def build_lib(force: bool = False) -> str:
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(os.path.join(HERE, f)) > os.path.getmtime(_LIB_PATH)
            for f in ("ff_oracle.c", "ff_oracle.h", "score_tables.h")):
        subprocess.check_call(["make", "-s", "-C", HERE], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_lib())
        u64p, i64p = C.POINTER(C.c_uint64), C.POINTER(C.c_int64)
        _lib.ffo_pack_from_index.argtypes = [C.c_int, C.POINTER(_CPack)]
        _lib.ffo_encode.restype = C.c_uint64
        _lib.ffo_encode.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int)]
        _lib.ffo_decode.argtypes = [C.c_uint64, C.c_int, C.c_char_p]
        _lib.ffo_mismatches.argtypes = [C.POINTER(_CPack), C.c_uint64, C.c_uint64, C.c_uint64]
        _lib.ffo_bin_comparitor.argtypes = [C.POINTER(_CPack), C.c_uint64, C.c_int, C.c_int, u64p, u64p]
        _lib.ffo_mismatch_bin.argtypes = [C.POINTER(_CPack), C.c_uint64, C.c_uint64, C.c_uint64]
        _lib.ffo_discover_blocks.argtypes = [C.POINTER(_CPack), C.c_int, u64p, i64p, u64p, C.c_int64, C.c_int,
                                             C.c_int, C.c_int, C.POINTER(C.POINTER(_CHits))]
        _lib.ffo_discover_soa.argtypes = [C.POINTER(_CPack), C.c_int, u64p, i64p, u64p, C.c_int64, C.c_int,
                                          C.c_int, C.c_int, C.POINTER(C.POINTER(_CHits))]
        _lib.ffo_hits_free.argtypes = [C.POINTER(_CHits)]
        _lib.ffo_bulge_align.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                         C.POINTER(C.c_int)]
        _lib.ffo_bulge_align.restype = None
        _lib.ffo_discover_bulge.argtypes = [C.POINTER(_CPack), u64p, C.c_int64, u64p, C.c_int64, C.c_int, C.c_int,
                                            C.c_int, C.c_int, C.POINTER(C.POINTER(_CHits))]
        _lib.ffo_cfd_pair.restype = C.c_double
        _lib.ffo_cfd_pair.argtypes = [C.c_uint64, C.c_uint64]
        _lib.ffo_cfd_guide.argtypes = [C.c_uint64, u64p, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                       C.POINTER(C.c_double)]
        _lib.ffo_hsu_offtarget.restype = C.c_double
        _lib.ffo_hsu_offtarget.argtypes = [C.c_uint64, C.c_uint64]
        _lib.ffo_hsu_guide.restype = C.c_double
        _lib.ffo_hsu_guide.argtypes = [C.POINTER(_CPack), C.c_uint64, u64p, C.c_int64]
    return _lib


def _cpack(pack: ParameterPack) -> _CPack:
    cp = _CPack()
    if lib().ffo_pack_from_index(pack.index, C.byref(cp)) != 0:
        raise ValueError("bad enzyme index")
    assert cp.scan_len == pack.scan_len and cp.cmp_mask == pack.cmp_mask and bool(cp.five_prime) == pack.five_prime
    return cp


def _u64p(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def _i64p(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


# -- bit kernels -----------------------------------------------------------------------------------
def encode(bases: str, count: int = 1) -> int:
    """BitEncoding.bitEncodeString (BitEncoding.scala:46-67)."""
    err = C.c_int(0)
    v = lib().ffo_encode(bases.encode(), len(bases), count, C.byref(err))
    if err.value:
        raise ValueError("Unable to encode " + bases + " count " + str(count))
    return int(v)


def decode(enc: int, length: int) -> Tuple[str, int]:
    """BitEncoding.bitDecodeString (BitEncoding.scala:85-99) -> (bases, count)."""
    buf = C.create_string_buffer(length + 1)
    cnt = lib().ffo_decode(enc, length, buf)
    return buf.value.decode(), cnt


def mismatches(pack: ParameterPack, a: int, b: int, additional_mask: int = STRING_MASK) -> int:
    """BitEncoding.mismatches (BitEncoding.scala:127-132)."""
    return lib().ffo_mismatches(C.byref(_cpack(pack)), a, b, additional_mask)


def bin_code(bin_str: str) -> int:
    v = 0
    for ch in bin_str:
        v = v * 4 + "ACGT".index(ch)
    return v


def bin_comparitor(pack: ParameterPack, bin_str: str, right_shift_bases: int = 0) -> Tuple[int, int]:
    """BitEncoding.binToLongComparitor (BitEncoding.scala:153-170) -> (binLong, guideMask)."""
    bl, bm = C.c_uint64(0), C.c_uint64(0)
    lib().ffo_bin_comparitor(C.byref(_cpack(pack)), bin_code(bin_str), len(bin_str), right_shift_bases,
                             C.byref(bl), C.byref(bm))
    return bl.value, bm.value


def mismatch_bin(pack: ParameterPack, bin_str: str, guide: int, right_shift_bases: int = 0) -> int:
    """BitEncoding.mismatchBin (BitEncoding.scala:142-144)."""
    bl, bm = bin_comparitor(pack, bin_str, right_shift_bases)
    return lib().ffo_mismatch_bin(C.byref(_cpack(pack)), bl, bm, guide)


# -- position longs -- bitcoding/BitPosition.scala:51-92 ------------------------------------------
def pos_encode(contig_id: int, start: int, length: int, forward: bool) -> int:
    return (contig_id << 32) | start | (0 if forward else (1 << 60)) | (length << 52)


def pos_decode(p: int) -> Tuple[int, int, int, bool]:
    return (p >> 32) & 0xFFFFF, p & 0xFFFFFFFF, (p >> 52) & 0xFF, ((p >> 60) & 0xF) == 0


# ------------------------------------------------------------------------------------------------
# discovery results
@dataclass
class Hits:
    row_ptr: np.ndarray
    targets: np.ndarray
    mismatches: np.ndarray
    total_count: np.ndarray
    overflowed: np.ndarray
    pos_ptr: Optional[np.ndarray] = None
    positions: Optional[np.ndarray] = None
    n_compares: int = 0
    n_target_compares: int = 0
    n_targets_scanned: int = 0
    saturated: bool = False
    bins_visited: int = 0
    bulge: Optional[np.ndarray] = None   # extension (discover_bulge): 0 none, 0x40|q RNA bulge, 0x80|q DNA bulge

    def row(self, g: int) -> Tuple[np.ndarray, np.ndarray]:
        lo, hi = int(self.row_ptr[g]), int(self.row_ptr[g + 1])
        return self.targets[lo:hi], self.mismatches[lo:hi]

    def row_positions(self, g: int) -> List[np.ndarray]:
        lo, hi = int(self.row_ptr[g]), int(self.row_ptr[g + 1])
        return [self.positions[int(self.pos_ptr[i]):int(self.pos_ptr[i + 1])] for i in range(lo, hi)]


def _take(hp, with_pos: bool) -> Hits:
    h = hp.contents
    n = h.n_guides
    row_ptr = np.ctypeslib.as_array(h.row_ptr, shape=(n + 1,)).copy()
    nh = int(row_ptr[-1])
    targets = np.ctypeslib.as_array(h.targets, shape=(max(nh, 1),))[:nh].copy()
    mm = np.ctypeslib.as_array(h.mismatches, shape=(max(nh, 1),))[:nh].copy()
    total = np.ctypeslib.as_array(h.total_count, shape=(max(n, 1),))[:n].copy()
    ovf = np.ctypeslib.as_array(h.overflowed, shape=(max(n, 1),))[:n].copy()
    pos_ptr = positions = None
    if with_pos and h.pos_ptr:
        pos_ptr = np.ctypeslib.as_array(h.pos_ptr, shape=(nh + 1,)).copy()
        npos = int(pos_ptr[-1])
        positions = np.ctypeslib.as_array(h.positions, shape=(max(npos, 1),))[:npos].copy()
    out = Hits(row_ptr, targets, mm, total, ovf, pos_ptr, positions, int(h.n_compares), int(h.n_target_compares),
               int(h.n_targets_scanned), bool(h.saturated), int(h.bins_visited))
    if h.bulge:
        out.bulge = np.ctypeslib.as_array(h.bulge, shape=(max(nh, 1),))[:nh].copy()
    lib().ffo_hits_free(hp)
    return out


@dataclass
class Database:
    """An inflated FlashFry database: every bin block concatenated, plus the header contents."""
    pack: ParameterPack
    bin_width: int
    longs: np.ndarray               # uint64, all blocks in bin order
    bin_off: np.ndarray             # int64 [n_bins+1], offsets in longs
    n_targets: np.ndarray           # int32 [n_bins] (header numberOfTargets)
    contigs: List[str] = field(default_factory=list)

    def soa(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
        """Decode the blocks (BlockManager.scala:266-351) into targets / positions arrays in database order.
        Returns (targets u64[N_t], bin_off_targets i64[n_bins+1], pos_off i64[N_t+1], positions u64[N_p])."""
        n_bins = len(self.bin_off) - 1
        tg, po, bo = [], [], np.zeros(n_bins + 1, np.int64)
        L = self.longs
        for b in range(n_bins):
            lo, hi = int(self.bin_off[b]), int(self.bin_off[b + 1])
            typ = int(L[lo])
            if typ == 1:
                i = lo + 1
            elif typ == 2:
                i = lo + 1 + 256
            else:
                raise ValueError("Invalid bin type, unknown value: %d" % typ)
            blk = L[i:hi]
            # walk [target, pos x count]* -- vectorised: find target slots by cumulative skipping
            j, n = 0, len(blk)
            idx = []
            while j < n:
                idx.append(j)
                j += 1 + (int(blk[j]) >> 48)
            idx = np.asarray(idx, np.int64)
            is_t = np.zeros(n, bool)
            is_t[idx] = True
            tg.append(blk[is_t])
            po.append(blk[~is_t])
            bo[b + 1] = bo[b] + len(idx)
        targets = np.concatenate(tg) if tg else np.zeros(0, np.uint64)
        positions = np.concatenate(po) if po else np.zeros(0, np.uint64)
        counts = (targets >> np.uint64(48)).astype(np.int64)
        pos_off = np.zeros(len(targets) + 1, np.int64)
        np.cumsum(counts, out=pos_off[1:])
        return targets, bo, pos_off, positions


def discover_blocks(db: Database, guides: Sequence[int], max_mismatch: int = 4, max_off_targets: int = 2000,
                    force_linear: bool = False) -> Hits:
    """The reference's `discover` scan over its own block format (SeekTraverser / LinearTraverser)."""
    g = np.ascontiguousarray(np.asarray(guides, dtype=np.uint64))
    longs = np.ascontiguousarray(db.longs, dtype=np.uint64)
    off = np.ascontiguousarray(db.bin_off, dtype=np.int64)
    hp = C.POINTER(_CHits)()
    rc = lib().ffo_discover_blocks(C.byref(_cpack(db.pack)), db.bin_width, _u64p(longs), _i64p(off), _u64p(g),
                                   len(g), max_mismatch, max_off_targets, int(force_linear), C.byref(hp))
    if rc != 0:
        raise RuntimeError("Invalid bin type (rc=%d)" % rc)
    return _take(hp, True)


def discover_soa(pack: ParameterPack, bin_width: int, targets: np.ndarray, bin_off: np.ndarray,
                 guides: Sequence[int], max_mismatch: int = 4, max_off_targets: int = 2000,
                 n_threads: int = 1) -> Hits:
    g = np.ascontiguousarray(np.asarray(guides, dtype=np.uint64))
    t = np.ascontiguousarray(targets, dtype=np.uint64)
    off = np.ascontiguousarray(bin_off, dtype=np.int64)
    hp = C.POINTER(_CHits)()
    rc = lib().ffo_discover_soa(C.byref(_cpack(pack)), bin_width, _u64p(t), _i64p(off), _u64p(g), len(g),
                                max_mismatch, max_off_targets, n_threads, C.byref(hp))
    if rc != 0:
        raise RuntimeError("discover_soa failed rc=%d" % rc)
    return _take(hp, False)


# -- EXTENSION: 1-bp bulge mode (not in the reference; PARITY UNPINNED -- semantics defined in ff_oracle.h) ------
BULGE_RNA, BULGE_DNA = 1, 2


def bulge_align(guide: int, target: int, flags: int = 3) -> Tuple[int, int, int]:
    """(mismatches, type 0 none / 1 RNA / 2 DNA, bulge position q) of the best alignment (ffo_bulge_align)."""
    mm, ty, pos = C.c_int(), C.c_int(), C.c_int()
    lib().ffo_bulge_align(guide, target, flags, C.byref(mm), C.byref(ty), C.byref(pos))
    return mm.value, ty.value, pos.value


def bulge_align_strings(g: str, t: str, flags: int = 3) -> Tuple[int, int, int]:
    """The same definition on 20-base strings, written as string surgery (independent of the C code):
    RNA bulge at q = delete guide base q, align with t[1:]; DNA bulge at q = delete target base q, align with g[1:]."""
    ham = lambda a, b: sum(x != y for x, y in zip(a, b))
    assert len(g) == 20 and len(t) == 20
    best = (ham(g, t), 0, 0)
    if flags & BULGE_RNA:
        for q in range(1, 19):
            best = min(best, (ham(g[:q] + g[q + 1:], t[1:]), 1, q))
    if flags & BULGE_DNA:
        for q in range(1, 19):
            best = min(best, (ham(g[1:], t[:q] + t[q + 1:]), 2, q))
    return best


def discover_bulge(pack: ParameterPack, targets: np.ndarray, guides: Sequence[int], max_mismatch: int = 4,
                   max_off_targets: int = 2000, flags: int = 3, n_threads: int = 1) -> Hits:
    """Brute-force bulge-mode discover over targets in database order (ffo_discover_bulge)."""
    g = np.ascontiguousarray(np.asarray(guides, dtype=np.uint64))
    t = np.ascontiguousarray(targets, dtype=np.uint64)
    hp = C.POINTER(_CHits)()
    rc = lib().ffo_discover_bulge(C.byref(_cpack(pack)), _u64p(t), len(t), _u64p(g), len(g), max_mismatch,
                                  max_off_targets, flags, n_threads, C.byref(hp))
    if rc != 0:
        raise RuntimeError("bulge mode needs a 23-bp 3'-PAM Cas9 pack (rc=%d)" % rc)
    return _take(hp, False)


def bin_offsets_from_sorted(pack: ParameterPack, bin_width: int, targets: np.ndarray) -> np.ndarray:
    """bin_off[b] for a 3'-PAM database whose targets are globally sorted (DB order == lexicographic)."""
    assert not pack.five_prime
    shift = np.uint64(2 * (pack.scan_len - bin_width))
    keys = (targets & np.uint64(STRING_MASK)) >> shift
    return np.searchsorted(keys, np.arange((1 << (2 * bin_width)) + 1, dtype=np.uint64), side="left").astype(np.int64)


# -- scorers ---------------------------------------------------------------------------------------
def cfd_pair(guide: int, ot: int) -> float:
    return lib().ffo_cfd_pair(guide, ot)


def cfd_guide(guide: int, ots: np.ndarray) -> Tuple[float, float, np.ndarray]:
    """Doench2016CFDScore.scoreGuide -> (max [thresholded], specificity, per-OT score with NaN = skipped)."""
    o = np.ascontiguousarray(ots, dtype=np.uint64)
    per = np.zeros(max(len(o), 1), np.float64)
    mx, sp = C.c_double(0), C.c_double(0)
    lib().ffo_cfd_guide(guide, _u64p(o), len(o), C.byref(mx), C.byref(sp), per.ctypes.data_as(C.POINTER(C.c_double)))
    return mx.value, sp.value, per[:len(o)]


def hsu_offtarget(guide: int, ot: int) -> float:
    return lib().ffo_hsu_offtarget(guide, ot)


def hsu_guide(pack: ParameterPack, guide: int, ots: np.ndarray) -> float:
    o = np.ascontiguousarray(ots, dtype=np.uint64)
    return lib().ffo_hsu_guide(C.byref(_cpack(pack)), guide, _u64p(o), len(o))


def minot(pack: ParameterPack, guide: int, ots: np.ndarray) -> Tuple[str, str, str]:
    """scoring/ClosestHit.scala:43-76 -> (basesDiffToClosestHit, closestHitCount, 0-1-2-3-4_mismatch)."""
    closest, count = None, 0
    hist = [0] * 5
    for t in ots:
        t = int(t)
        mm = mismatches(pack, t, guide)
        c = (t >> 48) & 0x7FFF
        if mm <= 4:
            hist[mm] += c
        if mm > 0 and (closest is None or mm < closest):
            closest, count = mm, c
        elif closest is not None and mm == closest:
            count += c
    h = ",".join(str(x) for x in hist)
    return ("UNK", "0", h) if closest is None else (str(closest), str(count), h)


def dangerous(pack: ParameterPack, bases: str, guide: int, ots: np.ndarray) -> Tuple[str, str, str]:
    """scoring/DangerousSequences.scala:49-68 (non-numeric output)."""
    gc = sum(1 for b in bases.upper() if b in "CG") / len(bases)
    p0 = "GC_" + java_double_str(gc) if (gc < .25 or gc > .75) else "NONE"
    lo, hi = (pack.pam_len, pack.scan_len) if pack.five_prime else (0, pack.scan_len - pack.pam_len)
    p1 = "PolyT" if "TTTT" in bases[lo:hi] else "NONE"
    in_genome = sum(((int(t) >> 48) & 0x7FFF) for t in ots if mismatches(pack, int(t), guide) == 0)
    p2 = "IN_GENOME=%d" % in_genome if in_genome > 0 else "NONE"
    return p0, p1, p2


# ------------------------------------------------------------------------------------------------
# Java Double.toString (Appendix B.11 of SURVEY.md): shortest round-trip digits in Java's layout
def java_double_str(x: float) -> str:
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Infinity" if x > 0 else "-Infinity"
    if x == 0:
        return "-0.0" if str(x).startswith("-") else "0.0"
    from decimal import Decimal
    sign_bit, dig, exp = Decimal(repr(float(x))).as_tuple()  # repr == shortest round-trip digits
    sign = "-" if sign_bit else ""
    digits = "".join(str(d) for d in dig).lstrip("0")
    e10 = len(digits) - 1 + exp            # decimal exponent of the first significant digit
    digits = digits.rstrip("0") or "0"
    ax = abs(x)
    if 1e-3 <= ax < 1e7:
        if e10 >= 0:
            whole = digits[:e10 + 1].ljust(e10 + 1, "0")
            frac = digits[e10 + 1:] or "0"
            return sign + whole + "." + frac
        return sign + "0." + "0" * (-e10 - 1) + digits
    frac = digits[1:] or "0"
    return sign + digits[0] + "." + frac + "E" + str(e10)


# ------------------------------------------------------------------------------------------------
# FASTA -> sites : reference/ReferenceEncoder.scala:46-70,114-169
@dataclass
class Site:
    contig: str
    bases: str
    forward: bool
    position: int
    context: Optional[str]


_COMP = str.maketrans("ACGTacgt", "TGCAtgca")


def revcomp(s: str) -> str:  # utils/Utils.scala:88
    return s.translate(_COMP)[::-1]


def read_fasta(path: str) -> List[Tuple[str, str]]:
    """ReferenceEncoder.findTargetSites :52-66 -- contig names get ' ' and tab -> '_', lines are upper-cased."""
    opener = gzip.open if path.endswith(".gz") else open
    out, name, buf = [], None, []
    with opener(path, "rt") as fh:
        for line in fh:
            line = line.rstrip("\r\n")
            if line.startswith(">"):
                if name is not None:
                    out.append((name, "".join(buf)))
                name, buf = line[1:].replace(" ", "_").replace("\t", "_"), []
            else:
                buf.append(line.upper())
    if name is not None:
        out.append((name, "".join(buf)))
    return out


def _match_positions(codes: np.ndarray, pattern: Tuple[str, ...]) -> np.ndarray:
    """All offsets i with codes[i+j] in pattern[j] for every j (regex with look-ahead => overlapping matches)."""
    L = len(pattern)
    n = len(codes) - L + 1
    if n <= 0:
        return np.zeros(0, np.int64)
    ok = np.ones(n, bool)
    for j, letters in enumerate(pattern):
        allowed = np.zeros(6, bool)
        for ch in letters:
            allowed["ACGT".index(ch)] = True
        ok &= allowed[codes[j:j + n]]
    return np.nonzero(ok)[0].astype(np.int64)


def _codes(seq: str) -> np.ndarray:
    lut = np.full(256, 4, np.uint8)
    for i, ch in enumerate("ACGT"):
        lut[ord(ch)] = i
    return lut[np.frombuffer(seq.encode("latin-1"), np.uint8)]


def _window_values(codes: np.ndarray, starts: np.ndarray, L: int, rc: bool) -> np.ndarray:
    """2-bit packed value (first base most significant) of the L-mer at each start; reverse-complemented if rc."""
    v = np.zeros(len(starts), np.uint64)
    for j in range(L):
        if rc:
            c = np.uint64(3) - codes[starts + (L - 1 - j)].astype(np.uint64)
        else:
            c = codes[starts + j].astype(np.uint64)
        v = (v << np.uint64(2)) | c
    return v


def find_target_sites(contigs: Iterable[Tuple[str, str]], pack: ParameterPack, flank: int) -> List[Site]:
    """SimpleSiteFinder.reset (ReferenceEncoder.scala:114-169): per contig all forward then all reverse matches."""
    out: List[Site] = []
    L = pack.scan_len
    for name, seq in contigs:
        codes = _codes(seq)
        for fwd, pattern in ((True, pack.fwd), (False, pack.rev)):
            for start in _match_positions(codes, pattern):
                start = int(start)
                end = start + L
                sub = seq[start:end]
                ctx = seq[max(0, start - flank):end + flank]
                if not fwd:
                    sub, ctx = revcomp(sub), revcomp(ctx)
                out.append(Site(name, sub, fwd, start, ctx if len(ctx) == L + 2 * flank else None))
    return out


def gc_content(bases: str) -> float:  # utils/Utils.scala:46
    return sum(1 for b in bases.upper() if b in "CG") / len(bases)


# ------------------------------------------------------------------------------------------------
# BGZF (SAM spec 4.1; htsjdk BlockCompressed{Input,Output}Stream is the un-vendored dependency, build.sbt:17)
BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
BGZF_BLOCK = 65498  # htsjdk DEFAULT_UNCOMPRESSED_BLOCK_SIZE = 64KiB - (header 18 + footer 8 + 2 + 10)


class BgzfWriter:
    """Mimics BlockCompressedOutputStream: write(), getPosition() -> virtual file pointer, close() -> EOF block."""

    def __init__(self, path: str, level: int = 5):
        self.fh = open(path, "wb")
        self.buf = bytearray()
        self.block_address = 0
        self.level = level

    def position(self) -> int:  # DatabaseWriter.scala:80 blockStream.getPosition
        return (self.block_address << 16) | len(self.buf)

    def _deflate(self, n: int):
        data = bytes(self.buf[:n])
        del self.buf[:n]
        co = zlib.compressobj(self.level, zlib.DEFLATED, -15)
        comp = co.compress(data) + co.flush()
        bsize = len(comp) + 25  # total member length - 1
        assert bsize < 65536
        member = (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize) + comp +
                  struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))
        self.fh.write(member)
        self.block_address += len(member)

    def write(self, data: bytes):
        mv = memoryview(data)
        while len(mv):
            room = BGZF_BLOCK - len(self.buf)
            self.buf += mv[:room]
            mv = mv[room:]
            if len(self.buf) == BGZF_BLOCK:
                self._deflate(BGZF_BLOCK)

    def close(self):
        if self.buf:
            self._deflate(len(self.buf))
        self.fh.write(BGZF_EOF)
        self.fh.close()


def bgzf_members(raw: bytes) -> List[Tuple[int, int]]:
    """[(file offset, member length)] by walking the BSIZE fields."""
    out, off = [], 0
    while off < len(raw):
        if raw[off:off + 4] != b"\x1f\x8b\x08\x04":
            raise ValueError("not a BGZF member at %d" % off)
        xlen = struct.unpack_from("<H", raw, off + 10)[0]
        p, end, bsize = off + 12, off + 12 + xlen, None
        while p < end:
            si1, si2, slen = raw[p], raw[p + 1], struct.unpack_from("<H", raw, p + 2)[0]
            if si1 == 66 and si2 == 67:
                bsize = struct.unpack_from("<H", raw, p + 4)[0]
            p += 4 + slen
        if bsize is None:
            raise ValueError("BGZF member without BC field")
        out.append((off, bsize + 1))
        off += bsize + 1
    return out


def bgzf_inflate_all(path: str) -> Tuple[bytes, Dict[int, int]]:
    """Inflate every member; returns (payload, {member file offset -> offset of its payload in the output})."""
    raw = open(path, "rb").read()
    chunks, where, total = [], {}, 0
    for off, ln in bgzf_members(raw):
        xlen = struct.unpack_from("<H", raw, off + 10)[0]
        body = raw[off + 12 + xlen: off + ln - 8]
        data = zlib.decompress(body, -15)
        crc, isize = struct.unpack_from("<II", raw, off + ln - 8)
        assert isize == len(data) and crc == (zlib.crc32(data) & 0xFFFFFFFF)
        where[off] = total
        chunks.append(data)
        total += len(data)
    return b"".join(chunks), where


# ------------------------------------------------------------------------------------------------
# header : reference/binary/BinaryHeader.scala:69-160
def bin_names(width: int) -> List[str]:
    """utils/BaseCombinationGenerator.scala:33-69 order == base-4 counting order."""
    out = []
    for v in range(4 ** width):
        out.append("".join("ACGT"[(v >> (2 * (width - 1 - i))) & 3] for i in range(width)))
    return out


def write_header(path: str, pack: ParameterPack, bin_width: int, offsets: List[Tuple[int, int, int]],
                 contigs: List[str]):
    with open(path, "w") as fh:
        fh.write("%d\n%d\n%d\n%d\n" % (MAGIC, VERSION, pack.index, 4 ** bin_width))
        for name, (vptr, nbytes, ntargets) in zip(bin_names(bin_width), offsets):
            fh.write("%s=%d,%d,%d\n" % (name, vptr, nbytes, ntargets))
        for i, c in enumerate(contigs):
            fh.write("%s=%d\n" % (c, i + 1))


def read_header(path: str):
    with open(path) as fh:
        lines = fh.read().split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    assert int(lines[0]) == MAGIC, "Binary file %s doesn't have the magic number expected at the top of the file" % path
    assert int(lines[1]) == VERSION, "Binary file %s doesn't have the correct version" % path
    pack = PACK_BY_INDEX[int(lines[2])]
    n_bins = int(lines[3])
    import math
    width = int(math.log(n_bins) / math.log(4))
    names = bin_names(width)
    offs = []
    for i, nm in enumerate(names):
        ln = lines[4 + i]
        k, v = ln.split("=")
        assert k == nm, "Failed to verify bin name, expected: %s isn't what we got %s" % (nm, k)
        a, b, c = v.split(",")[:3]
        offs.append((int(a), int(b), int(c)))
    contigs = [ln.split("=")[0] for ln in lines[4 + n_bins:] if ln]
    return pack, width, offs, contigs


def read_database(path: str) -> Database:
    """What SeekTraverser.fillBlock (SeekTraverser.scala:113-121) delivers for every bin, concatenated."""
    pack, width, offs, contigs = read_header(path + ".header")
    payload, where = bgzf_inflate_all(path)
    n_bins = len(offs)
    bin_off = np.zeros(n_bins + 1, np.int64)
    parts = []
    for b, (vptr, nbytes, _nt) in enumerate(offs):
        start = where[vptr >> 16] + (vptr & 0xFFFF)
        assert nbytes % 8 == 0
        parts.append(np.frombuffer(payload, dtype="<u8", count=nbytes // 8, offset=start))
        bin_off[b + 1] = bin_off[b] + nbytes // 8
    longs = np.concatenate(parts).astype(np.uint64)
    return Database(pack, width, longs, bin_off, np.asarray([o[2] for o in offs], np.int32), contigs)


# ------------------------------------------------------------------------------------------------
# `index` : BuildOffTargetDatabase.scala:57-89, BinWriter.scala:49-75, BlockReader.scala:87-154,
#           DatabaseWriter.scala:58-111, BlockManager.scala:362-442
def collapse_sites(values: np.ndarray, positions: np.ndarray):
    """BlockReader.loadBlock :104-126 + TargetPos.combine :141-154: sort by bases, merge equal 23-mers,
    count = min(sum, 32767), positions truncated to 32767.  Position order inside a target comes from an
    unstable quickSort in the reference (unpinned); a stable sort (discovery order) is used here."""
    order = np.argsort(values, kind="stable")
    v, p = values[order], positions[order]
    uniq, first, cnt = np.unique(v, return_index=True, return_counts=True)
    capped = np.minimum(cnt, 32767)
    keep = np.ones(len(v), bool)
    if (cnt > 32767).any():
        rank = np.arange(len(v)) - np.repeat(first, cnt)
        keep = rank < 32767
    return uniq, capped.astype(np.int64), p[keep]


def make_blocks(pack: ParameterPack, bin_width: int, uniq: np.ndarray, counts: np.ndarray, pos: np.ndarray):
    """Lay the collapsed targets out as reference blocks.  uniq must be in *database order* (bin-major)."""
    n_bins = 4 ** bin_width
    if pack.five_prime:
        bshift = np.uint64(2 * (pack.scan_len - (bin_width + pack.pam_len)))
    else:
        bshift = np.uint64(2 * (pack.scan_len - bin_width))
    bkey = ((uniq >> bshift) & np.uint64(n_bins - 1)).astype(np.int64)
    assert (np.diff(bkey) >= 0).all()
    tb = np.searchsorted(bkey, np.arange(n_bins + 1), side="left")
    pos_off = np.zeros(len(uniq) + 1, np.int64)
    np.cumsum(counts, out=pos_off[1:])
    tlong = uniq | (counts.astype(np.uint64) << np.uint64(48))
    blocks, ntargets = [], []
    sub_shift = np.uint64(2 * (pack.scan_len - (bin_width + 4))) if not pack.five_prime else np.uint64(0)
    for b in range(n_bins):
        lo, hi = int(tb[b]), int(tb[b + 1])
        n = hi - lo
        plo, phi = int(pos_off[lo]), int(pos_off[hi])
        body = np.empty(n + (phi - plo), np.uint64)
        slot = np.arange(n, dtype=np.int64) + (pos_off[lo:hi] - plo)       # index of each target long in the body
        is_t = np.zeros(len(body), bool)
        is_t[slot] = True
        body[is_t] = tlong[lo:hi]
        body[~is_t] = pos[plo:phi]
        if n > 500 and not pack.five_prime:  # DatabaseWriter.scala:85 -> createIndexedBlock(…, 4)
            sub = ((uniq[lo:hi] >> sub_shift) & np.uint64(255)).astype(np.int64)
            sizes = np.bincount(sub, weights=(1 + counts[lo:hi]), minlength=256).astype(np.int64)
            firsts = np.full(256, -1, np.int64)
            present, first_idx = np.unique(sub, return_index=True)
            firsts[present] = slot[first_idx]
            table = ((firsts.astype(np.uint64) << np.uint64(32)) | sizes.astype(np.uint64))  # BlockManager.scala:401
            blk = np.concatenate([np.asarray([2], np.uint64), table, body])
        else:                                 # createLinearBlock :424-442
            blk = np.concatenate([np.asarray([1], np.uint64), body])
        blocks.append(blk)
        ntargets.append(n)
    return blocks, ntargets


def write_database(path: str, pack: ParameterPack, bin_width: int, blocks: List[np.ndarray], ntargets: List[int],
                   contigs: List[str]):
    """DatabaseWriter.writeToBinnedFileSet :78-110."""
    w = BgzfWriter(path)
    offs = []
    for blk, nt in zip(blocks, ntargets):
        vptr = w.position()
        w.write(blk.astype("<u8").tobytes())  # Utils.longArrayToByteArray: native (little-endian) order
        offs.append((vptr, len(blk) * 8, nt))
    w.close()
    write_header(path + ".header", pack, bin_width, offs, contigs)


def sites_to_arrays(contigs: List[Tuple[str, str]], pack: ParameterPack):
    """Vectorised findTargetSites(flank=0) for whole chromosomes -> (values u64, positions u64)."""
    vals, poss = [], []
    L = pack.scan_len
    for ci, (_name, seq) in enumerate(contigs):
        codes = _codes(seq)
        for fwd, pattern in ((True, pack.fwd), (False, pack.rev)):
            st = _match_positions(codes, pattern)
            if len(st) == 0:
                continue
            vals.append(_window_values(codes, st, L, rc=not fwd))
            p = (np.uint64(ci + 1) << np.uint64(32)) | st.astype(np.uint64) | (np.uint64(L) << np.uint64(52))
            if not fwd:
                p = p | (np.uint64(1) << np.uint64(60))
            poss.append(p)
    if not vals:
        return np.zeros(0, np.uint64), np.zeros(0, np.uint64)
    return np.concatenate(vals), np.concatenate(poss)


def db_order_key(pack: ParameterPack, bin_width: int, uniq: np.ndarray) -> np.ndarray:
    """Sort key that yields database order: bin-major, then lexicographic within the bin
    (identical to plain lexicographic order for 3'-PAM enzymes)."""
    if not pack.five_prime:
        return uniq
    n_bins = 4 ** bin_width
    bshift = np.uint64(2 * (pack.scan_len - (bin_width + pack.pam_len)))
    return (((uniq >> bshift) & np.uint64(n_bins - 1)) << np.uint64(48)) | uniq


def build_database(fasta: str, out_path: str, enzyme: str = "spcas9ngg", bin_width: int = 7) -> Dict[str, int]:
    """`index` end to end.  Returns a few statistics."""
    pack = pack_by_name(enzyme)
    contigs = read_fasta(fasta)
    vals, poss = sites_to_arrays(contigs, pack)
    return build_database_from_sites(out_path, pack, bin_width, vals, poss, [c[0] for c in contigs])


def build_database_from_sites(out_path: str, pack: ParameterPack, bin_width: int, vals: np.ndarray,
                              poss: np.ndarray, contig_names: List[str]) -> Dict[str, int]:
    uniq, counts, pos = collapse_sites(vals, poss)
    if pack.five_prime:
        # regroup into bin-major order, keeping positions attached
        key = db_order_key(pack, bin_width, uniq)
        order = np.argsort(key, kind="stable")
        pos_off = np.zeros(len(uniq) + 1, np.int64)
        np.cumsum(counts, out=pos_off[1:])
        pos = np.concatenate([pos[pos_off[i]:pos_off[i + 1]] for i in order]) if len(order) else pos
        uniq, counts = uniq[order], counts[order]
    blocks, ntargets = make_blocks(pack, bin_width, uniq, counts, pos)
    write_database(out_path, pack, bin_width, blocks, ntargets, contig_names)
    return {"sites": int(len(vals)), "targets": int(len(uniq)), "max_count": int(counts.max()) if len(counts) else 0,
            "indexed_bins": int(sum(1 for n in ntargets if n > 500 and not pack.five_prime))}


# ------------------------------------------------------------------------------------------------
# discover / score TSV : targetio/TabDelimitedHandler.scala:38-154, crispr/CRISPRHit.scala:54-101
@dataclass
class Guide:
    site: Site
    encoding: int


def guides_from_fasta(fasta: str, pack: ParameterPack, flank: int = 6, min_gc: float = 0.0, max_gc: float = 1.0,
                      ) -> List[Guide]:
    """OffTargetDiscovery.run :93-104: find sites, GC filter, encode with count 1, sort by start
    (ResultsAggregator.scala:35; ties keep discovery order here, the reference's quickSort leaves them unspecified)."""
    sites = find_target_sites(read_fasta(fasta), pack, flank)
    sites = [s for s in sites if min_gc <= gc_content(s.bases) <= max_gc]
    guides = [Guide(s, encode(s.bases, 1)) for s in sites]
    guides.sort(key=lambda g: g.site.position)
    return guides


def hit_token(pack: ParameterPack, target: int, mm: int, positions: Optional[np.ndarray], contigs: List[str],
              scores: Optional[str] = None) -> str:
    """CRISPRHit.toOutput :54-88."""
    bases, count = decode(int(target), pack.scan_len)
    tok = "%s_%d_%d" % (bases, count, mm)
    if positions is not None and len(positions):
        parts = []
        for p in positions:
            cid, start, _ln, fwd = pos_decode(int(p))
            parts.append("%s:%d^%s" % (contigs[cid - 1], start, "F" if fwd else "R"))
        tok += "<" + "|".join(parts) + ">"
    if scores:
        tok += scores
    return tok


def write_discover_tsv(path: str, pack: ParameterPack, guides: List[Guide], hits: Hits, contigs: List[str],
                       with_positions: bool = False, score_columns: Sequence[str] = (),
                       score_values: Optional[List[List[str]]] = None, write_ots: bool = True,
                       per_ot_scores: Optional[List[List[Optional[str]]]] = None) -> None:
    """TabDelimitedOutput (TabDelimitedHandler.scala:104-154)."""
    with open(path, "w") as fh:
        cols = ["contig", "start", "stop", "target", "context", "overflow", "orientation"] + list(score_columns)
        cols += ["otCount", "offTargets"] if write_ots else ["otCount"]
        fh.write("\t".join(cols) + "\n")
        for gi, g in enumerate(guides):
            s = g.site
            lo, hi = int(hits.row_ptr[gi]), int(hits.row_ptr[gi + 1])
            row = [s.contig, str(s.position), str(s.position + len(s.bases)), s.bases, s.context or "NONE",
                   "OVERFLOW" if hits.overflowed[gi] else "OK", "FWD" if s.forward else "RVS"]
            if score_values is not None:
                row += score_values[gi]
            cnts = (hits.targets[lo:hi] >> np.uint64(48)).astype(np.int64)
            row.append(str(int(cnts.sum())))
            if write_ots:
                toks = []
                for i in range(lo, hi):
                    pos = None
                    if with_positions and hits.positions is not None:
                        pos = hits.positions[int(hits.pos_ptr[i]):int(hits.pos_ptr[i + 1])]
                    sc = per_ot_scores[gi][i - lo] if per_ot_scores is not None else None
                    toks.append(hit_token(pack, int(hits.targets[i]), int(hits.mismatches[i]), pos, contigs, sc))
                row.append(",".join(toks))
            fh.write("\t".join(row) + "\n")


@dataclass
class TsvGuide:
    site: Site
    encoding: int
    overflow_budget: int
    inherited_overflow: bool
    targets: List[int]
    recorded_mm: List[int]
    positions: List[Optional[List[Tuple[str, int, bool]]]]
    annotations: Dict[str, str]


def read_discover_tsv(path: str, pack: ParameterPack, max_mismatch: int = 2 ** 31 - 1,
                      filter_overflow: bool = True) -> List[TsvGuide]:
    """TabDelimitedInput (TabDelimitedHandler.scala:169-334): parse guides + `SEQ_count_mm<...>` tokens back."""
    opener = gzip.open if path.endswith(".gz") else open
    out = []
    with opener(path, "rt") as fh:
        header = fh.readline().rstrip("\n").split("\t")
        assert header[:7] == ["contig", "start", "stop", "target", "context", "overflow", "orientation"]
        rest = header[7:]
        with_ots = rest[-2:] == ["otCount", "offTargets"]
        annotations = rest[:-2] if with_ots else rest[:-1]
        for ln in fh:
            sp = ln.rstrip("\n").split("\t")
            site = Site(sp[0], sp[3], sp[6] == "FWD", int(sp[1]), None if sp[4] == "NONE" else sp[4])
            is_ovf = sp[5] != "OK"
            ot_count = int(sp[7 + len(annotations)])
            budget = ot_count if is_ovf else ot_count + 1  # :242-249
            g = TsvGuide(site, encode(sp[3], 1), budget, is_ovf, [], [], [],
                         {a: sp[7 + i] for i, a in enumerate(annotations)})
            total = 0
            if with_ots and len(sp) == len(header) and sp[-1]:
                for tok in sp[-1].split(","):
                    tok = tok.split("{")[0]
                    f = tok.split("_")
                    seq, cnt = f[0], int(f[1])
                    mm = int(f[2].split("<")[0])
                    if mm > max_mismatch:  # :293
                        continue
                    pos = None
                    if "<" in tok:
                        pos = []
                        for pe in tok[tok.index("<") + 1:tok.index(">")].split("|"):
                            ctg, r = pe.split(":")
                            st, strand = r.split("^")
                            pos.append((ctg, int(st), strand == "F"))
                    if total < budget:  # if (!ot.full) ot.addOT :311,317
                        g.targets.append(encode(seq, cnt))
                        g.recorded_mm.append(mm)
                        g.positions.append(pos)
                        total += cnt
            full = total >= budget
            if (not filter_overflow) or (not is_ovf and not full):  # :259
                out.append(g)
    return out


def md5_file(path: str) -> str:
    return hashlib.md5(open(path, "rb").read()).hexdigest()
