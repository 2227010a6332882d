"""TEST INFRASTRUCTURE (oracle side): stream a large synthetic database out in FlashFry's own on-disk format.

What it restates (src/main/scala/...): DatabaseWriter.writeToBinnedFileSet reference/binary/DatabaseWriter.scala:58-111
(one block per 7-mer bin in AAAAAAA..TTTTTTT order, indexed when the bin holds > 500 targets), createIndexedBlock /
createLinearBlock reference/binary/blocks/BlockManager.scala:362-442, BinaryHeader.writeHeader
reference/binary/BinaryHeader.scala:69-97, over the BGZF container of htsjdk 2.8.1 (SAM spec 4.1; 65 498-byte members).
ff_oracle.write_database does the same for small databases through Python objects per bin; this module handles 3e8
targets: numpy per run of bins, members deflated on all host threads (zlib releases the GIL), virtual pointers
computed from the recorded member sizes.  Positions are synthetic and a pure function of the occurrence's global
ordinal, so a loader can be checked without keeping 3 GB of them around:
    position(o) = start (o mod 2^31) | contig (1 + (o >> 31) mod 24) << 32 | 23 << 52 | strand (o & 1) << 60
Only tests/, bench.py's cold-start measurement and tools/ import this.
"""
from __future__ import annotations

import struct
import zlib
from concurrent.futures import ThreadPoolExecutor
from typing import Dict

import numpy as np

from . import ff_oracle as o

MEMBER = o.BGZF_BLOCK


def synthetic_positions(first_ordinal: int, n: int) -> np.ndarray:
    od = np.arange(first_ordinal, first_ordinal + n, dtype=np.uint64)
    return ((od & np.uint64(0x7FFFFFFF)) | ((np.uint64(1) + ((od >> np.uint64(31)) % np.uint64(24))) << np.uint64(32)) |
            (np.uint64(23) << np.uint64(52)) | ((od & np.uint64(1)) << np.uint64(60)))


def _member(data: bytes, level: int) -> bytes:
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    comp = co.compress(data) + co.flush()
    if len(comp) + 26 > 65536:  # incompressible: store (deflate "stored" blocks always fit a 65 498-byte payload)
        co = zlib.compressobj(0, zlib.DEFLATED, -15)
        comp = co.compress(data) + co.flush()
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(comp) + 25) + comp +
            struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


def write_big_database(path: str, pack: o.ParameterPack, targets: np.ndarray, bin_width: int = 7, threads: int = 16,
                       level: int = 1, bins_per_run: int = 128) -> Dict[str, float]:
    """targets: u64 longs with counts, in database order (3'-PAM packs: sorted).  Returns statistics."""
    assert not pack.five_prime, "bin-major reordering of 5'-PAM packs is not needed for the benchmark databases"
    n_bins = 4 ** bin_width
    seq = targets & np.uint64(0xFFFFFFFFFFFF)
    counts = (targets >> np.uint64(48)).astype(np.int64)
    bshift = np.uint64(2 * (pack.scan_len - bin_width))
    tb = np.searchsorted((seq >> bshift).astype(np.int64), np.arange(n_bins + 1), side="left")
    sub_shift = np.uint64(2 * (pack.scan_len - (bin_width + 4)))
    pos_off = np.zeros(len(targets) + 1, np.int64)
    np.cumsum(counts, out=pos_off[1:])
    bin_stream_off = np.zeros(n_bins, np.int64)
    bin_bytes = np.zeros(n_bins, np.int64)
    member_sizes = []
    carry = b""
    stream_off = 0
    indexed = 0
    with open(path, "wb") as fh, ThreadPoolExecutor(max_workers=threads) as pool:
        for b0 in range(0, n_bins, bins_per_run):
            b1 = min(n_bins, b0 + bins_per_run)
            parts = []
            for b in range(b0, b1):
                lo, hi = int(tb[b]), int(tb[b + 1])
                n = hi - lo
                plo, phi = int(pos_off[lo]), int(pos_off[hi])
                body = np.empty(n + (phi - plo), np.uint64)
                slot = np.arange(n, dtype=np.int64) + (pos_off[lo:hi] - plo)
                is_t = np.zeros(len(body), bool)
                is_t[slot] = True
                body[is_t] = targets[lo:hi]
                body[~is_t] = synthetic_positions(plo, phi - plo)
                if n > 500:  # DatabaseWriter.scala:85 -> createIndexedBlock(..., 4)
                    sub = ((seq[lo:hi] >> sub_shift) & np.uint64(255)).astype(np.int64)
                    sizes = np.bincount(sub, weights=(1 + counts[lo:hi]), minlength=256).astype(np.int64)
                    firsts = np.full(256, -1, np.int64)
                    present, first_idx = np.unique(sub, return_index=True)
                    firsts[present] = slot[first_idx]
                    table = (firsts.astype(np.uint64) << np.uint64(32)) | sizes.astype(np.uint64)
                    blk = np.concatenate([np.asarray([2], np.uint64), table, body])
                    indexed += 1
                else:
                    blk = np.concatenate([np.asarray([1], np.uint64), body])
                bin_stream_off[b] = stream_off
                bin_bytes[b] = len(blk) * 8
                stream_off += len(blk) * 8
                parts.append(blk.astype("<u8").tobytes())
            data = carry + b"".join(parts)
            n_full = len(data) // MEMBER if b1 < n_bins else (len(data) + MEMBER - 1) // MEMBER
            chunks = [data[i * MEMBER:(i + 1) * MEMBER] for i in range(n_full)]
            carry = data[n_full * MEMBER:] if b1 < n_bins else b""
            for m in pool.map(lambda c: _member(c, level), chunks):
                fh.write(m)
                member_sizes.append(len(m))
        fh.write(o.BGZF_EOF)
    member_off = np.zeros(len(member_sizes) + 1, np.int64)
    np.cumsum(np.asarray(member_sizes, np.int64), out=member_off[1:])
    offs = []
    for b in range(n_bins):
        m, inside = divmod(int(bin_stream_off[b]), MEMBER)
        offs.append(((int(member_off[m]) << 16) | inside, int(bin_bytes[b]), int(tb[b + 1] - tb[b])))
    o.write_header(path + ".header", pack, bin_width, offs, ["chrSynth%d" % (i + 1) for i in range(24)])
    return {"targets": int(len(targets)), "positions": int(pos_off[-1]), "inflated_bytes": int(stream_off),
            "file_bytes": int(member_off[-1] + len(o.BGZF_EOF)), "indexed_bins": indexed, "members": len(member_sizes)}
