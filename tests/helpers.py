"""Seeded synthetic inputs shared by the parity tests (no reference files needed)."""
import os

import numpy as np

BASES = np.frombuffer(b"ACGT", np.uint8)


def random_genome(seed: int, n_bases: int, repeat_unit: int = 0, n_repeats: int = 0, n_contigs: int = 1):
    """Random ACGT contigs with some N runs and, optionally, a repeat family (a unit pasted many times) so that some
    targets carry large occurrence counts."""
    rng = np.random.default_rng(seed)
    contigs = []
    for c in range(n_contigs):
        seq = BASES[rng.integers(0, 4, n_bases)].copy()
        if repeat_unit and n_repeats:
            unit = BASES[rng.integers(0, 4, repeat_unit)]
            for _ in range(n_repeats):
                p = int(rng.integers(0, n_bases - repeat_unit))
                seq[p:p + repeat_unit] = unit
        for _ in range(3):  # a few N runs break windows
            p = int(rng.integers(0, n_bases - 50))
            seq[p:p + int(rng.integers(1, 40))] = ord("N")
        contigs.append(("ctg%d test" % (c + 1), seq.tobytes().decode()))
    return contigs


def write_fasta(path, contigs, width=60, lower_fraction=0.0, seed=0):
    rng = np.random.default_rng(seed)
    with open(path, "w") as fh:
        for name, seq in contigs:
            fh.write(">" + name + "\n")
            for i in range(0, len(seq), width):
                line = seq[i:i + width]
                if lower_fraction and rng.random() < lower_fraction:
                    line = line.lower()
                fh.write(line + "\n")


def random_guides(oracle, pack, seed: int, n: int):
    """20 uniform bases + uniform N + GG, de-duplicated on the protospacer, encoded with count 1."""
    rng = np.random.default_rng(seed)
    proto = pack.scan_len - pack.pam_len
    seen, out = set(), []
    while len(out) < n:
        s = "".join("ACGT"[i] for i in rng.integers(0, 4, proto + 1))
        if s[:proto] in seen:
            continue
        seen.add(s[:proto])
        out.append(oracle.encode(s + "GG", 1))
    return np.asarray(out, np.uint64)


def planted_guides(pack, targets: np.ndarray, seed: int, n: int, max_subs: int = 4):
    """Random database targets with 0..max_subs random substitutions in the protospacer (count reset to 1)."""
    rng = np.random.default_rng(seed)
    proto = pack.scan_len - pack.pam_len
    out = []
    for _ in range(n):
        t = int(targets[int(rng.integers(0, len(targets)))]) & 0xFFFFFFFFFFFF
        for _ in range(int(rng.integers(0, max_subs + 1))):
            pos = int(rng.integers(0, proto))
            shift = 2 * (pack.scan_len - 1 - pos)
            t ^= int(rng.integers(1, 4)) << shift
        out.append(t | (1 << 48))
    return np.asarray(out, np.uint64)


def assert_hits_equal(gpu, ref, check_positions=False):
    assert (np.asarray(gpu.row_ptr) == np.asarray(ref.row_ptr)).all(), "row_ptr differs"
    assert (np.asarray(gpu.targets) == np.asarray(ref.targets)).all(), "targets differ"
    assert (np.asarray(gpu.mismatches) == np.asarray(ref.mismatches)).all(), "mismatch counts differ"
    assert (np.asarray(gpu.total_count) == np.asarray(ref.total_count)).all(), "total_count differs"
    assert (np.asarray(gpu.overflowed) == np.asarray(ref.overflowed)).all(), "overflow flags differ"
    if check_positions:
        assert (np.asarray(gpu.pos_ptr) == np.asarray(ref.pos_ptr)).all(), "pos_ptr differs"
        assert (np.asarray(gpu.positions) == np.asarray(ref.positions)).all(), "positions differ"


def family_database(oracle, seed: int, n_seeds: int, variants_per_seed: int):
    """A spCas9-NGG database whose targets are variants of a few seed protospacers: 0..5 substitutions and, for two
    thirds of them, one deleted or inserted base (the 20-base window stays PAM-anchored).  Returns
    (targets u64 sorted with counts, seed guides u64)."""
    rng = np.random.default_rng(seed)
    seeds = ["".join("ACGT"[i] for i in rng.integers(0, 4, 20)) for _ in range(n_seeds)]
    vals = set()
    for s in seeds:
        for _ in range(variants_per_seed):
            t = list(s)
            kind = int(rng.integers(0, 3))
            q = int(rng.integers(1, 19))
            if kind == 1:    # a guide base has no partner: delete it, a new base enters at the 5' end
                t = ["ACGT"[int(rng.integers(0, 4))]] + t[:q] + t[q + 1:]
            elif kind == 2:  # an extra genomic base: insert one, the 5'-most base leaves the window
                t = t[1:q + 1] + ["ACGT"[int(rng.integers(0, 4))]] + t[q + 1:]
            for _ in range(int(rng.integers(0, 6))):
                t[int(rng.integers(0, 20))] = "ACGT"[int(rng.integers(0, 4))]
            assert len(t) == 20
            vals.add(oracle.encode("".join(t) + "ACGT"[int(rng.integers(0, 4))] + "GG", 1) & 0xFFFFFFFFFFFF)
    v = np.asarray(sorted(vals), np.uint64)
    counts = rng.integers(1, 4, len(v)).astype(np.uint64)
    big = rng.random(len(v)) < 0.002
    counts[big] = rng.integers(500, 3000, int(big.sum())).astype(np.uint64)
    targets = v | (counts << np.uint64(48))
    guides = np.asarray([oracle.encode(s + "TGG", 1) for s in seeds], np.uint64)
    return targets, guides
