"""CPU-only checks of the product's host side: the shared library loads and exports the whole ABI, fails loudly
without a GPU, the CLI mirrors FlashFry's option surface, and the guide-sharding logic works over gloo (world 2)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

from conftest import ROOT

LIB = os.path.join(ROOT, "flashfry_b200", "libflashfry_b200.so")
CLI = os.path.join(ROOT, "flashfry_b200", "flashfry_b200_cli")


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(LIB):
        from flashfry_b200 import build
        build.build()
    return LIB


def test_library_exports_every_symbol_of_the_header(built):
    hdr = open(os.path.join(ROOT, "include", "flashfry_b200.h")).read()
    declared = set(re.findall(r"\b(ff_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 17
    lib = ctypes.CDLL(built)
    for name in sorted(declared):
        assert hasattr(lib, name), "missing export: " + name
    from flashfry_b200 import _native
    assert set(_native.SYMBOLS) == declared
    assert lib.ff_abi_version() == 2


def test_no_gpu_means_loud_failure_not_a_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import flashfry_b200.api as ff
    with pytest.raises(ff.FlashFryError) as e:
        ff.Context(0)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: no product source may include, import, link or call it."""
    bad = re.compile(r"#\s*include[^\n]*oracle|^\s*(from|import)\s+oracle|ff_oracle|ffo_|libff_oracle", re.M)
    for base, _dirs, files in os.walk(os.path.join(ROOT, "flashfry_b200")):
        if "_obj" in base or "__pycache__" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".inl")):
                txt = open(os.path.join(base, f), errors="replace").read()
                assert not bad.search(txt), os.path.join(base, f)
    assert not bad.search(open(os.path.join(ROOT, "include", "flashfry_b200.h")).read())
    ldd = subprocess.run(["ldd", LIB], capture_output=True, text=True).stdout if os.path.exists(LIB) else ""
    assert "oracle" not in ldd


def test_cli_surface(built, tmp_path):
    if not os.path.exists(CLI):
        pytest.skip("CLI not built")
    r = subprocess.run([CLI], capture_output=True, text=True)
    assert r.returncode == 2 and "discover" in r.stderr and "score" in r.stderr
    r = subprocess.run([CLI, "discover", "--fasta", "x.fa"], capture_output=True, text=True)
    assert r.returncode == 1 and "Missing required option '--database'" in r.stderr
    r = subprocess.run([CLI, "index"], capture_output=True, text=True)
    assert r.returncode == 2 and "outside the GPU hot path" in r.stderr
    # header validation happens before the GPU is touched and carries the reference's message (BinaryHeader.scala:121-124)
    (tmp_path / "db.header").write_text("42\n1\n3\n16384\n")
    r = subprocess.run([CLI, "discover", "-fasta", "x.fa", "-database", str(tmp_path / "db"), "-output", str(tmp_path / "o")],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "magic number" in r.stderr


def test_shard_range_partitions_exactly():
    from flashfry_b200.sharding import shard_range
    for n in (0, 1, 7, 8, 100000, 100003):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                assert 0 <= lo <= hi <= n
                cover += list(range(lo, hi)) if n < 1000 else [(lo, hi)]
            if n < 1000:
                assert cover == list(range(n))
            else:
                assert cover[0][0] == 0 and cover[-1][1] == n and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
                assert max(b - a for a, b in cover) - min(b - a for a, b in cover) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from flashfry_b200.sharding import shard_range, all_gather_counts
from oracle import ff_oracle as o
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# a tiny sorted database and guide set, identical on every rank (one index replica per rank)
rng = np.random.default_rng(3)
pack = o.PACK_BY_INDEX[3]
t = np.unique(rng.integers(0, 1 << 42, 30000, dtype=np.uint64)) << np.uint64(4) | np.uint64(0xA) | (np.uint64(1) << np.uint64(48))
guides = (t[rng.integers(0, len(t), 41)] & np.uint64(0xFFFFFFFFFFFF)) ^ (np.uint64(3) << np.uint64(20)) | (np.uint64(1) << np.uint64(48))
bo = o.bin_offsets_from_sorted(pack, 7, t)
full = o.discover_soa(pack, 7, t, bo, guides, 3, 2000)
lo, hi = shard_range(len(guides), rank, world)
mine = o.discover_soa(pack, 7, t, bo, guides[lo:hi], 3, 2000)      # stands in for the per-rank GPU call
allc = all_gather_counts(torch.from_numpy(mine.total_count.astype(np.int32)), len(guides))
assert allc.numpy().tolist() == full.total_count.tolist(), (rank, allc, full.total_count)
assert (mine.targets == full.targets[full.row_ptr[lo]:full.row_ptr[hi]]).all()
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_guide_sharding_world2_gloo(tmp_path):
    """N>1 path on CPU: two ranks shard the guides, each runs its shard, the all-gathered count vector equals the
    single-process result (the oracle stands in for the per-rank GPU call; tests may use it as the checker)."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(script), ROOT], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == 2


SELFTEST = os.path.join(ROOT, "flashfry_b200", "host_selftest")


def test_host_mirror_tsv_round_trips(built, tmp_path):
    """The C++ TabDelimitedInput/Output pair re-writes the reference's fixtures byte for byte
    (TabDelimitedHanderTest.scala:40-51 on fake.sites; the md5-pinned EMX1 files)."""
    import gzip
    from conftest import GOLDEN
    if not os.path.exists(SELFTEST):
        pytest.skip("host_selftest not built")
    fake = tmp_path / "fake.sites"
    fake.write_bytes(gzip.open(os.path.join(GOLDEN, "fake.sites.gz")).read())
    cases = [(2, str(fake), True), (3, os.path.join(GOLDEN, "EMX1.output"), False),
             (3, os.path.join(GOLDEN, "EMX1.output.with_positions"), True),
             (3, os.path.join(GOLDEN, "EMX1.output.scored_with_ots"), True)]
    for enzyme, path, positions in cases:
        out = tmp_path / "rt.tsv"
        cmd = [SELFTEST, "roundtrip", str(enzyme), path, str(out)] + (["positions"] if positions else [])
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert out.read_bytes() == open(path, "rb").read(), path


def test_host_mirror_site_finder_and_double_format(built, oracle, tmp_path):
    """SimpleSiteFinder in C++ == the oracle's restatement on a random multi-contig FASTA (all six enzymes), and
    javaDoubleToString == the oracle's Java formatting."""
    import helpers
    import numpy as np
    if not os.path.exists(SELFTEST):
        pytest.skip("host_selftest not built")
    contigs = helpers.random_genome(12, 5000, n_contigs=3)
    fa = tmp_path / "g.fa"
    helpers.write_fasta(str(fa), contigs, lower_fraction=0.3)
    for pack in oracle.PACKS.values():
        r = subprocess.run([SELFTEST, "sites", str(pack.index), str(fa), "6"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        got = [ln.split("\t") for ln in r.stdout.strip().split("\n") if ln]
        ref = oracle.find_target_sites(oracle.read_fasta(str(fa)), pack, 6)
        assert len(got) == len(ref) and len(ref) > 10, pack.name
        for g, s in zip(got, ref):
            assert g == [s.contig, str(s.position), s.bases, "FWD" if s.forward else "RVS", s.context or "NONE"]
    rng = np.random.default_rng(1)
    xs = [0.0, 1.0, 100.0, 1e7, 9999999.999, 1e-3, 9.99e-4, 0.023, 0.16666666680238093, 98.41774847095192, 1.5e-10, 2e22, 123456789.125]
    xs += list(rng.random(200)) + list(rng.random(50) * 1e-6) + list(rng.random(50) * 1e9)
    r = subprocess.run([SELFTEST, "double"] + [repr(float(x)) for x in xs], capture_output=True, text=True)
    assert r.stdout.strip().split("\n") == [oracle.java_double_str(float(x)) for x in xs]


def test_host_mirror_site_finder_reference_vectors(built, tmp_path):
    """SimpleSiteFinderTest.scala known answers through the C++ SimpleSiteFinder."""
    import json
    from conftest import GOLDEN
    if not os.path.exists(SELFTEST):
        pytest.skip("host_selftest not built")
    vec = json.load(open(os.path.join(GOLDEN, "reference_unit_vectors.json")))
    idx = {"CPF1": 1, "SPCAS9": 2, "SPCAS9NGG": 3, "SPCAS9NAG": 4, "SPCAS919": 5, "SPCAS9NGG19": 6}
    comp = str.maketrans("ACGT", "TGCA")
    for c in vec["site_finder_cases"]:
        fa = tmp_path / "c.fa"
        fa.write_text(">testContig\n" + c["seq"] + "\n")
        r = subprocess.run([SELFTEST, "sites", str(idx[c["pack"]]), str(fa), str(c["flank"])], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        got = [ln.split("\t") for ln in r.stdout.strip().split("\n") if ln]
        assert len(got) == len(c["sites"]), c
        for g, (kind, a, b) in zip(got, c["sites"]):
            want = c["seq"][a:b] if kind == "fwd" else c["seq"][a:b].translate(comp)[::-1]
            assert g[2] == want and int(g[1]) == a and g[3] == ("FWD" if kind == "fwd" else "RVS")
        if "context_defined" in c:
            assert [g[4] != "NONE" for g in got] == c["context_defined"]


def test_integration_sources_match_the_document_and_the_jni_shim_type_checks():
    """INTEGRATION.md's Java / C / Scala blocks are shipped as files under integration/; the JNI shim must compile
    (syntax + types) against the C ABI header with a minimal stand-in for <jni.h> (this image has no JDK), and every
    JNI entry point must have its `native` declaration in NativeBridge.java."""
    import subprocess
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```(java|c|scala)\n(.*?)```", md, re.S)
    files = {"java": ["integration/java/flashfry/NativeBridge.java"], "c": ["integration/jni/flashfry_b200_jni.c"],
             "scala": ["integration/scala/GpuTraverser.scala", "integration/scala/GpuScoreModel.scala"]}
    seen = {"java": 0, "c": 0, "scala": 0}
    for lang, body in blocks:
        path = os.path.join(ROOT, files[lang][seen[lang]])
        seen[lang] += 1
        assert open(path).read().endswith(body), path + " drifted from INTEGRATION.md"
    assert seen == {"java": 1, "c": 1, "scala": 2}
    shim = os.path.join(ROOT, "integration", "jni", "flashfry_b200_jni.c")
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    r = subprocess.run([gcc, "-std=c11", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-Wno-unused-parameter",
                        "-I", os.path.join(ROOT, "tests", "stubs"), "-I", os.path.join(ROOT, "include"), shim],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    java = open(os.path.join(ROOT, "integration", "java", "flashfry", "NativeBridge.java")).read()
    natives = set(re.findall(r"public static native [\w\[\].]+\s+(\w+)\(", java))
    exported = set(re.findall(r"Java_flashfry_NativeBridge_(\w+)\(", open(shim).read()))
    assert natives == exported and len(natives) >= 10


def test_library_carries_sm_100a_code_for_every_hot_kernel(built):
    """The product is hand-written CUDA compiled for sm_100a only: the shared object must hold sm_100a cubins with the
    scan / ordering / scoring kernels (no PTX-JIT fallback, no other architectures)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    elfs = subprocess.run([cuobjdump, "--list-elf", built], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", elfs))
    assert archs == {"sm_100a"}, archs
    syms = subprocess.run([cuobjdump, "-symbols", built], capture_output=True, text=True).stdout
    for kernel in ("k_bin_scan", "k_pair_scan", "k_slice_planes", "k_seed_scan", "k_pattern_scan", "k_overflow_cut", "k_cut_window", "k_sort_cut", "k_sort_long", "k_compact_rows", "k_gather",
                   "k_score", "k_hit_aggregates", "k_cell_offsets", "k_pair_scan2", "k_guide_place",
                   "k_peer_barrier", "k_peer_counts", "k_peer_compact", "k_peer_totals"):
        assert kernel in syms, "kernel missing from the library: " + kernel
    # the bin scan stages its bins with TMA bulk copies (cp.async.bulk -> UBLKCP in SASS) completing on an mbarrier
    sass = subprocess.run([cuobjdump, "-sass", "-fun", "_ZN2ff10k_bin_scanILi9EEEvNS_9BinParamsE", built], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass and "SYNCS" in sass, "k_bin_scan lost its TMA bulk copy / mbarrier"
    # part two's ring: TMA bulk copies into two buffers, no CTA-wide barrier inside its bin loop (one BAR at the start only)
    ring = subprocess.run([cuobjdump, "-sass", "-fun", "_ZN2ff12k_pair_scan2ILi11EEEvNS_10PairParamsE", built], capture_output=True, text=True).stdout
    assert "UBLKCP" in ring and "SYNCS" in ring, "k_pair_scan2 lost its TMA ring"
    assert len(re.findall(r"\bBAR\.SYNC", ring)) <= 2, "k_pair_scan2 grew CTA-wide barriers"
    # the database-sharded drain pushes candidates to peers with plain stores: the only system-scope atomic is the barrier's
    bar = subprocess.run([cuobjdump, "-sass", "-fun", "_ZN2ff14k_peer_barrierENS_8PeerPtrsEiijPjj", built], capture_output=True, text=True).stdout
    assert "ATOM" in bar or "RED" in bar, "k_peer_barrier lost its arrival atomic"


def test_c_abi_shard_range_is_the_one_definition_of_sharding(built):
    """ff_shard_range (pure function of the C ABI, no GPU needed) is what ff_multi_discover shards with; the
    one-process-per-GPU ranks (flashfry_b200.sharding) call the same symbol."""
    import flashfry_b200.api as api
    from flashfry_b200.sharding import shard_range
    for n in (0, 1, 5, 100000, 100003):
        for world in (1, 2, 8):
            edges = [api.shard_range(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and sum(c for _, c in edges) == n
            assert all(edges[r][0] + edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            assert [shard_range(n, r, world) for r in range(world)] == [(f, f + c) for f, c in edges]
    assert api.shard_range(10, 0, 0) == (0, 0) and api.shard_range(10, 2, 5) == (0, 0)
