"""The JNI shim (integration/jni/flashfry_b200_jni.c) EXECUTED: every Java_flashfry_NativeBridge_* function is called
through a functional mock JNIEnv (tests/stubs/jni_mock.c -- no JDK in this image) against the real library on the GPU,
and the results are compared with direct calls of the C ABI (tests/stubs/jni_exec.c).  Also: the CLI's TSV fast path
(ff_hits_write_tsv) against the object-building Traverser drop-in of the host mirror."""
import os
import subprocess

import pytest

import helpers

pytestmark = pytest.mark.gpu
ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def test_jni_shim_executes_against_the_gpu(small_db, tmp_path):
    import torch
    exe = str(tmp_path / "jni_exec")
    libdir = os.path.join(ROOT, "flashfry_b200")
    r = subprocess.run(["gcc", "-O1", "-Wall", "-Wno-unused-function", "-I", os.path.join(ROOT, "tests", "stubs"), "-I", os.path.join(ROOT, "include"),
                        "-o", exe, os.path.join(ROOT, "tests", "stubs", "jni_exec.c"), os.path.join(ROOT, "tests", "stubs", "jni_mock.c"),
                        "-L", libdir, "-lflashfry_b200", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, small_db[0], str(tmp_path), str(torch.cuda.device_count())], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "JNI_EXEC_OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("positions", [False, True])
def test_tsv_fast_path_equals_the_object_building_traverser(small_db, oracle, tmp_path, positions):
    """`flashfry_b200_cli discover` writes its TSV with ff_hits_write_tsv straight from the CSR; host_selftest traverse
    goes the reference's way (GpuTraverser.scan -> a CRISPRHit per hit -> TabDelimitedOutput.write).  Same bytes."""
    cli = os.path.join(ROOT, "flashfry_b200", "flashfry_b200_cli")
    selftest = os.path.join(ROOT, "flashfry_b200", "host_selftest")
    contigs = helpers.random_genome(101, 200_000, repeat_unit=60, n_repeats=400, n_contigs=2)  # the genome small_db indexes
    fa = str(tmp_path / "guides.fa")
    helpers.write_fasta(fa, [("region", contigs[0][1][5000:5600])])
    a, b = str(tmp_path / "fast.tsv"), str(tmp_path / "objects.tsv")
    cmd = [cli, "discover", "--database", small_db[0], "--fasta", fa, "--output", a, "--maximumOffTargets", "40"] + (["--positionOutput"] if positions else [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([selftest, "traverse", small_db[0], fa, b, "positions" if positions else "nopos", "4", "40"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ta, tb = open(a).read(), open(b).read()
    assert ta == tb and ta.count("\n") > 20 and "OVERFLOW" in ta and "OK" in ta
