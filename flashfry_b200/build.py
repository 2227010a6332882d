"""In-tree build of libflashfry_b200.so (CUDA, sm_100a only) and the host CLI.

    python -m flashfry_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box with the tree.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libflashfry_b200.so")
CLI = os.path.join(HERE, "flashfry_b200_cli")
SELFTEST = os.path.join(HERE, "host_selftest")

CU_SOURCES = ["ff_api.cu", "ff_db.cu", "ff_discover.cu", "ff_score.cu", "ff_multi.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-O2,-Wall", "--expt-relaxed-constexpr"]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _newer(src_files, target) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(f) > t for f in src_files)


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".hpp", ".inl"))]
    hs.append(os.path.join(HERE, "..", "include", "flashfry_b200.h"))
    host = os.path.join(CSRC, "host")
    if os.path.isdir(host):
        hs += [os.path.join(host, f) for f in os.listdir(host) if f.endswith((".hpp", ".h"))]
    return hs


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed:\n" + " ".join(cmd) + "\n" + r.stdout)
    if verbose and r.stdout.strip():
        print(r.stdout)
    return r.stdout


def build_variant(name: str, defines) -> str:
    """Experiment support: compile the library with extra -D flags into gpurun_variants/<name>/ (git-ignored, ships to
    the GPU box); select it at run time with FLASHFRY_B200_LIB=<path>."""
    out_dir = os.path.join(HERE, "..", "gpurun_variants", name)
    os.makedirs(out_dir, exist_ok=True)
    cc = nvcc()
    objs = []
    jobs = []
    for src in CU_SOURCES:
        o = os.path.join(out_dir, src.replace(".cu", ".o"))
        objs.append(o)
        jobs.append([cc] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-c", os.path.join(CSRC, src), "-o", o])
    with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
        list(ex.map(lambda c: _run(c, False), jobs))
    lib = os.path.join(out_dir, "libflashfry_b200.so")
    _run([cc, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lz", "-lpthread", "-ldl",
                                               "-Xcompiler", "-fPIC", "-cudart", "static"], False)
    for o in objs:
        os.remove(o)
    return lib


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    cc = nvcc()
    hdrs = _headers()
    jobs = []
    for src in CU_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or _newer([s] + hdrs, o):
            jobs.append([cc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o])
    if jobs:
        with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
            list(ex.map(lambda c: _run(c, verbose), jobs))
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in CU_SOURCES]
    if force or _newer(objs, LIB):
        _run([cc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lz", "-lpthread", "-ldl",
                                                   "-Xcompiler", "-fPIC", "-cudart", "static"], verbose)
    cli_src = os.path.join(CSRC, "host", "flashfry_cli.cpp")
    if os.path.exists(cli_src) and (force or _newer([cli_src] + hdrs + [LIB], CLI)):
        gxx = shutil.which("g++") or "g++"
        if os.path.exists("/usr/bin/g++"):
            gxx = "/usr/bin/g++"
        _run([gxx, "-O2", "-std=c++17", "-Wall", "-o", CLI, cli_src, "-I", os.path.join(HERE, "..", "include"),
              "-L", HERE, "-lflashfry_b200", "-Wl,-rpath,$ORIGIN", "-lz", "-lpthread"], verbose)
    st_src = os.path.join(CSRC, "host", "host_selftest.cpp")
    if os.path.exists(st_src) and (force or _newer([st_src] + hdrs + [LIB], SELFTEST)):
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++")
        _run([gxx, "-O2", "-std=c++17", "-Wall", "-o", SELFTEST, st_src, "-I", os.path.join(HERE, "..", "include"),
              "-L", HERE, "-lflashfry_b200", "-Wl,-rpath,$ORIGIN", "-lz", "-lpthread"], verbose)
    return LIB


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--variant":  # python -m flashfry_b200.build --variant NAME FF_X=1 FF_Y=2
        print(build_variant(sys.argv[2], sys.argv[3:]))
    else:
        print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
