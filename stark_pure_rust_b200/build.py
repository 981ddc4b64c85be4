"""Builds libstark_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m stark_pure_rust_b200.build [--force] [--verbose]

One object per translation unit, compiled in parallel, linked into stark_pure_rust_b200/libstark_b200.so.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libstark_b200.so")
CLI = os.path.join(HERE, "r1cs-stark")
SOURCES = ["api.cu", "ntt.cu", "ntt_b05.cu", "ntt_b6.cu", "ntt_b7.cu", "ntt_b8.cu", "merkle.cu", "poseidon.cu", "fri.cu", "prover.cu", "frontend.cu", "ext.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ARCH + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-O3,-Wall", "-c"] + os.environ.get("SB_NVCC_EXTRA", "").split()
LFLAGS = ARCH + ["-shared", "--cudart", "static", "-lpthread", "-ldl"]


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(HERE, "..", "include", "stark_b200.h"))
    return hs


def _compile(src, force, verbose):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    path = os.path.join(CSRC, src)
    if not force and os.path.exists(obj):
        t = os.path.getmtime(obj)
        if all(os.path.getmtime(d) <= t for d in [path] + _headers()):
            return obj, ""
    cmd = [NVCC] + CFLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", obj, path]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed on %s:\n%s%s" % (src, r.stdout, r.stderr))
    return obj, r.stdout + r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        res = list(ex.map(lambda s: _compile(s, force, verbose), SOURCES))
    objs = [o for o, _ in res]
    log = "".join(l for _, l in res)
    if verbose:
        sys.stderr.write(log)
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        r = subprocess.run([NVCC] + LFLAGS + ["-o", LIB] + objs, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    # the command-line front end (r1cs-stark/src/main.rs): plain C++ against the C ABI
    main_src = os.path.join(CSRC, "main.cpp")
    if force or not os.path.exists(CLI) or os.path.getmtime(CLI) < max(os.path.getmtime(main_src), os.path.getmtime(LIB)):
        r = subprocess.run(["g++", "-O2", "-std=c++17", "-o", CLI, main_src, "-L" + HERE, "-lstark_b200", "-Wl,-rpath,$ORIGIN"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("building r1cs-stark failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
