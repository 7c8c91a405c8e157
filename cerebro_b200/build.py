"""Builds the in-tree C-ABI library ``cerebro_b200/_native/libcerebro_b200.so`` with nvcc for
sm_100a.  Cross-compiles without a GPU.  ``python -m cerebro_b200.build [--force]``."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_native")
LIB = os.path.join(OUT_DIR, "libcerebro_b200.so")
HARNESS = os.path.join(OUT_DIR, "cerebro_harness")
STEREO_EMUL = os.path.join(OUT_DIR, "libstereo_emul.so")
ORB_EMUL = os.path.join(OUT_DIR, "liborb_emul.so")
SOURCES = ["capi.cu", "comm.cu", "search.cu", "pnp.cu", "netvlad.cu", "frontend.cu", "features.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-fvisibility=hidden",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found; cannot build libcerebro_b200.so")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "cerebro_b200.h"))
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OUT_DIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or r.returncode:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError("nvcc failed for %s" % src)
    if force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    # ROS-free C++ host shim + harness over the C ABI (plain g++, no CUDA headers needed)
    host = os.path.join(HERE, "host")
    hsrc = [os.path.join(host, "harness.cpp"), os.path.join(host, "cerebro_shim.hpp"), os.path.join(HERE, "..", "include", "cerebro_b200.h")]
    if force or _stale(HARNESS, hsrc + [LIB]):
        cmd = ["g++", "-O2", "-std=c++17", "-Wall", hsrc[0], "-o", HARNESS, "-L" + OUT_DIR, "-lcerebro_b200",
               "-Wl,-rpath,$ORIGIN", "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("host harness build failed")
    # CPU emulation of the stereo kernels' thread space over the shared per-thread bodies (tests only, never a fallback)
    esrc = [os.path.join(host, "stereo_emul.cpp"), os.path.join(CSRC, "stereo_core.h")]
    if force or _stale(STEREO_EMUL, esrc):
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", esrc[0], "-o", STEREO_EMUL]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("stereo emulation build failed")
    # same for the ORB / remap kernels (separate roundings everywhere: -ffp-contract=off)
    osrc = [os.path.join(host, "orb_emul.cpp"), os.path.join(CSRC, "orb_core.h"), os.path.join(CSRC, "orb_pipeline.h"), os.path.join(CSRC, "orb_pattern.h")]
    if force or _stale(ORB_EMUL, osrc):
        cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", osrc[0], "-o", ORB_EMUL]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("ORB emulation build failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
