"""Build of the device library (nip_b200/libnipgpu.so) and the host glue.

Everything is compiled in-tree with explicit nvcc / gcc command lines so that the
built `.so` files travel with the repository snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libnipgpu.so")
HOST_LIB = os.path.join(PKG, "libnip_gpu_backend.so")
REF_SRC = "/root/reference/src"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "--expt-relaxed-constexpr",
    "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
]
SOURCES = ["api.cu", "jtree.cu", "chain.cu", "dense.cu", "params.cu", "probe.cu", "group.cu", "factor.cu", "model.cpp"]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_device_library(force=False, verbose=False, extra_flags=()):
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "nipgpu.h")]
    if not force and not _stale(LIB, deps):
        return LIB
    objs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, os.path.splitext(src)[0] + ".o")
        if force or _stale(obj, deps):
            cmd = [_nvcc(), *NVCC_FLAGS, *extra_flags, "-c", os.path.join(CSRC, src), "-o", obj]
            if verbose:
                print(" ".join(cmd))
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("nvcc failed on " + src)
            if verbose and (r.stdout or r.stderr):
                print(r.stdout, r.stderr)
        objs.append(obj)
    cmd = [_nvcc(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-ldl", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    return LIB


def build_host_backend(force=False, verbose=False):
    """nip.h entry points on top of the C ABI; needs the reference headers
    (struct layouts), so it is only (re)built where /root/reference exists."""
    src = [os.path.join(PKG, "host", "nip_gpu_backend.c"), os.path.join(PKG, "host", "nip_model_export.c"),
           os.path.join(PKG, "host", "nip_data_bin.c"), os.path.join(PKG, "host", "nip_model_write.c")]
    if not os.path.isdir(REF_SRC) or not all(os.path.exists(s) for s in src):
        return HOST_LIB if os.path.exists(HOST_LIB) else None
    if not force and not _stale(HOST_LIB, src + [LIB]):
        return HOST_LIB
    cmd = ["gcc", "-std=gnu99", "-O2", "-fPIC", "-shared", "-Wall", "-Wno-unused",
           "-include", os.path.join(PKG, "host", "nip_errcodes_shim.h"),
           "-I" + REF_SRC, "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(PKG, "host"),
           "-o", HOST_LIB, *src, "-L" + PKG, "-lnipgpu", "-Wl,-rpath,$ORIGIN", "-Wl,-Bsymbolic-functions", "-lm"]
    if verbose:
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("host backend build failed")
    return HOST_LIB


if __name__ == "__main__":
    build_device_library(force="--force" in sys.argv, verbose=True)
    build_host_backend(force="--force" in sys.argv, verbose=True)
    print("built", LIB)
