"""Build libnaf_b200.so in-tree with nvcc for sm_100a (no torch headers, pure C ABI).

    python -m naf_b200.csrc.build [--force] [--verbose]
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "libnaf_b200.so")
STAMP = os.path.join(HERE, ".build_stamp")
SOURCES = ["naf_abi.cu", "naf_pack.cu", "naf_kpool.cu", "naf_xattn_generic.cu", "naf_xattn_simt.cu",
           "naf_xattn_tc.cu", "naf_xattn_tcws.cu", "naf_encoder.cu", "naf_conv_tc.cu"]
HEADERS = ["naf_common.cuh", "naf_umma.cuh", os.path.join(ROOT, "include", "naf_b200.h")]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _digest() -> str:
    h = hashlib.sha256()
    for f in SOURCES + HEADERS + [os.path.abspath(__file__)]:
        path = f if os.path.isabs(f) else os.path.join(HERE, f)
        with open(path, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def _flags(verbose: bool = False) -> list[str]:
    flags = ["-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
             "-Xcompiler", "-fPIC,-O3,-fvisibility=hidden",
             "-I", os.path.join(ROOT, "include"), "-I", HERE, "-DNAF_BUILDING_LIB", "-DNAF_WITH_TC"]
    if verbose:
        flags += ["-Xptxas", "-v"]
    return flags


def command(verbose: bool = False) -> list[str]:
    """The equivalent single nvcc invocation (what build() does, file by file in parallel)."""
    return [nvcc_path()] + _flags(verbose) + ["-shared", "-o", LIB] + [os.path.join(HERE, s) for s in SOURCES]


def build(force: bool = False, verbose: bool = False) -> str:
    dig = _digest()
    if not force and os.path.isfile(LIB) and os.path.isfile(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == dig:
                return LIB
    from concurrent.futures import ThreadPoolExecutor

    objdir = os.path.join(HERE, "_obj")
    os.makedirs(objdir, exist_ok=True)
    nvcc = nvcc_path()

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        res = subprocess.run([nvcc] + _flags(verbose) + ["-c", os.path.join(HERE, src), "-o", obj],
                             capture_output=True, text=True)
        return src, obj, res

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    failed = False
    for src, _, res in results:
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
        failed |= res.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libnaf_b200.so")
    res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] +
                         [obj for _, obj, _ in results], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libnaf_b200.so")
    with open(STAMP, "w") as fh:
        fh.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
