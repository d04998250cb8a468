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
           "naf_xattn_tcws.cu", "naf_xattn_tma.cu", "naf_xattn_union.cu", "naf_xattn_bwd.cu", "naf_xattn_bwd_tc.cu", "naf_encoder.cu", "naf_conv_tc.cu"]
HEADERS = ["naf_common.cuh", "naf_umma.cuh", "naf_tmap.cuh", os.path.join(ROOT, "include", "naf_b200.h")]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _read(f: str) -> bytes:
    path = f if os.path.isabs(f) else os.path.join(HERE, f)
    with open(path, "rb") as fh:
        return fh.read()


def _digest(sources=None, verbose: bool = False) -> str:
    """Hash of what an object (one source) or the library (all sources) is built from: the source(s),
    every header, this script and the compiler flags."""
    h = hashlib.sha256()
    for f in list(SOURCES if sources is None else sources) + HEADERS + [os.path.abspath(__file__)]:
        h.update(_read(f))
    h.update(" ".join(_flags(verbose) + EXTRA_FLAGS).encode())
    return h.hexdigest()


#: extra nvcc flags from the environment (experiment switches such as -DNAF_WS_EXP=1)
EXTRA_FLAGS = os.environ.get("NAF_NVCC_FLAGS", "").split()


def _flags(verbose: bool = False) -> list[str]:
    flags = ["-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
             "-Xcompiler", "-fPIC,-O3,-fvisibility=hidden",
             "-I", os.path.join(ROOT, "include"), "-I", HERE, "-DNAF_BUILDING_LIB", "-DNAF_WITH_TC"]
    if verbose:
        flags += ["-Xptxas", "-v"]
    return flags


def command(verbose: bool = False) -> list[str]:
    """The equivalent single nvcc invocation (what build() does, file by file in parallel)."""
    return [nvcc_path()] + _flags(verbose) + EXTRA_FLAGS + ["-shared", "-o", LIB] + [os.path.join(HERE, s) for s in SOURCES]


def build(force: bool = False, verbose: bool = False) -> str:
    """Incremental: an object is recompiled only when its source, a header, the flags or this script
    changed (per-object stamp files under _obj/); the library is relinked when any object changed."""
    dig = _digest(verbose=verbose)
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
        stamp = obj + ".stamp"
        d = _digest([src], verbose)
        if not force and os.path.isfile(obj) and os.path.isfile(stamp):
            with open(stamp) as fh:
                if fh.read().strip() == d:
                    return src, obj, None
        res = subprocess.run([nvcc] + _flags(verbose) + EXTRA_FLAGS + ["-c", os.path.join(HERE, src), "-o", obj],
                             capture_output=True, text=True)
        if res.returncode == 0:
            with open(stamp, "w") as fh:
                fh.write(d)
        return src, obj, res

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    failed = False
    for src, _, res in results:
        if res is None:
            continue
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
