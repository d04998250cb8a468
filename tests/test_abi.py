"""CPU: the C-ABI library loads, exports every symbol include/naf_b200.h declares, its structs
match the ctypes mirrors, and argument validation / kernel selection work without a GPU."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest
import torch

import naf_b200
from naf_b200 import _lib, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "naf_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"NAF_API\s+[\w\s\*]+?\b(naf_\w+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 8
    assert set(syms) == set(_lib.EXPORTS)
    for s in syms:
        assert getattr(lib, s) is not None


def test_abi_version_and_error_string():
    lib = _lib.load()
    assert lib.naf_abi_version() == _lib.ABI_VERSION
    assert isinstance(lib.naf_last_error(), bytes)


def test_struct_layouts_match_the_header():
    """Compile a tiny C program against the header and compare sizeof/offsetof with ctypes."""
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "naf_b200.h"
int main(void){
  printf("%zu %zu %zu %zu\n", sizeof(naf_kpool_params), offsetof(naf_kpool_params, B), offsetof(naf_kpool_params, x_stride_b), offsetof(naf_kpool_params, x_stride_x));
  printf("%zu %zu %zu %zu %zu\n", sizeof(naf_xattn_params), offsetof(naf_xattn_params, B), offsetof(naf_xattn_params, scale), offsetof(naf_xattn_params, q_stride_b), offsetof(naf_xattn_params, algo));
  printf("%zu %zu %zu %zu %zu\n", offsetof(naf_xattn_params, workspace), offsetof(naf_xattn_params, q_dtype), offsetof(naf_xattn_params, Kw), sizeof(naf_xattn_bwd_params), offsetof(naf_xattn_bwd_params, Kw));
  return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        cfile, exe = os.path.join(d, "t.c"), os.path.join(d, "t")
        open(cfile, "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), cfile, "-o", exe])
        l1, l2, l3 = subprocess.check_output([exe], text=True).strip().splitlines()
    kp, xp = _lib.KPoolParams, _lib.XAttnParams
    assert [int(v) for v in l1.split()] == [C.sizeof(kp), kp.B.offset, kp.x_stride_b.offset, kp.x_stride_x.offset]
    assert [int(v) for v in l2.split()] == [C.sizeof(xp), xp.B.offset, xp.scale.offset, xp.q_stride_b.offset, xp.algo.offset]
    bp = _lib.XAttnBwdParams
    assert [int(v) for v in l3.split()] == [xp.workspace.offset, xp.q_dtype.offset, xp.Kw.offset, C.sizeof(bp), bp.Kw.offset]


def test_kernel_selection_for_baseline_configs():
    sel = ops.select_algo
    assert sel((1, 256, 224, 224), (1, 384, 16, 16), 4, 7) == "cell_tcws"      # C1: latency bound, no pre-pass
    assert sel((8, 256, 896, 896), (8, 768, 32, 32), 4, 7) == "cell_tma"       # C2
    assert sel((4, 256, 1036, 1036), (4, 1024, 37, 37), 4, 11) == "cell_tma"   # C3: single-pass wide heads
    assert sel((16, 256, 1344, 1344), (16, 768, 24, 24), 4, 7) == "cell_tcws"  # C4: 56x56-pixel cells
    assert sel((4, 256, 2048, 2048), (4, 768, 32, 32), 4, 7) == "cell_tcws"    # C5: 64x64-pixel cells
    assert sel((1, 64, 32, 32), (1, 16, 13, 13), 4, 9) == "union_tc"    # non-integer ratio (tap tables)
    assert sel((8, 256, 32, 32), (8, 384, 13, 13), 4, 9) == "union_tc"  # the reference's training shape
    assert sel((8, 256, 256, 256), (8, 3, 256, 256), 1, 15) == "union_tc"  # denoising: ratio 1, C = 3, one head
    assert sel((1, 64, 32, 32), (1, 16, 13, 13), 4, 9, return_scores=True) == "generic"   # per-tap scores
    assert sel((1, 96, 32, 32), (1, 16, 13, 13), 4, 9) == "generic"     # head dim 24: not a multiple of 16
    assert sel((8, 256, 896, 896), (8, 768, 32, 32), 4, (7, 5)) == "generic"   # rectangular window (NATTEN (kh, kw))
    assert sel((8, 256, 896, 896), (8, 768, 32, 32), 4, (7, 7)) == "cell_tma"  # a pair with kh == kw is a square window
    assert sel((1, 256, 36, 36), (1, 32, 9, 9), 4, 7, return_scores=True) == "generic"      # 16-pixel cells
    assert sel((1, 256, 224, 224), (1, 384, 16, 16), 4, 7, return_scores=True) == "cell_tma"  # scores on the fast kernel


def test_validation_errors_map_to_reference_exceptions():
    with pytest.raises(ValueError):       # NATTEN: kernel_size * dilation > size
        ops.select_algo((1, 64, 36, 36), (1, 16, 4, 4), 4, 7)
    with pytest.raises(ValueError):       # even kernel
        ops.select_algo((1, 64, 36, 36), (1, 16, 9, 9), 4, 4)
    p = _lib.XAttnParams()
    p.q = p.k = p.v = p.out = 1 << 20
    p.B, p.D, p.C, p.heads, p.Ho, p.Wo, p.h, p.w, p.K = 1, 30, 16, 4, 36, 36, 9, 9, 3
    p.q_stride_b, p.q_stride_y, p.q_stride_x = 36 * 36 * 30, 36 * 30, 30
    rc = _lib.load().naf_xattn_select_algo(C.byref(p))
    assert rc == -_lib.NAF_ERR_BAD_SHAPE
    with pytest.raises(AssertionError, match="dim must be divisible by num_heads"):
        _lib.check(-rc, "x")
    assert _lib.load().naf_xattn_fwd_f32(None, None) == _lib.NAF_ERR_NULL


def test_cpu_tensors_are_rejected_loudly():
    q, k, v = torch.zeros(1, 8, 4, 4), torch.zeros(1, 8, 2, 2), torch.zeros(1, 4, 2, 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.xattn(q, k, v, 2, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        naf_b200.RoPE(8, num_heads=2)(q)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        naf_b200.CrossAttention(8, 2, (1, 1))(q, k, v)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libnaf_b200.so")
    with pytest.raises(_lib.NafLibraryError, match="no PyTorch/CPU fallback"):
        _lib.load()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "naf_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "/root/reference" not in text, f
