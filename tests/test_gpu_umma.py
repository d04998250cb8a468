"""GPU: unit test of the hand-written tcgen05 primitives (descriptors, TMEM, commit/mbarrier)
against torch.matmul, through a tiny probe kernel compiled from tests/cuda/umma_probe.cu."""
import ctypes as C
import os
import subprocess

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cuda", "umma_probe.cu")
LIB = os.path.join(ROOT, "tests", "cuda", "libumma_probe.so")


def build_probe():
    from naf_b200.csrc.build import nvcc_path

    if os.path.isfile(LIB) and os.path.getmtime(LIB) >= max(
            os.path.getmtime(SRC), os.path.getmtime(os.path.join(ROOT, "naf_b200", "csrc", "naf_umma.cuh"))):
        return LIB
    subprocess.check_call([nvcc_path(), "-std=c++17", "-O2", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
                           "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(ROOT, "naf_b200", "csrc"),
                           "-o", LIB, SRC])
    return LIB


@pytest.fixture(scope="module")
def probe():
    lib = C.CDLL(build_probe())
    lib.umma_probe.restype = C.c_int
    lib.umma_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    return lib


@pytest.mark.parametrize("N,K", [(64, 64), (192, 64), (256, 128), (16, 16), (96, 32)])
@pytest.mark.parametrize("b_mn", [0, 1])
@pytest.mark.parametrize("a_tmem", [0, 1])
def test_umma_probe(probe, N, K, b_mn, a_tmem):
    dev = torch.device("cuda", 0)
    g = torch.Generator(device="cpu").manual_seed(N * 1000 + K)
    A = torch.randn(128, K, generator=g).half().to(dev)
    Bnk = torch.randn(N, K, generator=g).half().to(dev)
    Bsrc = Bnk.t().contiguous() if b_mn else Bnk
    D = torch.full((128, N), float("nan"), device=dev)
    rc = probe.umma_probe(A.data_ptr(), Bsrc.data_ptr(), D.data_ptr(), N, K, b_mn, a_tmem,
                          torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    want = A.float() @ Bnk.float().t()
    err = (D - want).abs().max().item()
    assert err <= 1e-3 * max(1.0, want.abs().max().item()), (N, K, b_mn, a_tmem, err)
