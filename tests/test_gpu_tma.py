"""GPU: unit tests of the tensor-map TMA pieces (naf_b200/csrc/naf_tmap.cuh) through the probe kernels
of tests/cuda/tma_probe.cu: a K x K window of the channel-group-plane K / V tensors fetched by one
cp.async.bulk.tensor and consumed in place by tcgen05.mma through strided descriptors, and the
SWIZZLE_128B tensor store of an output tile."""
import ctypes as C
import os
import subprocess

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cuda", "tma_probe.cu")
LIB = os.path.join(ROOT, "tests", "cuda", "libtma_probe.so")


def build_probe():
    from naf_b200.csrc.build import nvcc_path

    deps = [SRC] + [os.path.join(ROOT, "naf_b200", "csrc", f) for f in ("naf_umma.cuh", "naf_tmap.cuh")]
    if os.path.isfile(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(d) for d in deps):
        return LIB
    subprocess.check_call([nvcc_path(), "-std=c++17", "-O2", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
                           "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(ROOT, "naf_b200", "csrc"),
                           "-o", LIB, SRC])
    return LIB


@pytest.fixture(scope="module")
def probe():
    lib = C.CDLL(build_probe())
    lib.tma_window_probe.restype = C.c_int
    lib.tma_window_probe.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_void_p]
    lib.tma_store_probe.restype = C.c_int
    lib.tma_store_probe.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    return lib


# K, GB (channel groups in the box), mode (0 = QK / K-major B, 1 = PV / MN-major B)
@pytest.mark.parametrize("K,GB,mode", [(7, 8, 0), (11, 8, 0), (3, 8, 0), (5, 8, 0), (9, 8, 0),
                                       (7, 24, 1), (11, 16, 1), (11, 32, 1), (7, 12, 1), (3, 4, 1), (9, 16, 1)])
def test_window_load_feeds_the_mma_in_place(probe, K, GB, mode):
    dev = torch.device("cuda", 0)
    G, h, w = GB + 5, K + 6, K + 9
    g = torch.Generator(device="cpu").manual_seed(K * 100 + GB)
    planes = torch.randn(G, h, w, 8, generator=g).half().to(dev)
    wx0, wy0, g0 = 3, 2, 4
    K2 = K * K
    TP = (K2 + 15) // 16 * 16
    win = planes[g0:g0 + GB, wy0:wy0 + K, wx0:wx0 + K].float()            # (GB, K, K, 8)
    mat = win.permute(1, 2, 0, 3).reshape(K2, GB * 8)                      # [tap][channel]
    if mode == 0:
        A = torch.randn(128, GB * 8, generator=g).half().to(dev)
        N = TP
        want = A.float() @ mat.t()                                         # (128, K2)
    else:
        A = torch.zeros(128, TP)
        A[:, :K2] = torch.randn(128, K2, generator=g)
        A = A.half().to(dev)
        N = GB * 8
        want = A.float()[:, :K2] @ mat                                     # (128, N)
    D = torch.full((128, N), float("nan"), device=dev)
    rc = probe.tma_window_probe(planes.data_ptr(), G, h, w, A.data_ptr(), D.data_ptr(), K, GB, mode, wx0, wy0, g0,
                                0x7E007E00, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    got = D[:, :K2] if mode == 0 else D
    err = (got - want).abs().max().item()
    assert err <= 1e-3 * max(1.0, want.abs().max().item()), (K, GB, mode, err)


@pytest.mark.parametrize("th,tw,NB", [(4, 28, 6), (7, 14, 3), (2, 56, 6), (2, 64, 6), (1, 37, 4), (4, 28, 2), (3, 37, 8)])
def test_swizzled_tensor_store(probe, th, tw, NB):
    dev = torch.device("cuda", 0)
    B, Ho, Wo, Cn = 2, 3 * th + 5, 2 * tw + 3, 32 * NB + 64
    out = torch.zeros(B, Ho, Wo, Cn, device=dev)
    g = torch.Generator(device="cpu").manual_seed(th * 100 + tw)
    src = torch.randn(th * tw, NB * 32, generator=g).to(dev)
    c0, x0, y0, b = 32, tw + 1, 2 * th, 1
    rc = probe.tma_store_probe(out.data_ptr(), B, Ho, Wo, Cn, src.data_ptr(), th, tw, NB, c0, x0, y0, b,
                               torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    want = torch.zeros_like(out)
    want[b, y0:y0 + th, x0:x0 + tw, c0:c0 + 32 * NB] = src.view(th, tw, NB * 32)
    assert torch.equal(out, want)
