// Unit probe for the hand-written tcgen05 primitives in naf_b200/csrc/naf_umma.cuh:
//   D[128 x N] (fp32) = A[128 x K] * B^T      A, B fp16
// one CTA of 128 threads; A from shared memory (K-major canonical layout) or from TMEM,
// B K-major ((N x K) row-major in global) or MN-major ((K x N) row-major in global).
// Built by tests/test_gpu_umma.py with nvcc for sm_100a; compared against torch.matmul.
#include <cuda_runtime.h>
#include <stdint.h>

#include "naf_umma.cuh"

using namespace naf::umma;

extern "C" __global__ void __launch_bounds__(128)
umma_probe_kernel(const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ D,
                  int N, int K, int b_mn_major, int a_from_tmem) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KC = K / 8;                       // 16-byte chunks along K
  uint8_t* sA = smem;                         // [KC][128][16 B]
  uint8_t* sB = smem + size_t(KC) * 128 * 16; // K-major: [KC][N][16 B]; MN-major: [K/8][N/8][8][16 B]

  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) { mbar_init(&mbar, 1); fence_mbar_init(); }
  // ---- stage A: thread <-> row
  for (int c = 0; c < KC; ++c) {
    const uint4 v = *reinterpret_cast<const uint4*>(A + size_t(tid) * K + c * 8);
    *reinterpret_cast<uint4*>(sA + (size_t(c) * 128 + tid) * 16) = v;
  }
  // ---- stage B
  if (!b_mn_major) {
    for (int i = tid; i < N * KC; i += 128) {
      const int n = i / KC, c = i % KC;
      const uint4 v = *reinterpret_cast<const uint4*>(B + size_t(n) * K + c * 8);
      *reinterpret_cast<uint4*>(sB + (size_t(c) * N + n) * 16) = v;
    }
  } else {
    const int NG = N / 8;
    for (int i = tid; i < K * NG; i += 128) {
      const int k = i / NG, g = i % NG;
      const uint4 v = *reinterpret_cast<const uint4*>(B + size_t(k) * N + g * 8);
      *reinterpret_cast<uint4*>(sB + (size_t(k / 8) * NG + g) * 128 + (k % 8) * 16) = v;
    }
  }
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tD = tmem;            // columns [0, N)
  const uint32_t tA = tmem + 256;      // columns [256, 256 + K/2)
  if (a_from_tmem) {
    // lane <-> row; K fp16 values packed two per 32-bit column
    for (int c = 0; c < KC; c += 2) {
      uint32_t r[8];
      const uint4 v0 = *reinterpret_cast<const uint4*>(A + size_t(tid) * K + c * 8);
      const uint4 v1 = *reinterpret_cast<const uint4*>(A + size_t(tid) * K + (c + 1) * 8);
      r[0] = v0.x; r[1] = v0.y; r[2] = v0.z; r[3] = v0.w;
      r[4] = v1.x; r[5] = v1.y; r[6] = v1.z; r[7] = v1.w;
      tmem_st8(tA + (uint32_t(warp * 32) << 16) + c * 4, r);
    }
    wait_st();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
  }
  if (tid == 0) {
    const uint32_t idesc = make_idesc_f16(128, N, false, b_mn_major != 0);
    for (int kk = 0; kk < K / 16; ++kk) {
      const uint64_t da = make_desc(smem_u32(sA) + kk * 2 * (128 * 16), 128 * 16, 128);
      uint64_t db;
      if (!b_mn_major) db = make_desc(smem_u32(sB) + kk * 2 * (N * 16), N * 16, 128);
      else db = make_desc(smem_u32(sB) + kk * 2 * (N / 8) * 128, (N / 8) * 128, 128);
      if (a_from_tmem) mma_f16_ts(tD, tA + kk * 8, db, idesc, kk > 0);
      else mma_f16_ss(tD, da, db, idesc, kk > 0);
    }
    commit(&mbar);
  }
  mbar_wait(&mbar, 0);
  fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tD + (uint32_t(warp * 32) << 16) + c0, r);
    wait_ld();
    for (int j = 0; j < 16; ++j) D[size_t(tid) * N + c0 + j] = __uint_as_float(r[j]);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

extern "C" __attribute__((visibility("default")))
int umma_probe(const void* A, const void* B, float* D, int N, int K, int b_mn_major,
               int a_from_tmem, void* stream) {
  const size_t smem = size_t(K / 8) * 128 * 16 + size_t(K) * N * 2;
  cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  umma_probe_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(A), static_cast<const __half*>(B), D, N, K, b_mn_major, a_from_tmem);
  return int(cudaGetLastError());
}
