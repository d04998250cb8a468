// Unit probes for the tensor-map TMA pieces of the attention kernel (naf_b200/csrc/naf_tmap.cuh):
//
//  tma_window_probe : one K x K window of a channel-group-plane tensor (G, h, w, 8) fp16 is fetched
//                     by ONE cp.async.bulk.tensor.4d into shared memory, where it lands as
//                     [channel group][tap][16 B].  That image is then used DIRECTLY as the B operand
//                     of tcgen05.mma through strided SWIZZLE_NONE descriptors:
//                       mode 0 (QK, K-major B):  D[128 x TP]   = A[128 x 8*GB] . W^T,  LBO = K*K*16, SBO = 128
//                       mode 1 (PV, MN-major B): D[128 x 8*GB] = A[128 x TP]   . W,    LBO = 128, SBO = K*K*16
//                     with TP = K*K rounded up to 16: the rows / taps beyond K*K read whatever follows.
//  tma_store_probe  : a th x tw pixel tile of NB*32 fp32 channels is staged in SWIZZLE_128B box images
//                     and written to a (B, Ho, Wo, C) tensor by NB tensor stores issued by one thread.
#include <cuda_runtime.h>
#include <stdint.h>

#include "naf_tmap.cuh"

using namespace naf::umma;
namespace tmx = naf::tmap;

extern "C" __global__ void __launch_bounds__(128)
tma_window_probe_kernel(const __grid_constant__ CUtensorMap map, const __half* __restrict__ A, float* __restrict__ D,
                        int K, int GB, int TP, int mode, int wx0, int wy0, int g0, uint32_t poison) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t mbar_ld, mbar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int K2 = K * K;
  const int win_bytes = GB * K2 * 16;
  const int KD = mode == 0 ? GB * 8 : TP;          // reduction length of the GEMM
  const int N = mode == 0 ? TP : GB * 8;
  uint8_t* sW = smem;                               // window image [g][tap][16 B]
  uint8_t* sPad = sW + win_bytes;                   // 256 B of zeros (what taps >= K*K of the last plane read)
  uint8_t* sA = smem + ((win_bytes + 256 + 4096 + 1023) / 1024) * 1024;   // [KD/8][128][16 B]
  // poison everything behind the window (NaN bit patterns), then the zero pad
  for (int i = tid; i < (int(sA - sW) / 4); i += 128) reinterpret_cast<uint32_t*>(sW)[i] = poison;
  __syncthreads();
  for (int i = tid; i < 64; i += 128) reinterpret_cast<uint32_t*>(sPad)[i] = 0u;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) { mbar_init(&mbar_ld, 1); mbar_init(&mbar_mma, 1); fence_mbar_init(); }
  for (int c = 0; c < KD / 8; ++c)
    *reinterpret_cast<uint4*>(sA + (size_t(c) * 128 + tid) * 16) = *reinterpret_cast<const uint4*>(A + size_t(tid) * KD + c * 8);
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    mbar_expect_tx(&mbar_ld, uint32_t(win_bytes));
    tmx::load4(sW, &map, 0, wx0, wy0, g0, &mbar_ld);
    mbar_wait(&mbar_ld, 0);
    fence_after_sync();
    const uint32_t idesc = make_idesc_f16(128, N, false, mode == 1);
    for (int kk = 0; kk < KD / 16; ++kk) {
      const uint64_t da = make_desc(smem_u32(sA) + kk * 2 * (128 * 16), 128 * 16, 128);
      const uint64_t db = mode == 0 ? make_desc(smem_u32(sW) + kk * 2 * (K2 * 16), K2 * 16, 128)
                                    : make_desc(smem_u32(sW) + kk * 256, 128, K2 * 16);
      mma_f16_ss(tmem, da, db, idesc, kk > 0);
    }
    commit(&mbar_mma);
  }
  mbar_wait(&mbar_mma, 0);
  fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t r[8];
    tmem_ld8(tmem + (uint32_t(warp * 32) << 16) + c0, r);
    wait_ld();
    for (int j = 0; j < 8; ++j) D[size_t(tid) * N + c0 + j] = __uint_as_float(r[j]);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

extern "C" __attribute__((visibility("default")))
int tma_window_probe(const void* planes, int G, int h, int w, const void* A, float* D, int K, int GB, int mode,
                     int wx0, int wy0, int g0, unsigned poison, void* stream) {
  CUtensorMap map;
  const uint64_t dims[4] = {8, uint64_t(w), uint64_t(h), uint64_t(G)};
  const uint64_t strides[3] = {16, uint64_t(w) * 16, uint64_t(h) * w * 16};
  const uint32_t box[4] = {8, uint32_t(K), uint32_t(K), uint32_t(GB)};
  if (!tmx::encode4(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, planes, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE))
    return -1;
  const int K2 = K * K, TP = (K2 + 15) / 16 * 16;
  const int win_bytes = GB * K2 * 16;
  const int KD = mode == 0 ? GB * 8 : TP;
  const size_t smem = size_t((win_bytes + 256 + 4096 + 1023) / 1024) * 1024 + size_t(KD / 8) * 128 * 16 + 1024;
  cudaFuncSetAttribute(tma_window_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  tma_window_probe_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(
      map, static_cast<const __half*>(A), D, K, GB, TP, mode, wx0, wy0, g0, poison);
  return int(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(128)
tma_store_probe_kernel(const __grid_constant__ CUtensorMap map, const float* __restrict__ src, int th, int tw, int NB,
                       int c0, int x0, int y0, int b) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x;
  const int npix = th * tw;
  const int box_stride = (npix * 128 + 1023) / 1024 * 1024;   // every box image 1024-byte aligned (swizzle atom)
  // thread <-> pixel of the tile (row-major), as the attention epilogue (thread <-> TMEM lane)
  if (tid < npix) {
    for (int j = 0; j < NB; ++j) {
      uint8_t* boxp = smem + size_t(j) * box_stride;
      for (int c = 0; c < 8; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(src + (size_t(tid) * NB + j) * 32 + c * 4);
        *reinterpret_cast<float4*>(boxp + tmx::swz128(tid, c)) = v;
      }
    }
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (tid == 0) {
    for (int j = 0; j < NB; ++j) tmx::store4(&map, c0 + 32 * j, x0, y0, b, smem + size_t(j) * box_stride);
    bulk_commit();
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

extern "C" __attribute__((visibility("default")))
int tma_store_probe(float* out, int B, int Ho, int Wo, int C, const float* src, int th, int tw, int NB, int c0, int x0,
                    int y0, int b, void* stream) {
  CUtensorMap map;
  const uint64_t dims[4] = {uint64_t(C), uint64_t(Wo), uint64_t(Ho), uint64_t(B)};
  const uint64_t strides[3] = {uint64_t(C) * 4, uint64_t(Wo) * C * 4, uint64_t(Ho) * Wo * C * 4};
  const uint32_t box[4] = {32, uint32_t(tw), uint32_t(th), 1};
  if (!tmx::encode4(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, out, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))
    return -1;
  const size_t smem = size_t(NB) * ((th * tw * 128 + 1023) / 1024 * 1024) + 1024;
  cudaFuncSetAttribute(tma_store_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  tma_store_probe_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(map, src, th, tw, NB, c0, x0, y0, b);
  return int(cudaGetLastError());
}
