// Tensor-map TMA (cp.async.bulk.tensor) support: host-side descriptor encoding through the driver
// entry point (no -lcuda link dependency) and the device-side load / store instructions.
// Inline PTX only; no CUTLASS dependency.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "naf_umma.cuh"

namespace naf {
namespace tmap {

// ---- host: cuTensorMapEncodeTiled through cudaGetDriverEntryPoint ------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// rank-4 tiled tensor map; dims / box innermost first; strides in BYTES for dims 1..3 (multiples of 16).
// Returns false if the driver entry point is missing or the descriptor is rejected.
inline bool encode4(CUtensorMap* out, CUtensorMapDataType dt, const void* base, const uint64_t (&dims)[4],
                    const uint64_t (&strides_b)[3], const uint32_t (&box)[4], CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t gd[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t gs[3] = {strides_b[0], strides_b[1], strides_b[2]};
  cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t es[4] = {1, 1, 1, 1};
  return fn(out, dt, 4, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---- device ------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}

// global (tensor map, 4-D box at coordinates c0..c3, innermost first) -> shared, completes `bytes of the
// box` on the mbarrier (the caller arms it with mbar_expect_tx)
__device__ __forceinline__ void load4(void* sdst, const CUtensorMap* m, int c0, int c1, int c2, int c3,
                                      uint64_t* mbar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(umma::smem_u32(sdst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
      "r"(umma::smem_u32(mbar))
      : "memory");
}

// shared (dense box image, swizzled as the map says) -> global; bulk-group completion
__device__ __forceinline__ void store4(const CUtensorMap* m, int c0, int c1, int c2, int c3, const void* ssrc) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(umma::smem_u32(ssrc))
               : "memory");
}

// the same with an L2 cache-policy operand (createpolicy.fractional.L2::evict_first / evict_last)
__device__ __forceinline__ void store4_hint(const CUtensorMap* m, int c0, int c1, int c2, int c3, const void* ssrc,
                                            uint64_t policy) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%1, %2, %3, %4}], [%5], %6;"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(umma::smem_u32(ssrc)), "l"(policy)
               : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}

// byte offset of 16-byte chunk `c` (0..7) of 128-byte row `row` inside a SWIZZLE_128B box image whose base
// is 1024-byte aligned: the chunk index is XORed with the row index modulo 8
__device__ __forceinline__ uint32_t swz128(uint32_t row, uint32_t c) { return row * 128u + ((c ^ (row & 7u)) << 4); }

}  // namespace tmap
}  // namespace naf
