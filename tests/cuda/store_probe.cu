// Micro-benchmark of the output store path of the attention epilogue, isolated from all compute: one CTA
// per SM writes "tiles" of th x tw pixels x DV fp32 channels of a (B, Ho, Wo, C) tensor from shared memory.
//   mode 0: tensor-map stores, SWIZZLE_128B, box = 32 channels x tw x th, `per_group` boxes per bulk group,
//           ping-pong over `slots` staging slots (wait_group.read <slots-1> before re-using a slot)
//   mode 1: 1-D bulk copies, one per (pixel row, half): 256 threads x (DV/2*4) bytes, each thread waits for
//           its own previous copy (naf_xattn_tcws.cu's epilogue)
//   mode 2: tensor-map stores WITHOUT swizzle, box = DV channels x tw x th (one box per tile)
#include <cuda_runtime.h>
#include <stdint.h>

#include "naf_tmap.cuh"

using namespace naf::umma;
namespace tmx = naf::tmap;

// Optional read stream (rd != nullptr): warps 8-11 read `rbytes` bytes per tile from a second buffer with the
// attention kernel's query-load instruction (32-byte ld.global.nc.L1::no_allocate), at most two tiles ahead of
// the store thread -- the read : write mix of the real launch without any compute.
extern "C" __global__ void __launch_bounds__(384, 1)
store_probe_kernel(const __grid_constant__ CUtensorMap map, float* __restrict__ out, int mode, int iters, int th, int tw,
                   int DV, int per_group, int slots, int cells_x, int cells_y, int heads, int Wo, int Ho, int C,
                   const float* __restrict__ rd, int rbytes, long long rd_elems, float* __restrict__ sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ volatile int s_tile;   // tiles the store thread has issued
  const int tid = threadIdx.x;
  if (tid == 0) s_tile = 0;
  if (tid >= 256) {
    __syncthreads();
    if (rd != nullptr && rbytes > 0) {
      const int rt = tid - 256;
      const int per_thread = rbytes / 32 / 128;   // 32-byte loads per reader thread and tile
      float acc = 0.f;
      long long pos = (long long)blockIdx.x * (rbytes / 4);
      for (int t = 0; t < iters * 7; ++t) {
        while (t > s_tile + 2) __nanosleep(64);
        for (int j = 0; j < per_thread; ++j) {
          float v[8];
          const float* src = rd + (pos + (long long)(j * 128 + rt) * 8) % rd_elems;
          asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                       : "l"(src));
          acc += v[0] + v[7];
        }
        pos += (long long)gridDim.x * (rbytes / 4);
      }
      if (acc == 123.456f) *sink = acc;
    }
    return;
  }
  // something finite in the staging area
  for (int i = tid; i < 200 * 1024 / 4; i += 256) reinterpret_cast<float*>(smem)[i] = float(i & 1023);
  fence_proxy_async_smem();
  __syncthreads();   // (the reader warps take part in this one barrier only)
  const int ncell = cells_x * cells_y * heads;
  int cell = blockIdx.x;
  if (mode == 0 || mode == 2) {
    if (tid == 0) {
      const int boxes_per_tile = mode == 0 ? DV / 32 : 1;
      const int box_bytes = mode == 0 ? 128 * 128 : ((th * tw * DV * 4 + 1023) / 1024 * 1024);
      int grp = 0;
      for (int it = 0; it < iters; ++it) {
        // walk cells round robin over the CTAs; tiles down the cell
        const int c = cell % ncell;
        cell += gridDim.x;
        const int head = c % heads, cx = (c / heads) % cells_x, cy = c / heads / cells_x;
        const int tiles = 7;
        for (int t = 0; t < tiles; ++t) {
          for (int b0 = 0; b0 < boxes_per_tile; b0 += per_group, ++grp) {
            const int slot = grp % slots;
            if (slots == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            else if (slots == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
            for (int j = 0; j < per_group && b0 + j < boxes_per_tile; ++j)
              tmx::store4(&map, head * DV + (mode == 0 ? (b0 + j) * 32 : 0), cx * tw, cy * th * tiles + t * th, 0,
                          smem + size_t(slot) * per_group * box_bytes + size_t(j) * box_bytes);
            bulk_commit();
          }
          s_tile = it * tiles + t + 1;
        }
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else {
    const int row = tid & 127, half = tid >> 7;
    const int bytes = DV / 2 * 4;
    uint8_t* my = smem + row * (DV * 4 + 16) + half * bytes;
    for (int it = 0; it < iters; ++it) {
      const int c = cell % ncell;
      cell += gridDim.x;
      const int head = c % heads, cx = (c / heads) % cells_x, cy = c / heads / cells_x;
      const int tiles = 7;
      for (int t = 0; t < tiles; ++t) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        if (row < th * tw) {
          const int y = cy * th * tiles + t * th + row / tw, x = cx * tw + row % tw;
          bulk_store(out + (int64_t(y) * Wo + x) * C + head * DV + half * (DV / 2), my, bytes);
        }
        bulk_commit();
        if (tid == 0) s_tile = it * tiles + t + 1;
      }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

extern "C" __attribute__((visibility("default")))
int store_probe_rw(float* out, int Ho, int Wo, int C, int mode, int iters, int th, int tw, int DV, int per_group, int slots,
                   int grid, void* stream, const float* rd, int rbytes, long long rd_elems, float* sink);

extern "C" __attribute__((visibility("default")))
int store_probe(float* out, int Ho, int Wo, int C, int mode, int iters, int th, int tw, int DV, int per_group, int slots,
                int grid, void* stream) {
  return store_probe_rw(out, Ho, Wo, C, mode, iters, th, tw, DV, per_group, slots, grid, stream, nullptr, 0, 1, nullptr);
}

extern "C" __attribute__((visibility("default")))
int store_probe_rw(float* out, int Ho, int Wo, int C, int mode, int iters, int th, int tw, int DV, int per_group, int slots,
                   int grid, void* stream, const float* rd, int rbytes, long long rd_elems, float* sink) {
  CUtensorMap map;
  const uint64_t dims[4] = {uint64_t(C), uint64_t(Wo), uint64_t(Ho), 1};
  const uint64_t strides[3] = {uint64_t(C) * 4, uint64_t(Wo) * C * 4, uint64_t(Ho) * Wo * C * 4};
  const uint32_t box0[4] = {32, uint32_t(tw), uint32_t(th), 1};
  const uint32_t box2[4] = {uint32_t(DV), uint32_t(tw), uint32_t(th), 1};
  if (!tmx::encode4(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, out, dims, strides, mode == 2 ? box2 : box0,
                    mode == 2 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B))
    return -1;
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(store_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int cells_x = Wo / tw, cells_y = Ho / (th * 7), heads = C / DV;
  store_probe_kernel<<<grid, 384, smem, static_cast<cudaStream_t>(stream)>>>(map, out, mode, iters, th, tw, DV, per_group,
                                                                           slots, cells_x, cells_y, heads, Wo, Ho, C, rd,
                                                                           rbytes, rd_elems, sink);
  return int(cudaGetLastError());
}
