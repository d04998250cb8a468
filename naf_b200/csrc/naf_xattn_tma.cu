// Cell kernel, TMA edition: the warp-specialised tcgen05 / TMEM pipeline of naf_xattn_tcws.cu with every
// bulk data movement except the per-pixel query loads handed to the tensor-map TMA engine.
//
//   * K / V windows.  A pre-pass (kv_planes_kernel) splits the low-resolution key and value maps ONCE into
//     fp16 hi / lo "channel-group planes"  (B, 2, C/8, h, w, 8).  The K x K window of a (cell, head) item
//     is then ONE cp.async.bulk.tensor.4d box per plane (UTMALDG): it lands in shared memory as
//     [channel group][tap][16 B], which IS a valid SWIZZLE_NONE UMMA operand image --
//         QK^T : K-major  B, LBO = K*K*16 (next channel group), SBO = 128 (next 8 taps)
//         P V  : MN-major B, LBO = 128 (next 8 taps),           SBO = K*K*16 (next channel group)
//     (tests/test_gpu_tma.py pins both against torch.matmul).  No thread converts or stores a window
//     element any more (naf_xattn_tcws.cu re-converted every K / V element once per window holding it,
//     49 or 121 times), and the window buffers shrink from TP-padded to exact K*K taps.
//   * Wide value heads.  O is accumulated in NH column halves of DVH channels against ONE P
//     (P stays in TMEM over the S columns): K = 11, dv = 256 (BASELINE config C3) runs in a single pass
//     -- scores, softmax and the fp16 split of P are computed once per tile instead of once per value slab.
//   * Output.  The epilogue stages normalised O in SWIZZLE_128B box images (conflict-free 16-byte
//     st.shared) and ONE thread issues a handful of cp.async.bulk.tensor.4d stores per tile (UTMASTG:
//     box = 32 channels x cell width x tile rows) instead of 256 per-thread bulk copies.
//
//   warps 0-7  "front": q half rows HBM -> registers (two tiles ahead), RoPE, fp16 hi/lo split, tcgen05.st;
//                       softmax on S (tcgen05.ld), un-normalised P back over S, row sums to TMEM
//   warps 8-15 "back" : drain O (tcgen05.ld), normalise, swizzled staging, tensor stores
//   warp 16    "mma"  : one thread issues every tcgen05.mma:  QK(0) QK(1) PV(0,*) QK(2) PV(1,*) ...
//   warp 17    "load" : one thread issues the window boxes of item i+1 while item i computes
//
// Tiles are whole pixel rows of the cell (th rows x rw pixels <= 128 TMEM lanes), so that a tile is a
// rectangle of the output image.  Shapes this kernel does not take (cell width > 128 pixels, value dims
// outside the table below) stay on naf_xattn_tcws.cu.
//
// Reference semantics: src/layers/attentions.py:16-29,53-75; RoPE src/layers/rope.py:137-153.
#include <cuda_bf16.h>

#include <cstring>
#include <type_traits>

#include "naf_common.cuh"
#include "naf_tmap.cuh"
#include "naf_umma.cuh"

namespace naf {

using namespace umma;

namespace {

constexpr int DQ = 64;
constexpr int KC = DQ / 8;
constexpr int NFRONT = 256, NBACK = 256, NTHREADS = NFRONT + NBACK + 64;
constexpr int MMA_WARP = (NFRONT + NBACK) / 32, LOAD_WARP = MMA_WARP + 1;
#ifndef NAF_TMA_EXP
#define NAF_TMA_EXP 0   // profiling variants (scripts/build_variant.py): 1 = no output stores, 2 = no q loads,
#endif                  // 4 = softmax without the exponentials (P = 1 on the valid taps), 8 = constant cos/sin (no table loads),
                        // 16 = P V without the Plo * Vhi pass
#ifndef NAF_TMA_NOB
#define NAF_TMA_NOB 1   // accumulator buffers in TMEM for the wide value heads.  2 = the head is accumulated in parts of
#endif                  // 96 (dv 192) or 64 (dv 256) columns alternating between two buffers, so that the drain of one part
                        // overlaps the P V of the next: measured SLOWER (pipeline alone 2.52 -> 3.08 ms at C2, 3.94 -> 5.16 ms
                        // at C3, profiles/r3_tma_experiments.txt).  A tcgen05.mma whose A operand lives in TMEM costs
                        // about 64 clocks however narrow N is, so halving N doubles the tensor time of P V.
#ifndef NAF_TMA_QBULK
#define NAF_TMA_QBULK 0   // bulk L2 prefetch of a cell's query rows by the loader warp, one item ahead (measured: C2 4.30 ->
#endif                    // 4.55 ms, x at target resolution 4.97 -> 6.5 ms: the output stream evicts the rows before they are used)
#ifndef NAF_TMA_STHINT
#define NAF_TMA_STHINT 0  // L2 policy of the output tensor stores: 0 none, 1 evict_first, 2 evict_last
#endif
#ifndef NAF_TMA_QHINT
#define NAF_TMA_QHINT 0   // 1 = evict_last policy on the query loads
#endif
#ifndef NAF_TMA_QPF
#define NAF_TMA_QPF 0   // L2 prefetch of the query rows, in tiles ahead of the register loads (0 = none: measured
#endif                  // neutral at 3 tiles, -5 % at 6: the loads are not what the front group waits for)
constexpr int kSmemLimit = 227 * 1024 - 1024;   // dynamic shared memory we allow ourselves (static barriers extra)

template <int TP>
struct TmaWindowOf { static constexpr int K = TP == 16 ? 3 : TP == 32 ? 5 : TP == 64 ? 7 : TP == 96 ? 9 : 11; };

constexpr int round_up(int v, int a) { return (v + a - 1) / a * a; }

// TP: padded taps; DVH: value channels per accumulator part; NH: parts per tile; ROUNDS: staging rounds per part;
// NOB: accumulator buffers in TMEM (2: the drain of part n overlaps the P V of part n+1)
template <int TP, int DVH, int NH, int ROUNDS, int NOB = 1>
struct TmaCfg {
  static constexpr int K = TmaWindowOf<TP>::K, K2 = K * K;
  static constexpr int DV = DVH * NH;
  // one fp16 plane of a window: [channel group][tap][16 B]; the MMAs read TP >= K2 taps, i.e. up to 240 B
  // past the last channel group: every plane is followed by >= 256 B that are zeroed once and never written
  static constexpr int kKPlane = KC * K2 * 16;
  static constexpr int kKStride = round_up(kKPlane + 256, 128);
  static constexpr int kKWin = 2 * kKStride;                    // hi | lo
  static constexpr int kVPlane = (DV / 8) * K2 * 16;
  static constexpr int kVStride = round_up(kVPlane + 256, 128);
  static constexpr int kVWin = 2 * kVStride;
  static constexpr int kRoundCols = DVH / ROUNDS;               // output columns staged per round
  static constexpr int kBox = 128 * 128;                        // one staging box image: 128 rows x 128 B
  static constexpr int kSlot = (kRoundCols / 32) * kBox;        // one staging slot: a round's boxes (fp32: 32 channels per
                                                                // box; bf16: 64, half the boxes)
  static constexpr int kMx = (2 * 2 + 4 * 2) * 128 * 4;         // row-max exchange between the two row halves [2][2][128]
                                                                // + softmax row sums for the epilogue [4][2][128]
  static constexpr int NKB = 2;
  static constexpr int kMin = NKB * kKWin + kVWin + kSlot + kMx + 1024;   // one V buffer, one staging slot
  // what is left goes first to a second staging slot (the tensor stores of round r read slot r&1 while round
  // r+1 is drained into the other one), then to a second V window buffer
  static constexpr int NSTG = (kMin + kSlot <= kSmemLimit) ? 2 : 1;
  static constexpr int kStage = NSTG * kSlot;
  static constexpr int NVB = (kMin + (NSTG - 1) * kSlot + kVWin <= kSmemLimit) ? 2 : 1;
  [[maybe_unused]] static constexpr int kOffK = 0;
  static constexpr int kOffV = NKB * kKWin;
  static constexpr int kOffStage = round_up(kOffV + NVB * kVWin, 1024);
  static constexpr int kOffMx = kOffStage + kStage;
  static constexpr int kSmemBytes = kOffMx + kMx;
  // (the row sums travel through shared memory, not TMEM: 128 + 2*128 + 128 columns is exactly 512, which gives
  // the 11x11 / dv 256 configuration its second Q stage)
  static constexpr int kQStages = (128 + 2 * TP + NOB * DVH <= 512) ? 2 : 1;
  [[maybe_unused]] static constexpr int kTmemQ = 0;
  static constexpr int kTmemS = 64 * kQStages;
  static constexpr int kTmemO = kTmemS + 2 * TP;
  static constexpr int kTmemUsed = kTmemO + NOB * DVH;
  static constexpr bool kFits = kTmemUsed <= 512 && kSmemBytes <= kSmemLimit && (DVH % ROUNDS) == 0 &&
                                (kRoundCols % 32) == 0 && DVH % 16 == 0 && DVH <= 256 && DV / 8 <= 256 &&
                                (NOB == 1 || (NOB == 2 && DVH <= 128));
};

__device__ __forceinline__ uint64_t tm_pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void tm_unpack2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t tm_fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

struct TmFastDiv {
  uint32_t d, magic, shift;
};
struct TmDivs {
  TmFastDiv ntiles, heads, w, h, rw, rep_y, rep_x;
};
__device__ __forceinline__ int tm_div(int n, const TmFastDiv& f) {
  return f.magic == 0 ? (n >> f.shift) : int(__umulhi(uint32_t(n), f.magic) >> f.shift);
}

struct TmItem {
  int b, ci, cj, head;
};
__device__ __forceinline__ TmItem tm_decode(int item, const TmDivs& dv) {
  TmItem c;
  int q = tm_div(item, dv.heads);
  c.head = item - q * int(dv.heads.d);
  item = q;
  q = tm_div(item, dv.w);
  c.cj = item - q * int(dv.w.d);
  item = q;
  q = tm_div(item, dv.h);
  c.ci = item - q * int(dv.h.d);
  c.b = q;
  return c;
}

struct TmGeom {
  int rh, rw;          // cell size in target pixels
  int th;              // image rows per tile; tile_rows = th * rw <= 128
  int ntiles;          // ceil(rh / th)
  int th_last;         // rows of the last tile
  int n_items;
  int groups_k, groups_v;   // channel groups per (batch, plane) of the K / V plane tensors: D/8, C/8
  int tab_bytes;            // RoPE table rows of one cell staged in shared memory per K window buffer: [cos_y | sin_y]
                            // rh rows of 64 B, then [cos_x | sin_x] as 4 chunk planes [chunk][column][16 B] of
                            // xchunk bytes each (a lane reads its column's 16 bytes: conflict-free); 0 = tables from global
  int xchunk;               // bytes of one chunk plane of an x table (rw * 16 rounded up to 128)
};

}  // namespace

// ---- pre-pass: fp32 or bf16 (B,h,w,Cn) map -> fp16 hi / lo channel-group planes (B, 2, Cn/8, h, w, 8) ----
// bf16 inputs (the reference under torch.autocast(bfloat16), train.py:120) are read as they are: no widened
// copy of the features is ever made.
__device__ __forceinline__ void load8_as_float(const float* p, float (&v)[8]) { ldg8(p, v); }
__device__ __forceinline__ void load8_as_float(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
kv_planes_kernel(const T* __restrict__ src, uint4* __restrict__ dst, int B, int h, int w, int Cn) {
  const int G = Cn / 8;
  const int64_t total = int64_t(B) * G * h * w;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    // x fastest (coalesced 16-byte plane writes), then y, channel group, batch
    const int x = int(i % w);
    int64_t r = i / w;
    const int y = int(r % h);
    r /= h;
    const int g = int(r % G);
    const int b = int(r / G);
    float v[8];
    load8_as_float(src + ((int64_t(b) * h + y) * w + x) * Cn + g * 8, v);
    uint4 hi, lo;
    split2_f16(v[0], v[1], hi.x, lo.x);
    split2_f16(v[2], v[3], hi.y, lo.y);
    split2_f16(v[4], v[5], hi.z, lo.z);
    split2_f16(v[6], v[7], hi.w, lo.w);
    const int64_t plane = int64_t(h) * w;
    dst[((int64_t(b) * 2 + 0) * G + g) * plane + int64_t(y) * w + x] = hi;
    dst[((int64_t(b) * 2 + 1) * G + g) * plane + int64_t(y) * w + x] = lo;
  }
}

// 18 warps: one SM sub-partition (16 K registers) hosts 5 of them, hence at most 96 registers per thread
template <int TP, int DVH, int NH, int ROUNDS, int NOB>
__global__ void __launch_bounds__(NTHREADS, 1)
xattn_cell_tma_kernel(naf_xattn_params p, TmGeom gm, TmDivs dv, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                      const __grid_constant__ CUtensorMap tmO2, const __grid_constant__ CUtensorMap tmCX,
                      const __grid_constant__ CUtensorMap tmSX) {
  using Cfg = TmaCfg<TP, DVH, NH, ROUNDS, NOB>;
  constexpr int K = Cfg::K, K2 = Cfg::K2;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_q_full[2], bar_s_full[2], bar_p_full[2], bar_o_full[2], bar_o_free[2];
  __shared__ uint64_t bar_k_full[2], bar_k_free[2], bar_v_full[2], bar_v_free[2];
  __shared__ uint32_t tmem_base_s;

  uint8_t* const sK = smem + Cfg::kOffK;
  uint8_t* const sV = smem + Cfg::kOffV;
  uint8_t* const stage_out = smem + Cfg::kOffStage;
  float* const mx = reinterpret_cast<float*>(smem + Cfg::kOffMx);   // [tile parity][half][row]
  float* const lsum = mx + 2 * 2 * 128;                             // [tile & 3][half][row]
  const float* const sTab = reinterpret_cast<const float*>(smem + Cfg::kSmemBytes);   // [K buffer][tab_bytes]

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform: the role branches are uniform branches
  const int rh = gm.rh, rw = gm.rw;
  const int npix = rh * rw;
  const int ntiles = gm.ntiles;
  const int tile_rows = gm.th * rw;
  const int my_items = (gm.n_items - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  auto item_of = [&](int it_seq) { return tm_decode(blockIdx.x + it_seq * gridDim.x, dv); };

  if (warp == MMA_WARP) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_q_full[s], NFRONT);
      mbar_init(&bar_s_full[s], 1);
      mbar_init(&bar_p_full[s], NFRONT);
      mbar_init(&bar_k_full[s], 1);
      mbar_init(&bar_k_free[s], gm.tab_bytes ? 1 + NFRONT / 32 : 1);   // MMA commit (+ the front warps: table rows)
      mbar_init(&bar_v_full[s], 1);
      mbar_init(&bar_v_free[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_o_full[s], 1);
      mbar_init(&bar_o_free[s], NBACK);
    }
    fence_mbar_init();
  }
  // zero the tails the MMAs over-read behind every window plane (never written afterwards)
  for (int i = tid; i < Cfg::NKB * 2 * (Cfg::kKStride - Cfg::kKPlane) / 16; i += NTHREADS) {
    const int per = (Cfg::kKStride - Cfg::kKPlane) / 16;
    const int pl = i / per, o = i - pl * per;
    *reinterpret_cast<uint4*>(sK + pl * Cfg::kKStride + Cfg::kKPlane + o * 16) = make_uint4(0, 0, 0, 0);
  }
  for (int i = tid; i < Cfg::NVB * 2 * (Cfg::kVStride - Cfg::kVPlane) / 16; i += NTHREADS) {
    const int per = (Cfg::kVStride - Cfg::kVPlane) / 16;
    const int pl = i / per, o = i - pl * per;
    *reinterpret_cast<uint4*>(sV + pl * Cfg::kVStride + Cfg::kVPlane + o * 16) = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const int total_tiles = my_items * ntiles;

  if (warp < 8) {
    // ======================================================================== FRONT
    const int rowgrp = warp & 3, half = warp >> 2;
    const int row = rowgrp * 32 + lane;
    const uint32_t lane_off = uint32_t(rowgrp * 32) << 16;
    constexpr int HALF = DQ / 2, P = DQ / 4, SC = TP / 2;
    const bool rope = p.cos_y != nullptr;
    const bool q_bf16 = p.q_dtype == NAF_DTYPE_BF16;
    const float qscale = p.scale * 1.4426950408889634f;

    float qa[P], qb[P];
    const bool tabs = gm.tab_bytes != 0;
#if NAF_TMA_QHINT
    const uint64_t q_policy = tmap::policy_evict_last();
#endif
    // of the q in the registers: its table rows (tables in shared memory: row / column inside the cell; tables
    // in global memory: the target pixel) and bit 0: first tile of the item, 1: last tile, 2: K buffer, 3: its parity
    int q_r = 0, q_c = 0, q_flags = 0;

    // element offset of this thread's query half row for tile g (clamped to the tile's last pixel)
    auto q_offset = [&](int g, int& it_seq, int& tile, int& ly, int& lx, int& y, int& x) -> int64_t {
      it_seq = tm_div(g, dv.ntiles);
      tile = g - it_seq * ntiles;
      const TmItem it = item_of(it_seq);
      int pi = tile * tile_rows + row;
      const int lim = min(npix, (tile + 1) * tile_rows);
      if (pi >= lim) pi = lim - 1;
      ly = tm_div(pi, dv.rw);
      lx = pi - ly * rw;
      y = it.ci * rh + ly;
      x = it.cj * rw + lx;
      return int64_t(it.b) * p.q_stride_b + it.head * DQ + P * half +
             int64_t(tm_div(y, dv.rep_y)) * p.q_stride_y + int64_t(tm_div(x, dv.rep_x)) * p.q_stride_x;
    };

    int f_g = 0;   // issue_q is called for g = 0, 1, 2, ... in order
    auto issue_q = [&]() {
      const int g = f_g++;
#if NAF_TMA_QPF > 0
      // Under the output stream an HBM read takes several microseconds; the register prefetch covers one
      // tile.  Pull the lines of tile g + NAF_TMA_QPF into L2 now: the two threads of a row cover the two
      // 128-byte lines of the head's query row (one line for bf16).
      if (g + NAF_TMA_QPF < total_tiles && !(NAF_TMA_EXP & 2)) {
        int a, b, c, d, e, f;
        const int64_t off = q_offset(g + NAF_TMA_QPF, a, b, c, d, e, f) + (half ? HALF : 0);
        prefetch_l2(q_bf16 ? static_cast<const void*>(reinterpret_cast<const __nv_bfloat16*>(p.q) + off)
                           : static_cast<const void*>(p.q + off));
      }
#endif
      int seq, tile, ly, lx, y, x;
      const int64_t qoff = q_offset(g, seq, tile, ly, lx, y, x);
      q_r = tabs ? ly : y;
      q_c = tabs ? lx : x;
      q_flags = int(tile == 0) | (int(tile + 1 == ntiles) << 1) | ((seq & 3) << 2);
#if NAF_TMA_EXP & 2
#pragma unroll
      for (int j = 0; j < P; ++j) { qa[j] = 0.01f * float(j + (lx & 7)); qb[j] = -0.02f * float(j); }
      (void)qoff;
#else
      if (q_bf16) {
        // 16 bf16 channels = 32 bytes per rotation half
        const __nv_bfloat16* qp = reinterpret_cast<const __nv_bfloat16*>(p.q) + qoff;
        load8_as_float(qp, *reinterpret_cast<float(*)[8]>(&qa[0]));
        load8_as_float(qp + 8, *reinterpret_cast<float(*)[8]>(&qa[8]));
        load8_as_float(qp + HALF, *reinterpret_cast<float(*)[8]>(&qb[0]));
        load8_as_float(qp + HALF + 8, *reinterpret_cast<float(*)[8]>(&qb[8]));
      } else {
        const float* qp = p.q + qoff;
#if NAF_TMA_QHINT
        ldg_hint8(qp, *reinterpret_cast<float(*)[8]>(&qa[0]), q_policy);
        ldg_hint8(qp + 8, *reinterpret_cast<float(*)[8]>(&qa[8]), q_policy);
        ldg_hint8(qp + HALF, *reinterpret_cast<float(*)[8]>(&qb[0]), q_policy);
        ldg_hint8(qp + HALF + 8, *reinterpret_cast<float(*)[8]>(&qb[8]), q_policy);
#else
        ldg_stream8(qp, *reinterpret_cast<float(*)[8]>(&qa[0]));
        ldg_stream8(qp + 8, *reinterpret_cast<float(*)[8]>(&qa[8]));
        ldg_stream8(qp + HALF, *reinterpret_cast<float(*)[8]>(&qb[0]));
        ldg_stream8(qp + HALF + 8, *reinterpret_cast<float(*)[8]>(&qb[8]));
#endif
      }
#endif
    };
    // rotate + scale + split the prefetched q and write it to TMEM:
    // columns [0,32) hi, [32,64) lo; two channels per 32-bit column
    constexpr int QS = Cfg::kQStages;
    auto stage_q = [&](int g) {   // g: the tile whose q is in the registers
      if (rope) {
        float c[P], sn[P];
#if NAF_TMA_EXP & 8
#pragma unroll
        for (int j = 0; j < P; ++j) { c[j] = 0.8f; sn[j] = 0.6f; }
#else
        if (tabs) {
          // the cell's table rows came with its K window (same buffer index, same barrier)
          const int kb = (q_flags >> 2) & 1;
          if (q_flags & 1) mbar_wait(&bar_k_full[kb], (q_flags >> 3) & 1);
          const float* tb = sTab + kb * (gm.tab_bytes / 4);
          // row tables: one 64-byte row per pixel row (a broadcast read); column tables: chunk planes
          const float* ct = half == 0 ? tb + q_r * P : tb + 2 * rh * P + q_c * 4;
          const float* st = ct + (half == 0 ? rh * P : gm.xchunk);              // xchunk bytes * 4 chunks = floats
          const int cs = half == 0 ? 4 : gm.xchunk / 4;                         // floats between 4-channel chunks
#pragma unroll
          for (int j = 0; j < P; j += 4) {
            *reinterpret_cast<float4*>(&c[j]) = *reinterpret_cast<const float4*>(ct + (j >> 2) * cs);
            *reinterpret_cast<float4*>(&sn[j]) = *reinterpret_cast<const float4*>(st + (j >> 2) * cs);
          }
        } else {
          const float* ct = half == 0 ? p.cos_y + int64_t(q_r) * P : p.cos_x + int64_t(q_c) * P;
          const float* st = half == 0 ? p.sin_y + int64_t(q_r) * P : p.sin_x + int64_t(q_c) * P;
          ldg8(ct, *reinterpret_cast<float(*)[8]>(&c[0]));
          ldg8(ct + 8, *reinterpret_cast<float(*)[8]>(&c[8]));
          ldg8(st, *reinterpret_cast<float(*)[8]>(&sn[0]));
          ldg8(st + 8, *reinterpret_cast<float(*)[8]>(&sn[8]));
        }
#endif
        const uint64_t qs2 = tm_pack2(qscale, qscale);
#pragma unroll
        for (int j = 0; j < P; j += 2) {
          const uint64_t a2 = tm_pack2(qa[j], qa[j + 1]), b2 = tm_pack2(qb[j], qb[j + 1]);
          const uint64_t c2 = tm_pack2(c[j], c[j + 1]), s2 = tm_pack2(sn[j], sn[j + 1]);
          const uint64_t ra = tm_fma2(b2 ^ 0x8000000080000000ull, s2, tm_fma2(a2, c2, 0ull));
          const uint64_t rb = tm_fma2(a2, s2, tm_fma2(b2, c2, 0ull));
          tm_unpack2(tm_fma2(ra, qs2, 0ull), qa[j], qa[j + 1]);
          tm_unpack2(tm_fma2(rb, qs2, 0ull), qb[j], qb[j + 1]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < P; ++j) {
          qa[j] *= qscale;
          qb[j] *= qscale;
        }
      }
      const uint32_t tq = tmem + Cfg::kTmemQ + (QS == 2 ? (g & 1) * 64 : 0) + lane_off;
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) split2_f16(qa[2 * j], qa[2 * j + 1], hi[j], lo[j]);
      tmem_st8(tq + 8 * half, hi);
      tmem_st8(tq + 32 + 8 * half, lo);
#pragma unroll
      for (int j = 0; j < 8; ++j) split2_f16(qb[2 * j], qb[2 * j + 1], hi[j], lo[j]);
      tmem_st8(tq + 16 + 8 * half, hi);
      tmem_st8(tq + 48 + 8 * half, lo);
      wait_st();
      fence_before_sync();
      mbar_arrive(&bar_q_full[QS == 2 ? (g & 1) : 0]);
      if (tabs && (q_flags & 2)) {
        // last tile of the item: this warp has read its table rows, the buffer may be refilled
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_k_free[(q_flags >> 2) & 1]);
      }
    };

    if (total_tiles > 0) {
      issue_q();
      stage_q(0);
      if (total_tiles > 1) issue_q();
    }
    for (int g = 0; g < total_tiles; ++g) {
      const int s = g & 1;
      // With one Q stage the prefetch of tile g+2 is issued after the softmax (its consumer, stage_q(g+2), comes
      // after the wait for S(g+1)).  With two stages stage_q(g+2) opens the next iteration, so the loads must be
      // in flight across the softmax (measured: issuing them late exposes the whole load latency once per tile).
      constexpr bool LATE_Q = QS == 1;
      if constexpr (QS == 2) {
        if (g + 1 < total_tiles) {
          stage_q(g + 1);
          if (!LATE_Q && g + 2 < total_tiles) issue_q();
        }
      }
      mbar_wait(&bar_s_full[s], (g >> 1) & 1);
      fence_after_sync();
      if constexpr (QS == 1) {
        if (g + 1 < total_tiles) stage_q(g + 1);   // one Q stage: QK(g) must have read it first
      }
      const uint32_t ts = tmem + Cfg::kTmemS + s * TP + lane_off;
      auto softmax_half = [&](auto half_c) {
        constexpr int H = decltype(half_c)::value;
        constexpr int base = H * SC;
        constexpr int NV = K2 - base < 0 ? 0 : (K2 - base > SC ? SC : K2 - base);   // valid taps of this half
        uint32_t mine[SC];
        if constexpr (SC % 16 == 0) {
#pragma unroll
          for (int c0 = 0; c0 < SC; c0 += 16) tmem_ld16(ts + base + c0, *reinterpret_cast<uint32_t(*)[16]>(&mine[c0]));
        } else {
#pragma unroll
          for (int c0 = 0; c0 < SC; c0 += 8) tmem_ld8(ts + base + c0, *reinterpret_cast<uint32_t(*)[8]>(&mine[c0]));
        }
        wait_ld();
        float m = -INFINITY;
#pragma unroll
        for (int j = 0; j < NV; ++j) m = fmaxf(m, __uint_as_float(mine[j]));
        if (p.scores != nullptr) {
          // return_weights=True (src/layers/attentions.py:27-28): the scaled pre-softmax logits, tap order
          // t_h*K + t_w.  S holds them times log2(e) (folded into the query scale for the exp2 softmax).
          const int it_s = tm_div(g, dv.ntiles), tile_s = g - it_s * ntiles;
          const TmItem its = item_of(it_s);
          const int pi = tile_s * tile_rows + row;
          if (pi < min(npix, (tile_s + 1) * tile_rows)) {
            const int py = tm_div(pi, dv.rw);
            const int y = its.ci * rh + py, x = its.cj * rw + (pi - py * rw);
            float* so = p.scores + (((int64_t(its.b) * p.heads + its.head) * p.Ho + y) * p.Wo + x) * K2 + base;
#pragma unroll
            for (int j = 0; j < NV; ++j) so[j] = __uint_as_float(mine[j]) * 0.6931471805599453f;
          }
        }
        // the two halves of a row exchange their partial maxima through shared memory; the same named
        // barrier orders the S reads of both halves before P overwrites the S columns
        float* mslot = mx + (s * 2) * 128 + row;
        mslot[H * 128] = m;
        fence_before_sync();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        fence_after_sync();
        m = fmaxf(m, mslot[(H ^ 1) * 128]);
        const uint64_t one2 = tm_pack2(1.f, 1.f), negm2 = tm_pack2(-m, -m);
        uint64_t l2 = 0ull;
        // e = 2^(s - m) in chunks of CH taps, each split into fp16 hi / lo and written back over S right away
        // (only one chunk of exponentials is ever live; the padded taps cost neither an exp nor a split)
        constexpr int CH = (SC % 16 == 0) ? 16 : 8;
#pragma unroll
        for (int c0 = 0; c0 < SC; c0 += CH) {
          uint32_t hi[CH / 2], lo[CH / 2];
#pragma unroll
          for (int j = 0; j < CH; j += 2) {
            const int t = c0 + j;
            if (t + 1 < NV) {
              float d0, d1;
              tm_unpack2(tm_fma2(tm_pack2(__uint_as_float(mine[t]), __uint_as_float(mine[t + 1])), one2, negm2), d0, d1);
#if NAF_TMA_EXP & 4
              const float e0 = 1.f + 0.f * d0, e1 = 1.f + 0.f * d1;
#else
              const float e0 = fast_exp2(d0), e1 = fast_exp2(d1);
#endif
              l2 = tm_fma2(tm_pack2(e0, e1), one2, l2);
              split2_f16(e0, e1, hi[j / 2], lo[j / 2]);
            } else if (t < NV) {
              const float e0 = fast_exp2(__uint_as_float(mine[t]) - m);
              l2 = tm_fma2(tm_pack2(e0, 0.f), one2, l2);
              split2_f16(e0, 0.f, hi[j / 2], lo[j / 2]);
            } else {
              hi[j / 2] = lo[j / 2] = 0u;
            }
          }
          if constexpr (CH == 16) {
            tmem_st8(ts + (base + c0) / 2, hi);
            tmem_st8(ts + TP / 2 + (base + c0) / 2, lo);
          } else {
            tmem_st4(ts + (base + c0) / 2, hi);
            tmem_st4(ts + TP / 2 + (base + c0) / 2, lo);
          }
        }
        float l0, l1;
        tm_unpack2(l2, l0, l1);
        // row sum for the epilogue, slot [tile & 3][half] (the epilogue may lag two tiles behind); published by
        // the release of the p_full arrive below, observed through the MMA thread's commit to o_full
        lsum[((g & 3) * 2 + H) * 128 + row] = l0 + l1;
      };
      if (half == 0) softmax_half(std::integral_constant<int, 0>{});
      else softmax_half(std::integral_constant<int, 1>{});
      wait_st();
      fence_before_sync();
      mbar_arrive(&bar_p_full[s]);   // release: l and P (TMEM) are visible to the consumers
      if constexpr (LATE_Q) {
        if (g + 2 < total_tiles) issue_q();
      }
    }
  } else if (warp < MMA_WARP) {
    // ======================================================================== BACK (two threads per row)
    const int bw = warp - 8;
    const int rowgrp = bw & 3, half = bw >> 2;
    const int row = rowgrp * 32 + lane;
    const uint32_t lane_off = uint32_t(rowgrp * 32) << 16;
    constexpr int RC = Cfg::kRoundCols;   // output columns staged per round
    constexpr int HC = RC / 2;            // ... of which this thread owns a contiguous half
    const bool bf16_out = p.out_dtype == NAF_DTYPE_BF16;
    const bool issuer = tid == NFRONT;
    if (issuer) {
      tmap::prefetch_desc(&tmO);
      tmap::prefetch_desc(&tmO2);
    }
#if NAF_TMA_STHINT == 1
    const uint64_t st_policy = tmap::policy_evict_first();
#elif NAF_TMA_STHINT == 2
    const uint64_t st_policy = tmap::policy_evict_last();
#endif
    int n = 0;      // accumulator uses so far: (tile, half) pairs
    int g = 0;
    int rc = 0;     // staging rounds so far
    for (int it_seq = 0; it_seq < my_items; ++it_seq) {
      const TmItem it = item_of(it_seq);
      for (int tile = 0; tile < ntiles; ++tile, ++g) {
        const bool last_tile = tile + 1 == ntiles;
        const int y0 = it.ci * rh + tile * gm.th, x0 = it.cj * rw;
        float inv_l = 0.f;
#pragma unroll
        for (int hh = 0; hh < NH; ++hh, ++n) {
          const int ob = NOB == 2 ? (n & 1) : 0;                 // accumulator buffer of this part
          const int opar = NOB == 2 ? ((n >> 1) & 1) : (n & 1);   // ... and the parity of this use
          const uint32_t tO = tmem + Cfg::kTmemO + ob * DVH + lane_off;
          mbar_wait(&bar_o_full[ob], opar);
          fence_after_sync();
          if (hh == 0) {
            inv_l = 1.f / (lsum[((g & 3) * 2) * 128 + row] + lsum[((g & 3) * 2 + 1) * 128 + row]);
          }
          // Accumulator halves of <= 128 columns (64 per thread) are pulled into registers in one go and O is
          // handed back to the tensor core BEFORE the staging rounds: with two halves per tile the next PV
          // waits for exactly this drain.  Wider accumulators are streamed 16 columns at a time per round.
          constexpr bool PRELOAD = DVH <= 128;
          constexpr int PC = PRELOAD ? DVH / 2 : 16;     // columns of this thread held at once
          uint32_t pre[PC];
          if constexpr (PRELOAD) {
            // this thread's columns of round rd: [rd*RC + half*HC, +HC)  ->  pre[rd*HC .. rd*HC+HC)
#pragma unroll
            for (int rd = 0; rd < ROUNDS; ++rd)
#pragma unroll
              for (int c = 0; c < HC; c += 16)
                tmem_ld16(tO + rd * RC + half * HC + c, *reinterpret_cast<uint32_t(*)[16]>(&pre[rd * HC + c]));
            wait_ld();
            fence_before_sync();
            mbar_arrive(&bar_o_free[ob]);
          }
#pragma unroll
          for (int rd = 0; rd < ROUNDS; ++rd, ++rc) {
            // Staging slot of this round.  With two slots nobody waits here: the stores that last read this
            // slot (round rc-2) were waited for by the issuer before the barrier of round rc-1.
            uint8_t* const slot = stage_out + (Cfg::NSTG == 2 ? (rc & 1) * Cfg::kSlot : 0);
            const uint32_t to = tO + rd * RC + half * HC;
            const uint64_t inv2 = tm_pack2(inv_l, inv_l);
            // normalise 16 accumulator columns and write them into the swizzled box images
            auto stage16 = [&](const uint32_t* r, int c) {
              if (bf16_out) {
                // 64 channels per 128-byte box row
#pragma unroll
                for (int j = 0; j < 16; j += 8) {
                  uint4 pk;
                  uint32_t* pw = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(r[j + 2 * e]) * inv_l,
                                                                    __uint_as_float(r[j + 2 * e + 1]) * inv_l);
                    pw[e] = *reinterpret_cast<const uint32_t*>(&h2);
                  }
                  const int col = half * HC + c + j;
                  *reinterpret_cast<uint4*>(slot + (col >> 6) * Cfg::kBox + tmap::swz128(row, (col & 63) >> 3)) = pk;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                  const uint64_t o01 = tm_fma2(tm_pack2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), inv2, 0ull);
                  const uint64_t o23 = tm_fma2(tm_pack2(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])), inv2, 0ull);
                  const int col = half * HC + c + j;
                  *reinterpret_cast<uint4*>(slot + (col >> 5) * Cfg::kBox + tmap::swz128(row, (col & 31) >> 2)) =
                      make_uint4(uint32_t(o01), uint32_t(o01 >> 32), uint32_t(o23), uint32_t(o23 >> 32));
                }
              }
            };
            if constexpr (PRELOAD) {
#pragma unroll
              for (int c = 0; c < HC; c += 16) stage16(&pre[rd * HC + c], c);
            } else {
              uint32_t r[2][16];
              tmem_ld16(to, r[0]);
#pragma unroll
              for (int c = 0; c < HC / 16; ++c) {
                wait_ld();
                if (c + 1 < HC / 16) tmem_ld16(to + (c + 1) * 16, r[(c + 1) & 1]);
                else if (rd == ROUNDS - 1) {
                  // the whole accumulator has left TMEM: the next PV may overwrite O
                  fence_before_sync();
                  mbar_arrive(&bar_o_free[ob]);
                }
                stage16(r[c & 1], c * 16);
              }
            }
            fence_proxy_async_smem();
            // every store issued so far has read its slot before anybody writes the OTHER slot (next round)
            if (Cfg::NSTG == 2 && issuer) bulk_wait_read<0>();
            asm volatile("bar.sync 2, 256;" ::: "memory");
            if (issuer) {
              const CUtensorMap* tm = (last_tile && gm.th_last != gm.th) ? &tmO2 : &tmO;
              const int c0 = it.head * Cfg::DV + hh * DVH + rd * RC;
              const int per_box = bf16_out ? 64 : 32;
              const int nbox = RC / per_box;
              if (!(NAF_TMA_EXP & 1))
                for (int j = 0; j < nbox; ++j) {
#if NAF_TMA_STHINT
                  tmap::store4_hint(tm, c0 + j * per_box, x0, y0, it.b, slot + j * Cfg::kBox, st_policy);
#else
                  tmap::store4(tm, c0 + j * per_box, x0, y0, it.b, slot + j * Cfg::kBox);
#endif
                }
              bulk_commit();
              if (Cfg::NSTG == 1) bulk_wait_read<0>();
            }
            if (Cfg::NSTG == 1) asm volatile("bar.sync 3, 256;" ::: "memory");   // single slot: free again
          }
        }
      }
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all output writes performed
  } else if (warp == LOAD_WARP) {
    // ======================================================================== WINDOW LOADER
    if (lane == 0) {
      tmap::prefetch_desc(&tmK);
      tmap::prefetch_desc(&tmV);
      if (gm.tab_bytes) {
        tmap::prefetch_desc(&tmCX);
        tmap::prefetch_desc(&tmSX);
      }
    }
    // Bulk L2 prefetch of the cell's query rows, one item ahead: the CTA of head 0 pulls the rows of ALL heads
    // (whole pixels, one contiguous run per guidance row), so that DRAM sees one sequential read burst per cell
    // instead of sector-sized reads scattered between the output writes; the per-tile register loads then hit L2.
    const bool qbulk = NAF_TMA_QBULK && p.q_stride_x == p.D && !(NAF_TMA_EXP & 2);
    const int qes = p.q_dtype == NAF_DTYPE_BF16 ? 2 : 4;
    for (int it_seq = 0; it_seq < my_items; ++it_seq) {
      const TmItem it = item_of(it_seq);
      if (lane == 0) {
        const int wy0 = window_origin(it.ci, p.h, K), wx0 = window_origin(it.cj, p.w, K);
        const int kb = it_seq & 1;
        if (it_seq >= 2) mbar_wait(&bar_k_free[kb], ((it_seq >> 1) - 1) & 1);
        mbar_expect_tx(&bar_k_full[kb], 2 * Cfg::kKPlane + (gm.tab_bytes ? 2 * (rh + rw) * (DQ / 4) * 4 : 0));
        if (gm.tab_bytes) {
          float* tb = const_cast<float*>(sTab) + kb * (gm.tab_bytes / 4);
          constexpr int P = DQ / 4;
          bulk_load(tb, p.cos_y + int64_t(it.ci) * rh * P, rh * P * 4, &bar_k_full[kb]);
          bulk_load(tb + rh * P, p.sin_y + int64_t(it.ci) * rh * P, rh * P * 4, &bar_k_full[kb]);
          float* tx = tb + 2 * rh * P;
          const int xc = gm.xchunk / 4;   // floats per chunk plane
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            tmap::load4(tx + c * xc, &tmCX, 4 * c, it.cj * rw, 0, 0, &bar_k_full[kb]);
            tmap::load4(tx + (4 + c) * xc, &tmSX, 4 * c, it.cj * rw, 0, 0, &bar_k_full[kb]);
          }
        }
        uint8_t* kdst = sK + kb * Cfg::kKWin;
        const int gk = (it.b * 2) * gm.groups_k + it.head * KC;
        tmap::load4(kdst, &tmK, 0, wx0, wy0, gk, &bar_k_full[kb]);
        tmap::load4(kdst + Cfg::kKStride, &tmK, 0, wx0, wy0, gk + gm.groups_k, &bar_k_full[kb]);
      }
      __syncwarp();
      if (qbulk && it.head == 0) {
        const int ey0 = tm_div(it.ci * rh, dv.rep_y), ey1 = tm_div(it.ci * rh + rh - 1, dv.rep_y);
        const int ex0 = tm_div(it.cj * rw, dv.rep_x), ex1 = tm_div(it.cj * rw + rw - 1, dv.rep_x);
        const uint32_t bytes = uint32_t(ex1 - ex0 + 1) * uint32_t(p.D) * qes;
        const uint8_t* base = reinterpret_cast<const uint8_t*>(p.q) +
                              (int64_t(it.b) * p.q_stride_b + int64_t(ex0) * p.q_stride_x) * qes;
        for (int ey = ey0 + lane; ey <= ey1; ey += 32) bulk_prefetch_l2(base + int64_t(ey) * p.q_stride_y * qes, bytes);
      }
      if (lane == 0) {
        const int wy0 = window_origin(it.ci, p.h, K), wx0 = window_origin(it.cj, p.w, K);
        const int vb = Cfg::NVB == 2 ? (it_seq & 1) : 0;
        const int vuse = Cfg::NVB == 2 ? (it_seq >> 1) : it_seq;     // how often this buffer has been filled before
        if (vuse >= 1) mbar_wait(&bar_v_free[vb], (vuse - 1) & 1);
        mbar_expect_tx(&bar_v_full[vb], 2 * Cfg::kVPlane);
        uint8_t* vdst = sV + vb * Cfg::kVWin;
        const int gv = (it.b * 2) * gm.groups_v + it.head * (Cfg::DV / 8);
        tmap::load4(vdst, &tmV, 0, wx0, wy0, gv, &bar_v_full[vb]);
        tmap::load4(vdst + Cfg::kVStride, &tmV, 0, wx0, wy0, gv + gm.groups_v, &bar_v_full[vb]);
      }
      __syncwarp();
    }
  } else {
    // ======================================================================== MMA ISSUER
#ifndef NAF_TMA_ELECT
// 1: issue under elect.sync (descriptors in uniform registers: back-to-back UTCHMMA, what the backward kernel uses).
// Measured here (session 4, same box): C2 unchanged (4.29 / 4.34 vs 4.29 / 4.31 ms), C3 SLOWER (5.99 / 6.06 vs 5.59 /
// 5.63 ms: the faster issuer runs the next tile's Q K^T ahead of the P V the drain is waiting for).  Off.
#define NAF_TMA_ELECT 0
#endif
    __syncwarp();
    if (NAF_TMA_ELECT ? elect_one_sync() : lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc_f16(128, TP, false, false);
      constexpr uint32_t idesc_pv = make_idesc_f16(128, DVH, false, true);
      constexpr int QS = Cfg::kQStages;
      int qk_seq = 0, qk_tile = 0, pv_seq = 0, pv_tile = 0;   // (item, tile) of the next QK / PV
      int n = 0;                                              // accumulator uses issued so far
      auto issue_qk = [&](int g) {
        const int s = g & 1;
        const int it_seq = qk_seq;
        const int kb = it_seq & 1;
        if (qk_tile == 0) {
          mbar_wait(&bar_k_full[kb], (it_seq >> 1) & 1);
        }
        const uint8_t* w = sK + kb * Cfg::kKWin;
        const bool last_of_item = qk_tile + 1 == ntiles;
        if (++qk_tile == ntiles) { qk_tile = 0; ++qk_seq; }
        if constexpr (QS == 2) mbar_wait(&bar_q_full[s], (g >> 1) & 1);
        else mbar_wait(&bar_q_full[0], g & 1);
        fence_after_sync();
        const uint32_t tq = tmem + Cfg::kTmemQ + (QS == 2 ? s * 64 : 0);
        const uint32_t tS = tmem + Cfg::kTmemS + s * TP;
        // S = Qhi*Khi^T + Qlo*Khi^T + Qhi*Klo^T
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t a0 = tq + (pass == 1 ? 32 : 0);
          const uint32_t b0 = smem_u32(w + (pass == 2 ? Cfg::kKStride : 0));
#pragma unroll
          for (int kk = 0; kk < DQ / 16; ++kk) {
            // K-major B: 16 channels = 2 channel groups per MMA; LBO = one channel-group plane, SBO = 8 taps
            const uint64_t db = make_desc(b0 + kk * 2 * (K2 * 16), K2 * 16, 128);
            mma_f16_ts(tS, a0 + kk * 8, db, idesc_qk, (pass | kk) != 0);
          }
        }
        commit(&bar_s_full[s]);
        if (last_of_item) commit(&bar_k_free[kb]);   // every MMA reading this K window has been issued
      };
#ifndef NAF_TMA_QK_MID
#define NAF_TMA_QK_MID 0   // 1: value heads in several accumulator halves issue the next tile's Q K^T after the FIRST half
#endif                     //    (it then runs while that half drains) instead of before the whole P V
      auto issue_pv = [&](int g, bool qk_mid) {
        const int s = g & 1;
        mbar_wait(&bar_p_full[s], (g >> 1) & 1);
        const int vb = Cfg::NVB == 2 ? (pv_seq & 1) : 0;
        const int vuse = Cfg::NVB == 2 ? (pv_seq >> 1) : pv_seq;
        if (pv_tile == 0) mbar_wait(&bar_v_full[vb], vuse & 1);
        const uint8_t* w = sV + vb * Cfg::kVWin;
        const bool last_of_item = pv_tile + 1 == ntiles;
        if (++pv_tile == ntiles) { pv_tile = 0; ++pv_seq; }
        const uint32_t tP = tmem + Cfg::kTmemS + s * TP;
#pragma unroll
        for (int hh = 0; hh < NH; ++hh, ++n) {
          const int ob = NOB == 2 ? (n & 1) : 0;
          const uint32_t tO = tmem + Cfg::kTmemO + ob * DVH;
          mbar_wait(&bar_o_free[ob], ((NOB == 2 ? (n >> 1) : n) + 1) & 1);   // drained by the epilogue of the previous use
          fence_after_sync();
          // O = Phi*Vhi + Plo*Vhi + Phi*Vlo on the channel groups of this half
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            if ((NAF_TMA_EXP & 16) && pass == 1) continue;   // profiling: two-pass P V
            const uint32_t a0 = tP + (pass == 1 ? TP / 2 : 0);
            const uint32_t b0 = smem_u32(w + (pass == 2 ? Cfg::kVStride : 0) + hh * (DVH / 8) * (K2 * 16));
#pragma unroll
            for (int kk = 0; kk < TP / 16; ++kk) {
              // MN-major B: 16 taps = 2 groups of 8 taps per MMA; LBO = 8 taps, SBO = one channel-group plane
              const uint64_t db = make_desc(b0 + kk * 256, 128, K2 * 16);
              mma_f16_ts(tO, a0 + kk * 8, db, idesc_pv, (pass | kk) != 0);
            }
          }
          commit(&bar_o_full[ob]);
          if (qk_mid && hh == 0) issue_qk(g + 1);
        }
        if (last_of_item) commit(&bar_v_free[vb]);
      };
      if (total_tiles > 0) issue_qk(0);
      constexpr bool kMid = NAF_TMA_QK_MID && NH >= 2;
      for (int g = 0; g < total_tiles; ++g) {
        if (!kMid && g + 1 < total_tiles) issue_qk(g + 1);
        issue_pv(g, kMid && g + 1 < total_tiles);
      }
    }
    __syncwarp();
  }

  fence_before_sync();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------ host side
namespace {

int tm_taps_pad(int K) { return (K * K + 15) / 16 * 16; }

TmFastDiv tm_make_div(uint32_t d) {
  TmFastDiv f;
  f.d = d;
  uint32_t s = 0;
  while ((1u << s) < d) ++s;
  if ((d & (d - 1)) == 0) {
    f.magic = 0;
    f.shift = s;
  } else {
    f.magic = uint32_t(((uint64_t(1) << (31 + s)) + d - 1) / d);
    f.shift = s - 1;
  }
  return f;
}

// tile geometry: whole cell rows, balanced
void tm_tiles(int rh, int rw, int& th, int& ntiles, int& th_last) {
  th = 128 / rw;
  if (th < 1) th = 1;
  if (th > rh) th = rh;
  ntiles = (rh + th - 1) / th;
  th = (rh + ntiles - 1) / ntiles;
  ntiles = (rh + th - 1) / th;
  th_last = rh - (ntiles - 1) * th;
}

size_t tm_workspace_bytes(const naf_xattn_params& p) {
  // fp16 hi + lo planes of K and V: the same byte count as the fp32 maps
  return size_t(p.B) * p.h * p.w * (size_t(p.D) + p.C) * 4 + 256;
}

template <int TP, int DVH, int NH, int ROUNDS, int NOB>
int launch_tma(const naf_xattn_params& p, cudaStream_t st) {
  using Cfg = TmaCfg<TP, DVH, NH, ROUNDS, NOB>;
  if constexpr (!Cfg::kFits) {
    return fail(NAF_ERR_UNSUPPORTED, "xattn(cell-tma): tile %dx%dx%d/%d/%d does not fit", TP, DVH, NH, ROUNDS, NOB);
  } else {
    auto kern = xattn_cell_tma_kernel<TP, DVH, NH, ROUNDS, NOB>;
    const int sms = device_sm_count();
    TmGeom gm;
    gm.rh = p.Ho / p.h;
    gm.rw = p.Wo / p.w;
    tm_tiles(gm.rh, gm.rw, gm.th, gm.ntiles, gm.th_last);
    const int64_t items = int64_t(p.B) * p.h * p.w * p.heads;
    gm.n_items = int(items);
    gm.groups_k = p.D / 8;
    gm.groups_v = p.C / 8;
    // RoPE table rows of a cell in shared memory (one set per K window buffer) when there is room for them
    gm.xchunk = (gm.rw * 16 + 127) / 128 * 128;
    gm.tab_bytes = p.cos_y ? 2 * gm.rh * (DQ / 4) * 4 + 8 * gm.xchunk : 0;
    if (Cfg::kSmemBytes + Cfg::NKB * gm.tab_bytes > kSmemLimit) gm.tab_bytes = 0;
    const int smem_bytes = Cfg::kSmemBytes + Cfg::NKB * gm.tab_bytes;
    const cudaError_t e = ensure_dyn_smem(kern, smem_bytes);
    if (e != cudaSuccess)
      return fail(NAF_ERR_CUDA, "xattn(cell-tma): smem opt-in failed: %s", cudaGetErrorString(e));
    TmDivs dv;
    dv.ntiles = tm_make_div(uint32_t(gm.ntiles));
    dv.heads = tm_make_div(uint32_t(p.heads));
    dv.w = tm_make_div(uint32_t(p.w));
    dv.h = tm_make_div(uint32_t(p.h));
    dv.rw = tm_make_div(uint32_t(gm.rw));
    dv.rep_y = tm_make_div(uint32_t(p.rep_y));
    dv.rep_x = tm_make_div(uint32_t(p.rep_x));

    // ---- pre-pass: K and V -> fp16 hi / lo channel-group planes in the caller's workspace
    uint8_t* ws = static_cast<uint8_t*>(p.workspace);
    ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 127) & ~uintptr_t(127));
    uint8_t* kplanes = ws;
    uint8_t* vplanes = ws + size_t(p.B) * p.h * p.w * p.D * 4;
    prefer_max_shared(kv_planes_kernel<float>);
    prefer_max_shared(kv_planes_kernel<__nv_bfloat16>);
    {
      const int64_t nk = int64_t(p.B) * (p.D / 8) * p.h * p.w, nv = int64_t(p.B) * (p.C / 8) * p.h * p.w;
      const int64_t cap = int64_t(sms) * 16;
      const int64_t bk_ = (nk + 255) / 256, bv_ = (nv + 255) / 256;
      const unsigned gk_ = unsigned(bk_ < cap ? bk_ : cap), gv_ = unsigned(bv_ < cap ? bv_ : cap);
      if (p.k_dtype == NAF_DTYPE_BF16)
        kv_planes_kernel<<<gk_, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(p.k), reinterpret_cast<uint4*>(kplanes), p.B, p.h, p.w, p.D);
      else
        kv_planes_kernel<<<gk_, 256, 0, st>>>(p.k, reinterpret_cast<uint4*>(kplanes), p.B, p.h, p.w, p.D);
      if (p.v_dtype == NAF_DTYPE_BF16)
        kv_planes_kernel<<<gv_, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(p.v), reinterpret_cast<uint4*>(vplanes), p.B, p.h, p.w, p.C);
      else
        kv_planes_kernel<<<gv_, 256, 0, st>>>(p.v, reinterpret_cast<uint4*>(vplanes), p.B, p.h, p.w, p.C);
      int rc = check_launch("xattn_kv_planes");
      if (rc != NAF_OK) return rc;
    }
    // ---- tensor maps
    CUtensorMap tmK, tmV, tmO, tmO2, tmCX, tmSX;
    memset(&tmCX, 0, sizeof(tmCX));
    memset(&tmSX, 0, sizeof(tmSX));
    if (gm.tab_bytes) {
      // column tables (Wo, 16) fp32: box = 4 channels x rw columns, one load per chunk plane
      const uint64_t dt_[4] = {16, uint64_t(p.Wo), 1, 1};
      const uint64_t st_[3] = {64, uint64_t(p.Wo) * 64, uint64_t(p.Wo) * 64};
      const uint32_t bt_[4] = {4, uint32_t(gm.rw), 1, 1};
      if (!tmap::encode4(&tmCX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, p.cos_x, dt_, st_, bt_, CU_TENSOR_MAP_SWIZZLE_NONE) ||
          !tmap::encode4(&tmSX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, p.sin_x, dt_, st_, bt_, CU_TENSOR_MAP_SWIZZLE_NONE))
        return fail(NAF_ERR_CUDA, "xattn(cell-tma): cuTensorMapEncodeTiled failed (rope tables)");
    }
    {
      const uint64_t dk[4] = {8, uint64_t(p.w), uint64_t(p.h), uint64_t(p.B) * 2 * (p.D / 8)};
      const uint64_t dvv[4] = {8, uint64_t(p.w), uint64_t(p.h), uint64_t(p.B) * 2 * (p.C / 8)};
      const uint64_t sp[3] = {16, uint64_t(p.w) * 16, uint64_t(p.h) * p.w * 16};
      const uint32_t bk[4] = {8, uint32_t(Cfg::K), uint32_t(Cfg::K), uint32_t(KC)};
      const uint32_t bv[4] = {8, uint32_t(Cfg::K), uint32_t(Cfg::K), uint32_t(Cfg::DV / 8)};
      const bool bf16 = p.out_dtype == NAF_DTYPE_BF16;
      const uint64_t es = bf16 ? 2 : 4;
      const uint64_t d_o[4] = {uint64_t(p.C), uint64_t(p.Wo), uint64_t(p.Ho), uint64_t(p.B)};
      const uint64_t so[3] = {uint64_t(p.C) * es, uint64_t(p.Wo) * p.C * es, uint64_t(p.Ho) * p.Wo * p.C * es};
      const uint32_t bo[4] = {bf16 ? 64u : 32u, uint32_t(gm.rw), uint32_t(gm.th), 1};
      const uint32_t bo2[4] = {bf16 ? 64u : 32u, uint32_t(gm.rw), uint32_t(gm.th_last), 1};
      const CUtensorMapDataType odt = bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
      if (!tmap::encode4(&tmK, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, kplanes, dk, sp, bk, CU_TENSOR_MAP_SWIZZLE_NONE) ||
          !tmap::encode4(&tmV, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, vplanes, dvv, sp, bv, CU_TENSOR_MAP_SWIZZLE_NONE) ||
          !tmap::encode4(&tmO, odt, p.out, d_o, so, bo, CU_TENSOR_MAP_SWIZZLE_128B) ||
          !tmap::encode4(&tmO2, odt, p.out, d_o, so, bo2, CU_TENSOR_MAP_SWIZZLE_128B))
        return fail(NAF_ERR_CUDA, "xattn(cell-tma): cuTensorMapEncodeTiled failed");
    }
    const int grid = int(items < sms ? items : sms);
    kern<<<grid, NTHREADS, smem_bytes, st>>>(p, gm, dv, tmK, tmV, tmO, tmO2, tmCX, tmSX);
    return check_launch("xattn_cell_tma");
  }
}

// (value head dim) -> (DVH, NH, ROUNDS) per padded tap count; 0 = no configuration
struct TmPlan {
  int dvh = 0, nh = 0, rounds = 0, nob = 1;
};

template <int TP>
constexpr TmPlan tm_plan(int dv, bool bf16) {
  // whole head in one accumulator when TMEM allows, else two halves against one P; rounds chosen so that two
  // staging slots fit (bf16 stores need 64-channel boxes, i.e. rounds of a multiple of 64 columns)
  TmPlan pl;
  switch (dv) {
    case 32: if (!bf16 && TmaCfg<TP, 32, 1, 1>::kFits) pl = {32, 1, 1}; break;
    case 64: if (TmaCfg<TP, 64, 1, 1>::kFits) pl = {64, 1, 1}; break;
    case 96: if (!bf16 && TmaCfg<TP, 96, 1, 1>::kFits) pl = {96, 1, 1}; break;
    case 128: if (TmaCfg<TP, 128, 1, 2>::kFits) pl = {128, 1, 2}; break;
    // (NAF_TMA_NOB == 2: two accumulator buffers of narrower parts -- a measured-slower build option, see above)
    case 192:
      if (bf16) { if (TmaCfg<TP, 192, 1, 1>::kFits) pl = {192, 1, 1, 1}; }
      else if (NAF_TMA_NOB == 2 && TmaCfg<TP, 96, 2, 1, 2>::kFits) pl = {96, 2, 1, 2};
      else if (TmaCfg<TP, 192, 1, 2>::kFits) pl = {192, 1, 2, 1};
      break;
    case 256:
      if (!bf16 && NAF_TMA_NOB == 2 && TmaCfg<TP, 64, 4, 2, 2>::kFits) pl = {64, 4, 2, 2};
      else if (TmaCfg<TP, 256, 1, 4>::kFits) pl = {256, 1, 4, 1};
      else if (bf16) { if (TmaCfg<TP, 128, 2, 2>::kFits) pl = {128, 2, 2, 1}; }
      else if (TmaCfg<TP, 128, 2, 4>::kFits) pl = {128, 2, 4, 1};
      break;
    default: break;
  }
  return pl;
}

TmPlan tm_plan_for(int K, int dv, bool bf16) {
  switch (tm_taps_pad(K)) {
    case 16: return tm_plan<16>(dv, bf16);
    case 32: return tm_plan<32>(dv, bf16);
    case 64: return tm_plan<64>(dv, bf16);
    case 96: return tm_plan<96>(dv, bf16);
    case 128: return tm_plan<128>(dv, bf16);
    default: return TmPlan{};
  }
}

template <int TP, int DVFULL, bool BF16>
int launch_tma_plan(const naf_xattn_params& p, cudaStream_t st) {
  constexpr TmPlan pl = tm_plan<TP>(DVFULL, BF16);
  if constexpr (pl.dvh == 0) {
    return fail(NAF_ERR_UNSUPPORTED, "xattn(cell-tma): no plan for dv=%d", DVFULL);
  } else {
    return launch_tma<TP, pl.dvh, pl.nh, pl.rounds, pl.nob>(p, st);
  }
}

template <int TP, int DVFULL>
int launch_tma_dv(const naf_xattn_params& p, cudaStream_t st) {
  return p.out_dtype == NAF_DTYPE_BF16 ? launch_tma_plan<TP, DVFULL, true>(p, st) : launch_tma_plan<TP, DVFULL, false>(p, st);
}

template <int TP>
int launch_tma_tp(const naf_xattn_params& p, cudaStream_t st) {
  switch (p.C / p.heads) {
    case 32: return launch_tma_dv<TP, 32>(p, st);
    case 64: return launch_tma_dv<TP, 64>(p, st);
    case 96: return launch_tma_dv<TP, 96>(p, st);
    case 128: return launch_tma_dv<TP, 128>(p, st);
    case 192: return launch_tma_dv<TP, 192>(p, st);
    case 256: return launch_tma_dv<TP, 256>(p, st);
    default: return fail(NAF_ERR_UNSUPPORTED, "xattn(cell-tma): no plan for dv=%d", p.C / p.heads);
  }
}

}  // namespace

size_t xattn_cell_tma_workspace(const naf_xattn_params& p) { return tm_workspace_bytes(p); }

bool xattn_cell_tma_supported(const naf_xattn_params& p, const char** why) {
  if (p.Kw != 0 && p.Kw != p.K) { *why = "rectangular window"; return false; }
  const int dq = p.D / p.heads, dv = p.C / p.heads;
  if (p.row_tap || p.col_tap) { *why = "tap tables given (non-integer ratio path)"; return false; }
  if (p.Ho % p.h || p.Wo % p.w) { *why = "target size is not a multiple of the feature size"; return false; }
  if (dq != DQ) { *why = "head dim must be 64"; return false; }
  if (p.K < 3 || p.K > 11) { *why = "kernel_size must be 3, 5, 7, 9 or 11"; return false; }
  if (p.h < p.K || p.w < p.K) { *why = "feature map smaller than the window"; return false; }
  const TmPlan pl = tm_plan_for(p.K, dv, p.out_dtype == NAF_DTYPE_BF16);
  if (!pl.dvh) { *why = "no tile configuration for this window / value head dim / output type"; return false; }
  const int rh = p.Ho / p.h, rw = p.Wo / p.w;
  if (rw > 128) { *why = "cells wider than 128 pixels"; return false; }
  if (rh * rw < 64) { *why = "fewer than 64 pixels per cell"; return false; }
  if (p.out_dtype == NAF_DTYPE_BF16 && p.C % 8) { *why = "bf16 store needs C % 8 == 0"; return false; }
  if (!p.workspace || size_t(p.workspace_bytes) < tm_workspace_bytes(p)) { *why = "workspace missing or too small (naf_xattn_workspace_bytes)"; return false; }
  const int qa_ = p.q_dtype == NAF_DTYPE_BF16 ? 16 : 8;   // 32-byte aligned query half rows for either element type
  if (!aligned32(p.q) || !aligned16(p.k) || !aligned16(p.v) || !aligned16(p.out) || (p.C % 8) ||
      (p.q_stride_b % qa_) || (p.q_stride_y % qa_) || (p.q_stride_x % qa_)) {
    *why = "pointers/strides not aligned";
    return false;
  }
  if ((p.k_dtype == NAF_DTYPE_F32 && !aligned32(p.k)) || (p.v_dtype == NAF_DTYPE_F32 && !aligned32(p.v))) {
    *why = "fp32 k / v not 32-byte aligned";
    return false;
  }
  if (p.cos_y && !(aligned32(p.cos_y) && aligned32(p.sin_y) && aligned32(p.cos_x) && aligned32(p.sin_x))) {
    *why = "rope tables not 32-byte aligned";
    return false;
  }
  if (int64_t(p.B) * p.h * p.w * p.heads >= (int64_t(1) << 31)) { *why = "too many items"; return false; }
  if (int64_t(p.B) * 2 * (p.C / 8) >= (int64_t(1) << 31)) { *why = "too many planes"; return false; }
  return true;
}

int launch_xattn_cell_tma(const naf_xattn_params& p, cudaStream_t st) {
  switch (tm_taps_pad(p.K)) {
    case 16: return launch_tma_tp<16>(p, st);
    case 32: return launch_tma_tp<32>(p, st);
    case 64: return launch_tma_tp<64>(p, st);
    case 96: return launch_tma_tp<96>(p, st);
    case 128: return launch_tma_tp<128>(p, st);
    default: return fail(NAF_ERR_UNSUPPORTED, "xattn(cell-tma): kernel_size %d", p.K);
  }
}

}  // namespace naf
