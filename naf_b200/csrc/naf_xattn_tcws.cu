// Cell kernel, warp-specialised tensor-core pipeline (tcgen05 / TMEM / bulk-copy epilogue).
//
// Same math as naf_xattn_tc.cu (two 3-pass fp16 hi/lo GEMMs per 128-row tile of one low-res cell
// and head, fp32 accumulation in TMEM), restructured so that HBM loads, tensor-core work, softmax
// and HBM stores of DIFFERENT tiles overlap inside one persistent CTA per SM:
//
//   warps 0-7   "front" : two threads per pixel row.  q half-rows HBM -> registers (prefetched two
//                          tiles ahead), RoPE, fp16 hi/lo split, tcgen05.st into TMEM (Q is the
//                          A operand of S = Q Kwin^T straight from TMEM, no smem staging);
//                          softmax on S (tcgen05.ld), P back to TMEM over S.
//   warps 8-15  "back"  : drain O from TMEM, normalise, stage the output row slabs in shared
//                          memory and hand them to the TMA engine (cp.async.bulk shared->global):
//                          output stores cost no LSU wavefronts and never stall a warp.
//   warp 16     "mma"   : one thread issues every tcgen05.mma in the order
//                          QK(0) QK(1) PV(0) QK(2) PV(1) ...  and commits to mbarriers.
//   warps 17-18 "stage" : convert the NEXT item's K / V windows (fp32 -> fp16 hi/lo, UMMA layouts)
//                          into the second window buffer -- for long items / large windows; otherwise
//                          the back group does it between two items and these warps idle.
//
// An "item" is (batch, cell, head, value-slab): wide value heads (dv = 256 with an 11x11 window)
// are split into `vsplit` slabs of DV channels so that windows and accumulators fit; the slabs of
// one head recompute the (cheap) scores.  TMEM: Q x 64 | S/P[2] x TP | O x DV | l x 8 columns.
// Tiles are balanced (ceil(npix / ntiles) rows each) so a 28x28 cell is 7 x 112 rows.
//
// Reference semantics: src/layers/attentions.py:16-29,53-75; RoPE src/layers/rope.py:137-153.
#include <cuda_bf16.h>

#include <type_traits>

#include "naf_common.cuh"
#include "naf_umma.cuh"

namespace naf {

using namespace umma;

namespace {

constexpr int DQ = 64;
constexpr int KC = DQ / 8;
constexpr int NFRONT = 256, NBACK = 256, NSTAGE = 64, NTHREADS = NFRONT + NBACK + 32 + NSTAGE;
constexpr int MMA_WARP = (NFRONT + NBACK) / 32;
#ifndef NAF_WS_EPI
#define NAF_WS_EPI 0     // output stores: 0 = one bulk copy (TMA engine) per thread and half row; 1 = per-warp transposed
#endif                   // read-back + 128-byte-run st.global (measured 11 % slower at C2: 5.16 vs 4.63 ms)
#ifndef NAF_WS_DB
#define NAF_WS_DB 0      // (measured slower: 5.25 vs 4.60 ms at C2) two-round epilogues alternate between two staging slabs per row (a bulk copy may still be
#endif                   // reading slab A while the thread refills slab B)
#ifndef NAF_WS_BLOCKED
#define NAF_WS_BLOCKED 0 // item -> CTA map: 0 = round robin (the 4 heads of a cell run on neighbouring SMs at the same time), 1 = each CTA
#endif                   // walks a contiguous range of items (measured slower: 5.0 vs 4.6 ms at C2)
#ifndef NAF_WS_EXP
#define NAF_WS_EXP 0     // profiling variants: 1 = no output stores, 2 = no q loads, 4 = no window staging after the first item, 8 = 16-byte stores
#endif

template <int TP>
struct WsWindowOf { static constexpr int K = TP == 16 ? 3 : TP == 32 ? 5 : TP == 64 ? 7 : TP == 96 ? 9 : 11; };

// TP: padded taps; DV: value channels per item; ROUNDS: staging rounds of the epilogue
template <int TP, int DV, int ROUNDS>
struct WsCfg {
  static constexpr int kSmemK = KC * TP * 16;          // one of hi / lo
  static constexpr int kSmemV = TP * DV * 2;           // one of hi / lo
  static constexpr int kWin = 2 * kSmemK + 2 * kSmemV; // one window buffer (K hi|lo|V hi|lo)
  static constexpr int kRoundCols = DV / ROUNDS;       // output columns staged per round
  // epilogues of exactly two rounds keep both rounds' slabs side by side in the row (double buffer)
  static constexpr bool kDoubleStage = NAF_WS_DB && ROUNDS == 2;
  static constexpr int kSlabs = kDoubleStage ? 2 : 1;
  static constexpr int kRowBytes = kSlabs * kRoundCols * 4 + 16;  // staged row + 16 B pad (bank spread)
  static constexpr int kStage = 128 * kRowBytes;
  static constexpr int kSmemTotal = 2 * kWin + kStage;
  // Q: 64 columns per stage (32 hi + 32 lo); double buffered when TMEM allows, so that the next
  // tile's Q is staged while the tensor core still works on QK of the current one
  static constexpr int kQStages = (128 + 2 * TP + DV + 8 <= 512) ? 2 : 1;
  static constexpr int kTmemQ = 0;
  static constexpr int kTmemS = 64 * kQStages;         // S/P[2]: 2 x TP columns
  static constexpr int kTmemO = kTmemS + 2 * TP;       // O: DV columns
  static constexpr int kTmemL = kTmemO + DV;           // row sums l: 4 slots (tile & 3) x 2 halves
  static constexpr int kTmemUsed = kTmemL + 8;
  static constexpr bool kFits = kTmemUsed <= 512 && kSmemTotal + 1024 <= 227 * 1024 &&
                                (DV % ROUNDS) == 0 && (kRoundCols % 32) == 0;
};

__device__ __forceinline__ void ws_split8(const float* x, uint4& hi, uint4& lo) {
  split2_f16(x[0], x[1], hi.x, lo.x);
  split2_f16(x[2], x[3], hi.y, lo.y);
  split2_f16(x[4], x[5], hi.z, lo.z);
  split2_f16(x[6], x[7], hi.w, lo.w);
}

// packed fp32x2 helpers (FFMA2): two floats in one 64-bit register
__device__ __forceinline__ uint64_t ws_pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void ws_unpack2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t ws_fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

// Division by a launch-constant divisor as multiply-high + shift (exact for the small operands
// used here: n < 2^31 / d, d <= 2^15); the table lives in the kernel parameter constant bank, so
// the per-tile index arithmetic of 272 threads costs no registers and ~3 instructions per divide.
struct FastDiv {
  uint32_t d, magic, shift;
};
struct WsDivs {
  FastDiv ntiles, vsplit, heads, w, h, rw, rep_y, rep_x;
};
__device__ __forceinline__ int fdiv(int n, const FastDiv& f) {
  return f.magic == 0 ? (n >> f.shift) : int(__umulhi(uint32_t(n), f.magic) >> f.shift);
}

struct ItemCoord {
  int b, ci, cj, head, vh;
};
__device__ __forceinline__ ItemCoord decode_item(int item, const WsDivs& dv) {
  ItemCoord c;
  int q = fdiv(item, dv.vsplit);
  c.vh = item - q * int(dv.vsplit.d);
  item = q;
  q = fdiv(item, dv.heads);
  c.head = item - q * int(dv.heads.d);
  item = q;
  q = fdiv(item, dv.w);
  c.cj = item - q * int(dv.w.d);
  item = q;
  q = fdiv(item, dv.h);
  c.ci = item - q * int(dv.h.d);
  c.b = q;
  return c;
}

// Stage the K and V windows of one item into a window buffer (fp32 -> fp16 hi/lo, UMMA
// canonical layouts).  Executed by `nthreads` threads with linear id `t`.  The global loads of U
// elements are issued back to back before the first conversion, so their latencies overlap.
template <int TP, int DV, int ROUNDS>
__device__ __forceinline__ void stage_windows(uint8_t* win, const naf_xattn_params& p,
                                              const ItemCoord& it, int vchan0, int t, int nthreads) {
  using Cfg = WsCfg<TP, DV, ROUNDS>;
  constexpr int K = WsWindowOf<TP>::K, K2 = K * K;
  constexpr int U = 4;
  uint8_t* sKhi = win;
  uint8_t* sKlo = sKhi + Cfg::kSmemK;
  uint8_t* sVhi = sKlo + Cfg::kSmemK;
  uint8_t* sVlo = sVhi + Cfg::kSmemV;
  const int wy0 = window_origin(it.ci, p.h, K);
  const int wx0 = window_origin(it.cj, p.w, K);
  const float* kbase = p.k + (int64_t(it.b * p.h + wy0) * p.w + wx0) * p.D + it.head * DQ;
  const float* vbase = p.v + (int64_t(it.b * p.h + wy0) * p.w + wx0) * p.C + vchan0;
  // K: canonical K-major [chunk c][tap n][16 B], zero rows for n >= K2
  for (int i0 = t; i0 < TP * KC; i0 += U * nthreads) {
    float x[U][8];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * nthreads;
      const int n = i % TP, c = i / TP;
      const bool ok = i < TP * KC && n < K2;
      const int tt = n / K, uu = n - tt * K;
      ldg8(ok ? kbase + (int64_t(tt) * p.w + uu) * p.D + c * 8 : kbase, x[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * nthreads;
      if (i < TP * KC) {
        const int n = i % TP, c = i / TP;
        uint4 hi = make_uint4(0, 0, 0, 0), lo = hi;
        if (n < K2) ws_split8(x[u], hi, lo);
        *reinterpret_cast<uint4*>(sKhi + (c * TP + n) * 16) = hi;
        *reinterpret_cast<uint4*>(sKlo + (c * TP + n) * 16) = lo;
      }
    }
  }
  // V: canonical MN-major [tap group][channel group][tap%8][16 B]
  constexpr int NG = DV / 8;
  for (int i0 = t; i0 < TP * NG; i0 += U * nthreads) {
    float x[U][8];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * nthreads;
      const int k = ((i >> 3) / NG) * 8 + (i & 7), g = (i >> 3) % NG;
      const bool ok = i < TP * NG && k < K2;
      const int tt = k / K, uu = k - tt * K;
      ldg8(ok ? vbase + (int64_t(tt) * p.w + uu) * p.C + g * 8 : vbase, x[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * nthreads;
      if (i < TP * NG) {
        const int kk = i & 7, g = (i >> 3) % NG, kg = (i >> 3) / NG;
        uint4 hi = make_uint4(0, 0, 0, 0), lo = hi;
        if (kg * 8 + kk < K2) ws_split8(x[u], hi, lo);
        const int off = (kg * NG + g) * 128 + kk * 16;
        *reinterpret_cast<uint4*>(sVhi + off) = hi;
        *reinterpret_cast<uint4*>(sVlo + off) = lo;
      }
    }
  }
}

}  // namespace

template <int TP, int DV, int ROUNDS>
__global__ void __launch_bounds__(NTHREADS, 1)
xattn_cell_tcws_kernel(naf_xattn_params p, int rh, int rw, int n_items, int vsplit, WsDivs dv, int use_stagers) {
  using Cfg = WsCfg<TP, DV, ROUNDS>;
  constexpr int K2 = WsWindowOf<TP>::K * WsWindowOf<TP>::K;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_q_full[2], bar_s_full[2], bar_p_full[2], bar_win_full[2], bar_win_free[2], bar_o_full, bar_o_free;
  __shared__ uint32_t tmem_base_s;

  uint8_t* win0 = smem;
  uint8_t* win1 = smem + Cfg::kWin;
  uint8_t* stage_out = smem + 2 * Cfg::kWin;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int npix = rh * rw;
  const int ntiles = (npix + 127) >> 7;
  const int tile_rows = (npix + ntiles - 1) / ntiles;   // balanced tiles, <= 128 rows
#if NAF_WS_BLOCKED
  // contiguous ranges: the first (n_items % grid) CTAs take one extra item
  const int items_lo = n_items / int(gridDim.x), items_rem = n_items - items_lo * int(gridDim.x);
  const int my_items = items_lo + (int(blockIdx.x) < items_rem ? 1 : 0);
  const int my_first = int(blockIdx.x) * items_lo + min(int(blockIdx.x), items_rem);
#else
  const int my_items = (n_items - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
#endif
  const int dv_full = p.C / p.heads;                    // = vsplit * DV
#if NAF_WS_BLOCKED
  auto item_of = [&](int it_seq) { return decode_item(my_first + it_seq, dv); };
#else
  auto item_of = [&](int it_seq) { return decode_item(blockIdx.x + it_seq * gridDim.x, dv); };
#endif
  auto vchan_of = [&](const ItemCoord& it) { return it.head * dv_full + it.vh * DV; };

  if (warp == MMA_WARP) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_q_full[s], NFRONT);
      mbar_init(&bar_s_full[s], 1);
      mbar_init(&bar_p_full[s], NFRONT);
      mbar_init(&bar_win_full[s], use_stagers ? NSTAGE : NBACK);
      mbar_init(&bar_win_free[s], 1);
    }
    mbar_init(&bar_o_full, 1);
    mbar_init(&bar_o_free, NBACK);
    fence_mbar_init();
  }
  // windows of this CTA's first item: staged by everybody
  if (my_items > 0 && tid < NFRONT + NBACK) {
    const ItemCoord it0 = item_of(0);
    stage_windows<TP, DV, ROUNDS>(win0, p, it0, vchan_of(it0), tid, NFRONT + NBACK);
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
    const float qscale = p.scale * 1.4426950408889634f;

    float qa[P], qb[P];
    int q_y = 0, q_x = 0;

    int f_g = 0;   // issue_q is called for g = 0, 1, 2, ... in order
    auto issue_q = [&]() {
      const int g = f_g++;
      const int it_seq = fdiv(g, dv.ntiles), tile = g - it_seq * ntiles;
      const ItemCoord it = item_of(it_seq);
      int pi = tile * tile_rows + row;
      const int lim = min(npix, (tile + 1) * tile_rows);
      if (row >= tile_rows || pi >= lim) pi = lim - 1;
      const int py = fdiv(pi, dv.rw);
      q_y = it.ci * rh + py;
      q_x = it.cj * rw + (pi - py * rw);
      const float* qp = p.q + int64_t(it.b) * p.q_stride_b + it.head * DQ + P * half +
                        int64_t(fdiv(q_y, dv.rep_y)) * p.q_stride_y + int64_t(fdiv(q_x, dv.rep_x)) * p.q_stride_x;
#if NAF_WS_EXP & 2
#pragma unroll
      for (int j = 0; j < P; ++j) { qa[j] = 0.01f * float(j + (pi & 7)); qb[j] = -0.02f * float(j); }
      (void)qp;
#else
      ldg_stream8(qp, *reinterpret_cast<float(*)[8]>(&qa[0]));
      ldg_stream8(qp + 8, *reinterpret_cast<float(*)[8]>(&qa[8]));
      ldg_stream8(qp + HALF, *reinterpret_cast<float(*)[8]>(&qb[0]));
      ldg_stream8(qp + HALF + 8, *reinterpret_cast<float(*)[8]>(&qb[8]));
#endif
    };
    // rotate + scale + split the prefetched q and write it to TMEM:
    // columns [0,32) hi, [32,64) lo; two channels per 32-bit column
    constexpr int QS = Cfg::kQStages;
    auto stage_q = [&](int g) {   // g: the tile whose q is in the registers
      if (rope) {
        const float* ct = half == 0 ? p.cos_y + int64_t(q_y) * P : p.cos_x + int64_t(q_x) * P;
        const float* st = half == 0 ? p.sin_y + int64_t(q_y) * P : p.sin_x + int64_t(q_x) * P;
        float c[P], sn[P];
        ldg8(ct, *reinterpret_cast<float(*)[8]>(&c[0]));
        ldg8(ct + 8, *reinterpret_cast<float(*)[8]>(&c[8]));
        ldg8(st, *reinterpret_cast<float(*)[8]>(&sn[0]));
        ldg8(st + 8, *reinterpret_cast<float(*)[8]>(&sn[8]));
        // rotation on channel pairs (packed fp32x2): qa' = (a c - b s) qscale, qb' = (b c + a s) qscale
        const uint64_t qs2 = ws_pack2(qscale, qscale);
#pragma unroll
        for (int j = 0; j < P; j += 2) {
          const uint64_t a2 = ws_pack2(qa[j], qa[j + 1]), b2 = ws_pack2(qb[j], qb[j + 1]);
          const uint64_t c2 = ws_pack2(c[j], c[j + 1]), s2 = ws_pack2(sn[j], sn[j + 1]);
          const uint64_t ra = ws_fma2(b2 ^ 0x8000000080000000ull, s2, ws_fma2(a2, c2, 0ull));
          const uint64_t rb = ws_fma2(a2, s2, ws_fma2(b2, c2, 0ull));
          ws_unpack2(ws_fma2(ra, qs2, 0ull), qa[j], qa[j + 1]);
          ws_unpack2(ws_fma2(rb, qs2, 0ull), qb[j], qb[j + 1]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < P; ++j) {
          qa[j] *= qscale;
          qb[j] *= qscale;
        }
      }
      // channel c lives in packed column c/2: a part -> columns [8h, 8h+8), b part -> [16+8h, ..)
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
    };

    if (total_tiles > 0) {
      issue_q();
      stage_q(0);
      if (total_tiles > 1) issue_q();
    }
    for (int g = 0; g < total_tiles; ++g) {
      const int s = g & 1;
      // (a) next tile's Q goes to TMEM early so that QK(g+1) runs under softmax(g).  With two Q
      // stages the buffer of tile g+1 was released by QK(g-1) (observed through s_full(g-1) in the
      // previous iteration); with one stage QK(g) itself must have finished first.
      if constexpr (QS == 2) {
        if (g + 1 < total_tiles) {
          stage_q(g + 1);
          if (g + 2 < total_tiles) issue_q();
        }
      }
      mbar_wait(&bar_s_full[s], (g >> 1) & 1);
      fence_after_sync();
      if constexpr (QS == 1) {
        if (g + 1 < total_tiles) {
          stage_q(g + 1);
          if (g + 2 < total_tiles) issue_q();
        }
      }
      // (b) softmax of tile g.  The body is compiled once per row half so that the tap mask
      // (columns >= K*K are padding) is a compile-time constant: no per-element predicates, and the
      // padded taps cost neither an exp nor a split.
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
        // The two halves of a row exchange their partial maxima through the 16-byte pad of the
        // row's staging slot (slot [tile parity][half]), and the same named barrier orders the S
        // reads of both halves before P overwrites the S columns.
        float* mpad = reinterpret_cast<float*>(stage_out + row * Cfg::kRowBytes + Cfg::kSlabs * Cfg::kRoundCols * 4) + s * 2;
        mpad[H] = m;
        fence_before_sync();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        fence_after_sync();
        m = fmaxf(m, mpad[H ^ 1]);
        // e = 2^(s - m) on value pairs (packed fp32x2 subtract and accumulate)
        const uint64_t one2 = ws_pack2(1.f, 1.f), negm2 = ws_pack2(-m, -m);
        uint64_t l2 = 0ull;
        float ev[SC];
#pragma unroll
        for (int j = 0; j < SC; j += 2) {
          if (j + 1 < NV) {
            float d0, d1;
            ws_unpack2(ws_fma2(ws_pack2(__uint_as_float(mine[j]), __uint_as_float(mine[j + 1])), one2, negm2), d0, d1);
            ev[j] = fast_exp2(d0);
            ev[j + 1] = fast_exp2(d1);
            l2 = ws_fma2(ws_pack2(ev[j], ev[j + 1]), one2, l2);
          } else if (j < NV) {
            ev[j] = fast_exp2(__uint_as_float(mine[j]) - m);
            ev[j + 1] = 0.f;
            l2 = ws_fma2(ws_pack2(ev[j], 0.f), one2, l2);
          } else {
            ev[j] = ev[j + 1] = 0.f;
          }
        }
        float l0, l1;
        ws_unpack2(l2, l0, l1);
        // row sum for the epilogue: spare TMEM column, slot [tile & 3][half] (the epilogue may lag
        // two tiles behind, so two slots would not be enough)
        tmem_st1(tmem + Cfg::kTmemL + lane_off + (g & 3) * 2 + H, __float_as_uint(l0 + l1));
        if constexpr (SC % 16 == 0) {
#pragma unroll
          for (int c0 = 0; c0 < SC; c0 += 16) {
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (c0 + 2 * j < NV) split2_f16(ev[c0 + 2 * j], ev[c0 + 2 * j + 1], hi[j], lo[j]);
              else hi[j] = lo[j] = 0u;
            }
            tmem_st8(ts + (base + c0) / 2, hi);
            tmem_st8(ts + TP / 2 + (base + c0) / 2, lo);
          }
        } else {
#pragma unroll
          for (int c0 = 0; c0 < SC; c0 += 8) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (c0 + 2 * j < NV) split2_f16(ev[c0 + 2 * j], ev[c0 + 2 * j + 1], hi[j], lo[j]);
              else hi[j] = lo[j] = 0u;
            }
            tmem_st4(ts + (base + c0) / 2, hi);
            tmem_st4(ts + TP / 2 + (base + c0) / 2, lo);
          }
        }
      };
      if (half == 0) softmax_half(std::integral_constant<int, 0>{});
      else softmax_half(std::integral_constant<int, 1>{});
      wait_st();
      fence_before_sync();
      mbar_arrive(&bar_p_full[s]);   // release: l and P (TMEM) are visible to the consumers
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
    uint8_t* my_stage = stage_out + row * Cfg::kRowBytes + half * HC * 4;
    int g = 0;
    for (int it_seq = 0; it_seq < my_items; ++it_seq) {
      const ItemCoord it = item_of(it_seq);
      // without stager warps (short items / small windows) this group converts the NEXT item's windows
      // itself; the buffer's last reader, item it_seq-1, has been fully drained by this group
      if (!use_stagers && it_seq + 1 < my_items) {
        const ItemCoord itn = item_of(it_seq + 1);
        stage_windows<TP, DV, ROUNDS>(((it_seq + 1) & 1) ? win1 : win0, p, itn, vchan_of(itn), tid - NFRONT, NBACK);
        fence_proxy_async_smem();
        mbar_arrive(&bar_win_full[(it_seq + 1) & 1]);
      }
      // element offset of this thread's half row slab inside the pixel; the output element is 4 bytes
      // (fp32) or 2 bytes (bf16, rounded on this final store)
      const int64_t obase = int64_t(it.b) * p.Ho * p.Wo * p.C + vchan_of(it) + half * HC;
      for (int tile = 0; tile < ntiles; ++tile, ++g) {
        const int pi = tile * tile_rows + row;
        const bool valid = row < tile_rows && pi < npix;
        const int py = fdiv(pi, dv.rw);
        const int y = it.ci * rh + py, x = it.cj * rw + (pi - py * rw);
        const int64_t oelem = obase + (int64_t(y) * p.Wo + x) * p.C;
        float* orow = static_cast<float*>(p.out) + oelem;   // fp32 view (EPI=1 path and fp32 stores)
        float inv_l = 0.f;
#if NAF_WS_EPI
        // lanes per contiguous run of the read-back (8 x 16 B = 128 B when the half row allows it)
        constexpr int PW = (HC % 32 == 0) ? 8 : 4;
        constexpr int RPI = 32 / PW, NCH = HC * 4 / (PW * 16);
        const int rsub = lane / PW, piece = lane % PW;
        const uint8_t* warp_stage = stage_out + (rowgrp * 32) * Cfg::kRowBytes + half * HC * 4;
        // global address of this lane's output row (0 = no row), handed to the other lanes by shuffle
        const uint64_t my_gp = valid ? reinterpret_cast<uint64_t>(orow) : 0ull;
#endif
#pragma unroll
        for (int rd = 0; rd < ROUNDS; ++rd) {
#if !NAF_WS_EPI
          // the previous bulk store from THIS slab must have finished reading the thread's bytes
          if constexpr (Cfg::kDoubleStage) bulk_wait_read<1>();
          else bulk_wait_read<0>();
#endif
          if (rd == 0) {
            mbar_wait(&bar_o_full, g & 1);
            fence_after_sync();
            uint32_t l0, l1;
            tmem_ld2(tmem + Cfg::kTmemL + lane_off + (g & 3) * 2, l0, l1);
            wait_ld();
            inv_l = 1.f / (__uint_as_float(l0) + __uint_as_float(l1));
          }
          const uint32_t to = tmem + Cfg::kTmemO + lane_off + rd * RC + half * HC;
          uint32_t r[2][16];
          tmem_ld16(to, r[0]);
#pragma unroll
          for (int c = 0; c < HC / 16; ++c) {
            wait_ld();
            if (c + 1 < HC / 16) tmem_ld16(to + (c + 1) * 16, r[(c + 1) & 1]);
            else if (rd == ROUNDS - 1) {
              // the whole O row is in registers / staged: the next PV may overwrite O
              fence_before_sync();
              mbar_arrive(&bar_o_free);
            }
            if (bf16_out) {
              // 16 values -> 16 bf16 = two 16-byte chunks of the (half as long) staged slab
#pragma unroll
              for (int j = 0; j < 16; j += 8) {
                uint4 pk;
                uint32_t* pw = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(r[c & 1][j + 2 * e]) * inv_l,
                                                                  __uint_as_float(r[c & 1][j + 2 * e + 1]) * inv_l);
                  pw[e] = *reinterpret_cast<const uint32_t*>(&h2);
                }
                *reinterpret_cast<uint4*>(my_stage + (Cfg::kDoubleStage ? rd * RC * 4 : 0) + (c * 16 + j) * 2) = pk;
              }
            } else {
              const uint64_t inv2 = ws_pack2(inv_l, inv_l);
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const uint64_t o01 = ws_fma2(ws_pack2(__uint_as_float(r[c & 1][j]), __uint_as_float(r[c & 1][j + 1])), inv2, 0ull);
                const uint64_t o23 = ws_fma2(ws_pack2(__uint_as_float(r[c & 1][j + 2]), __uint_as_float(r[c & 1][j + 3])), inv2, 0ull);
                *reinterpret_cast<uint4*>(my_stage + (Cfg::kDoubleStage ? rd * RC * 4 : 0) + (c * 16 + j) * 4) =
                    make_uint4(uint32_t(o01), uint32_t(o01 >> 32), uint32_t(o23), uint32_t(o23 >> 32));
              }
            }
          }
#if NAF_WS_EPI
          // the warp reads its 32 half rows back transposed: every store instruction writes RPI
          // runs of PW*16 contiguous bytes (one run per pixel row)
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 32 / RPI; ++i) {
            const int rr = i * RPI + rsub;
            const uint32_t gp_lo = __shfl_sync(0xffffffffu, uint32_t(my_gp), rr);
            const uint32_t gp_hi = __shfl_sync(0xffffffffu, uint32_t(my_gp >> 32), rr);
            float* gp = reinterpret_cast<float*>((uint64_t(gp_hi) << 32) | gp_lo);
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
              const float4 v = *reinterpret_cast<const float4*>(warp_stage + (Cfg::kDoubleStage ? rd * RC * 4 : 0) +
                                                                rr * Cfg::kRowBytes + (ch * PW + piece) * 16);
              if (gp) stg_stream(gp + rd * RC + (ch * PW + piece) * 4, v);
            }
          }
          __syncwarp();   // read-back done before the next round / tile overwrites the slab
#else
          // this thread's slab -> TMA engine (it only reads bytes this thread wrote)
          fence_proxy_async_smem();
          if (valid && !(NAF_WS_EXP & 1)) {
            const uint8_t* src = my_stage + (Cfg::kDoubleStage ? rd * RC * 4 : 0);
            if (bf16_out) bulk_store(static_cast<__nv_bfloat16*>(p.out) + oelem + rd * RC, src, HC * 2);
            else bulk_store(orow + rd * RC, src, (NAF_WS_EXP & 8) ? 16 : HC * 4);
          }
          bulk_commit();
#endif
        }
      }
    }
#if !NAF_WS_EPI
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all output writes performed
#endif
  } else if (warp > MMA_WARP) {
    // ======================================================================== WINDOW STAGERS
    // Two warps convert the K / V windows of item i+1 into the other window buffer while item i
    // runs, so the pipeline never waits for a window (staged by the epilogue warps this cost 0.7 ms
    // of 4.7 ms at C2).  A buffer is handed back by the MMA warp after the last PV that reads it.
    const int st = tid - (NFRONT + NBACK + 32);
    for (int it_seq = 1; use_stagers && it_seq < my_items; ++it_seq) {
      const int b = it_seq & 1;
      if (it_seq >= 2) mbar_wait(&bar_win_free[b], ((it_seq >> 1) - 1) & 1);
      const ItemCoord itn = item_of(it_seq);
      if (!(NAF_WS_EXP & 4)) stage_windows<TP, DV, ROUNDS>(b ? win1 : win0, p, itn, vchan_of(itn), st, NSTAGE);
      fence_proxy_async_smem();
      mbar_arrive(&bar_win_full[b]);
    }
  } else {
    // ======================================================================== MMA ISSUER
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc_f16(128, TP, false, false);
      constexpr uint32_t idesc_pv = make_idesc_f16(128, DV, false, true);
      const uint32_t tO = tmem + Cfg::kTmemO;
      constexpr int QS = Cfg::kQStages;
      int qk_seq = 0, qk_tile = 0, pv_seq = 0, pv_tile = 0;   // (item, tile) of the next QK / PV
      auto issue_qk = [&](int g) {
        const int s = g & 1;
        const int it_seq = qk_seq;
        if (qk_tile == 0 && it_seq > 0) mbar_wait(&bar_win_full[it_seq & 1], ((it_seq - 1) >> 1) & 1);
        const uint8_t* w = (it_seq & 1) ? win1 : win0;
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
          const uint32_t b0 = smem_u32(w + (pass == 2 ? Cfg::kSmemK : 0));
#pragma unroll
          for (int kk = 0; kk < DQ / 16; ++kk) {
            const uint64_t db = make_desc(b0 + kk * 2 * (TP * 16), TP * 16, 128);
            mma_f16_ts(tS, a0 + kk * 8, db, idesc_qk, (pass | kk) != 0);
          }
        }
        commit(&bar_s_full[s]);
      };
      auto issue_pv = [&](int g) {
        const int s = g & 1;
        mbar_wait(&bar_p_full[s], (g >> 1) & 1);
        mbar_wait(&bar_o_free, (g + 1) & 1);   // O drained by the epilogue of tile g-1
        fence_after_sync();
        const uint8_t* w = (pv_seq & 1) ? win1 : win0;
        const int pv_item = pv_seq;
        const bool last_of_item = pv_tile + 1 == ntiles;
        if (++pv_tile == ntiles) { pv_tile = 0; ++pv_seq; }
        const uint32_t tP = tmem + Cfg::kTmemS + s * TP;
        // O = Phi*Vhi + Plo*Vhi + Phi*Vlo
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t a0 = tP + (pass == 1 ? TP / 2 : 0);
          const uint32_t b0 = smem_u32(w + 2 * Cfg::kSmemK + (pass == 2 ? Cfg::kSmemV : 0));
#pragma unroll
          for (int kk = 0; kk < TP / 16; ++kk) {
            const uint64_t db = make_desc(b0 + kk * 2 * (DV / 8) * 128, (DV / 8) * 128, 128);
            mma_f16_ts(tO, a0 + kk * 8, db, idesc_pv, (pass | kk) != 0);
          }
        }
        commit(&bar_o_full);
        // every MMA that reads this item's windows has been issued: hand the buffer back to the stagers
        if (last_of_item) commit(&bar_win_free[pv_item & 1]);
      };
      if (total_tiles > 0) issue_qk(0);
      for (int g = 0; g < total_tiles; ++g) {
        if (g + 1 < total_tiles) issue_qk(g + 1);
        issue_pv(g);
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

int ws_taps_pad(int K) { return (K * K + 15) / 16 * 16; }

FastDiv make_fastdiv(uint32_t d) {
  // power of two: plain shift (magic == 0).  Otherwise, with s = ceil(log2 d):
  // magic = ceil(2^(31+s) / d) in (2^31, 2^32), n / d == umulhi(n, magic) >> (s - 1), exact for
  // every 0 <= n < 2^31.
  FastDiv f;
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

template <int TP, int DV, int ROUNDS>
int launch_ws(const naf_xattn_params& p, int vsplit, cudaStream_t st) {
  using Cfg = WsCfg<TP, DV, ROUNDS>;
  if constexpr (!Cfg::kFits) {
    return fail(NAF_ERR_UNSUPPORTED, "xattn(cell-tcws): tile %dx%d/%d does not fit", TP, DV, ROUNDS);
  } else {
    auto kern = xattn_cell_tcws_kernel<TP, DV, ROUNDS>;
    cudaError_t e = ensure_dyn_smem(kern, Cfg::kSmemTotal);
    if (e != cudaSuccess)
      return fail(NAF_ERR_CUDA, "xattn(cell-tcws): smem opt-in failed: %s", cudaGetErrorString(e));
    const int sms = device_sm_count();
    const int64_t items = int64_t(p.B) * p.h * p.w * p.heads * vsplit;
    const int grid = int(items < sms ? items : sms);
    const int rh = p.Ho / p.h, rw = p.Wo / p.w;
    WsDivs dv;
    dv.ntiles = make_fastdiv(uint32_t((rh * rw + 127) / 128));
    dv.vsplit = make_fastdiv(uint32_t(vsplit));
    dv.heads = make_fastdiv(uint32_t(p.heads));
    dv.w = make_fastdiv(uint32_t(p.w));
    dv.h = make_fastdiv(uint32_t(p.h));
    dv.rw = make_fastdiv(uint32_t(rw));
    dv.rep_y = make_fastdiv(uint32_t(p.rep_y));
    dv.rep_x = make_fastdiv(uint32_t(p.rep_x));
    // Dedicated stager warps pay off when an item is long enough to hide a 64-thread conversion of the
    // next windows (measured: C3 9.4 -> 8.7 ms, C5 14.6 -> 13.7 ms) and lose when items are short or the
    // windows small (C2 4.62 -> 4.76 ms, C1 0.064 -> 0.070 ms): then the epilogue warps stage, as before.
    const int ntiles = (rh * rw + 127) / 128;
    const int use_stagers = ntiles >= 6 && (TP >= 96 || ntiles >= 12);
    kern<<<grid, NTHREADS, Cfg::kSmemTotal, st>>>(p, rh, rw, int(items), vsplit, dv, use_stagers);
    return check_launch("xattn_cell_tcws");
  }
}

// The plan for a (taps, value head dim) pair: slab width DV = dv / vsplit and staging rounds,
// the first combination (fewest slabs, then fewest rounds) that fits TMEM and shared memory.
struct WsPlan {
  int dv_tile = 0, rounds = 0;
};

template <int TP, int DV>
constexpr int ws_rounds() {
  if (NAF_WS_DB && DV % 64 == 0 && WsCfg<TP, DV, 2>::kFits) return 2;   // double-buffered staging
  if (WsCfg<TP, DV, 1>::kFits) return 1;
  if (DV % 64 == 0 && WsCfg<TP, DV, 2>::kFits) return 2;
  if (DV % 128 == 0 && WsCfg<TP, DV, 4>::kFits) return 4;
  return 0;
}

template <int TP>
WsPlan ws_plan(int dv) {
  WsPlan pl;
  for (int vs = 1; vs <= 4; vs *= 2) {
    if (dv % vs) break;
    const int t = dv / vs;
    int r = 0;
    switch (t) {
      case 32: r = ws_rounds<TP, 32>(); break;
      case 64: r = ws_rounds<TP, 64>(); break;
      case 96: r = ws_rounds<TP, 96>(); break;
      case 128: r = ws_rounds<TP, 128>(); break;
      case 192: r = ws_rounds<TP, 192>(); break;
      case 256: r = ws_rounds<TP, 256>(); break;
      default: r = 0;
    }
    if (r) {
      pl.dv_tile = t;
      pl.rounds = r;
      return pl;
    }
  }
  return pl;
}

WsPlan ws_plan_for(int K, int dv) {
  switch (ws_taps_pad(K)) {
    case 16: return ws_plan<16>(dv);
    case 32: return ws_plan<32>(dv);
    case 64: return ws_plan<64>(dv);
    case 96: return ws_plan<96>(dv);
    case 128: return ws_plan<128>(dv);
    default: return WsPlan{};
  }
}

template <int TP, int DV>
int launch_ws_rounds(const naf_xattn_params& p, int rounds, int vsplit, cudaStream_t st) {
  switch (rounds) {
    case 1: return launch_ws<TP, DV, 1>(p, vsplit, st);
    case 2: if constexpr (DV % 64 == 0) return launch_ws<TP, DV, 2>(p, vsplit, st); break;
    case 4: if constexpr (DV % 128 == 0) return launch_ws<TP, DV, 4>(p, vsplit, st); break;
    default: break;
  }
  return fail(NAF_ERR_UNSUPPORTED, "xattn(cell-tcws): bad plan");
}

template <int TP>
int launch_ws_tp(const naf_xattn_params& p, cudaStream_t st) {
  const int dv = p.C / p.heads;
  const WsPlan pl = ws_plan<TP>(dv);
  if (!pl.dv_tile) return fail(NAF_ERR_UNSUPPORTED, "xattn(cell-tcws): no plan for dv=%d", dv);
  const int vsplit = dv / pl.dv_tile;
  switch (pl.dv_tile) {
    case 32: return launch_ws_rounds<TP, 32>(p, pl.rounds, vsplit, st);
    case 64: return launch_ws_rounds<TP, 64>(p, pl.rounds, vsplit, st);
    case 96: return launch_ws_rounds<TP, 96>(p, pl.rounds, vsplit, st);
    case 128: return launch_ws_rounds<TP, 128>(p, pl.rounds, vsplit, st);
    case 192: return launch_ws_rounds<TP, 192>(p, pl.rounds, vsplit, st);
    case 256: return launch_ws_rounds<TP, 256>(p, pl.rounds, vsplit, st);
    default: return fail(NAF_ERR_UNSUPPORTED, "xattn(cell-tcws): bad plan");
  }
}

}  // namespace

bool xattn_cell_tcws_supported(const naf_xattn_params& p, const char** why) {
  if (p.Kw != 0 && p.Kw != p.K) { *why = "rectangular window"; return false; }
  const int dq = p.D / p.heads, dv = p.C / p.heads;
  if (p.row_tap || p.col_tap) { *why = "tap tables given (non-integer ratio path)"; return false; }
  if (p.Ho % p.h || p.Wo % p.w) { *why = "target size is not a multiple of the feature size"; return false; }
  if (p.scores) { *why = "score output requested"; return false; }
  if (p.q_dtype != NAF_DTYPE_F32 || p.k_dtype != NAF_DTYPE_F32 || p.v_dtype != NAF_DTYPE_F32) { *why = "fp32 inputs only"; return false; }
  if (NAF_WS_EPI && p.out_dtype != NAF_DTYPE_F32) { *why = "fp32 output only in this build"; return false; }
  if (dq != DQ) { *why = "head dim must be 64"; return false; }
  if (p.K < 3) { *why = "kernel_size must be 3, 5, 7, 9 or 11"; return false; }
  if (dv % 32) { *why = "value head dim must be a multiple of 32"; return false; }
  const WsPlan pl = ws_plan_for(p.K, dv);
  if (!pl.dv_tile) { *why = "window / value tile does not fit shared memory + TMEM of the pipelined kernel"; return false; }
  if ((p.Ho / p.h) * (p.Wo / p.w) < 64) { *why = "fewer than 64 pixels per cell"; return false; }
  if (!aligned32(p.q) || !aligned32(p.k) || !aligned32(p.v) || !aligned32(p.out) ||
      (p.q_stride_b % 8) || (p.q_stride_y % 8) || (p.q_stride_x % 8)) {
    *why = "pointers/strides not 32-byte aligned";
    return false;
  }
  if (p.cos_y && !(aligned32(p.cos_y) && aligned32(p.sin_y) && aligned32(p.cos_x) && aligned32(p.sin_x))) {
    *why = "rope tables not 32-byte aligned";
    return false;
  }
  if (int64_t(p.B) * p.h * p.w * p.heads * (dv / pl.dv_tile) >= (int64_t(1) << 31)) { *why = "too many items"; return false; }
  return true;
}

int launch_xattn_cell_tcws(const naf_xattn_params& p, cudaStream_t st) {
  switch (ws_taps_pad(p.K)) {
    case 16: return launch_ws_tp<16>(p, st);
    case 32: return launch_ws_tp<32>(p, st);
    case 64: return launch_ws_tp<64>(p, st);
    case 96: return launch_ws_tp<96>(p, st);
    case 128: return launch_ws_tp<128>(p, st);
    default: return fail(NAF_ERR_UNSUPPORTED, "xattn(cell-tcws): kernel_size %d", p.K);
  }
}

}  // namespace naf
