// Union-window kernel: the tensor-core path (tcgen05 / TMEM) for the requests the cell kernels cannot
// take -- tap tables (non-integer ratios with duplicated taps: training's 32 <- 13,
// utils/training.py:28-50), ratio 1 (denoising: C = 3, one head, K up to 15, denoising.py:209-213,436-451),
// integer ratios with cells of a few pixels, any head dim that is a multiple of 16.
//
// The windows of neighbouring target pixels are no longer one shared window, but they overlap almost
// entirely.  One CTA owns (batch, head, 8 x 16 tile of target pixels = 128 TMEM lanes).  The low-resolution
// cells any pixel of the tile touches form a rectangle U = [i0,i1] x [j0,j1] (its "union window"); the
// neighbourhood attention of the tile is DENSE attention of the 128 pixels against U under a multiplicity
// mask:  with  m[p][c] = #{(t,u) : (row_tap[y_p][t], col_tap[x_p][u]) = c} = mr[p][c.row] * mc[p][c.col],
//     out[p] = sum_c m[p][c] e^{s[p][c]} v[c] / sum_c m[p][c] e^{s[p][c]},      s = scale * Q K_U^T
// which is exactly softmax over the K*K taps followed by the weighted sum (a cell hit by two taps counts
// twice; cells outside the pixel's window have m = 0).  U is walked in chunks of 128 cells (whole rows of
// U, padded to 16 or 32 columns) with an online softmax (running max / sum, O rescaled in TMEM when the max
// moves), so K = 15 at ratio 1 (22 x 30 cells) fits:
//     per chunk:   S[128 x 128] = Q[128 x dq] K_chunk^T        tcgen05.mma, Q is the A operand in TMEM
//                  P = m * 2^(S - max)                          two threads per pixel row
//                  O[128 x dv] (+)= P[128 x 128] V_chunk        P is the A operand in TMEM
// Precision scheme of the cell kernels: every operand split fp32 -> fp16 hi + fp16 lo, three MMA passes
// (hi*hi + lo*hi + hi*lo), fp32 accumulation.  Head dims are run-time values (one instantiation).
// These launches are small (a 32 x 32 or 256 x 256 target); the kernel is deliberately not pipelined.
//
// Reference semantics: src/layers/attentions.py:16-29,53-75; tap order irrelevant here (no score output:
// return_weights requests stay on the generic kernel).
#include <cuda_bf16.h>

#include <climits>

#include "naf_common.cuh"
#include "naf_umma.cuh"

namespace naf {

using namespace umma;

namespace {

constexpr int NT = 256;       // threads per CTA: two per pixel row / TMEM lane
constexpr int TH = 8;         // target tile: 8 rows x 16 columns = 128 pixels
constexpr int TW = 16;
constexpr int NC = 128;       // union cells per chunk = S columns
constexpr int RMAX = 32;      // largest union side a tile may have
constexpr int RROWS = RMAX + 8;  // rows of the row-multiplicity table (chunks may overrun RH by < 8)

__device__ __forceinline__ void split4(const float (&x)[4], uint32_t (&hi)[2], uint32_t (&lo)[2]) {
  split2_f16(x[0], x[1], hi[0], lo[0]);
  split2_f16(x[2], x[3], hi[1], lo[1]);
}

// 8 consecutive channels [ch0, ch0+8) of a pixel vector with `n` valid channels; zeros beyond
__device__ __forceinline__ void load8_guard(const float* base, int ch0, int n, bool vec, float (&x)[8]) {
  if (vec && ch0 + 8 <= n) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(base + ch0));
    const float4 b = __ldg(reinterpret_cast<const float4*>(base + ch0 + 4));
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
    x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = (ch0 + e < n) ? __ldg(base + ch0 + e) : 0.f;
  }
}

struct UnionGeom {
  int rh, rw;          // Ho / h, Wo / w (the integer rule's ratios; unused with tables)
  int tiles_y, tiles_x;
  int dq, dv, dvp;     // head dims; dvp = dv rounded up to 16
  int tmem_cols;
  int vec_v;           // V rows may be read with 128-bit loads
  int vec_o;           // fp32 output rows may be written with 128-bit stores
};

// One chunk's softmax for this thread's 64 S columns.  RWP = padded union width (16 or 32): column j of
// the chunk is union row j / RWP (of this chunk), union column j % RWP.
template <int RWP>
__device__ __forceinline__ void softmax_chunk(uint32_t tS, uint32_t lane_off, int hf, int row, int chunk,
                                              const uint8_t (*sMr)[128], const uint32_t (&mcw)[8],
                                              float (*red_m)[128], float (*red_l)[128], float& m_run,
                                              float& l_run, float& alpha) {
  constexpr int SC = NC / 2;            // columns per thread
  constexpr int RPT = SC / RWP;         // union rows per thread: 4 or 2
  constexpr int RPC = NC / RWP;         // union rows per chunk
  uint32_t s[SC];
#pragma unroll
  for (int c0 = 0; c0 < SC; c0 += 16)
    tmem_ld16(tS + lane_off + hf * SC + c0, *reinterpret_cast<uint32_t(*)[16]>(&s[c0]));
  int mrv[RPT];
#pragma unroll
  for (int q = 0; q < RPT; ++q) mrv[q] = sMr[chunk * RPC + hf * RPT + q][row];
  wait_ld();
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < SC; ++j) {
    const int jj = j % RWP;
    const int mult = mrv[j / RWP] * int((mcw[jj >> 2] >> (8 * (jj & 3))) & 0xffu);
    if (mult) m = fmaxf(m, __uint_as_float(s[j]));
  }
  red_m[hf][row] = m;
  fence_before_sync();   // the S columns were read: P may overwrite them after the barrier
  __syncthreads();
  fence_after_sync();
  const float m_new = fmaxf(m_run, fmaxf(red_m[0][row], red_m[1][row]));
  const float m_use = m_new == -INFINITY ? 0.f : m_new;
  alpha = fast_exp2(m_run - m_use);     // m_run = -inf -> 0
  m_run = m_new;
  float l = 0.f;
#pragma unroll
  for (int j = 0; j < SC; ++j) {
    const int jj = j % RWP;
    const int mult = mrv[j / RWP] * int((mcw[jj >> 2] >> (8 * (jj & 3))) & 0xffu);
    const float e = mult ? float(mult) * fast_exp2(__uint_as_float(s[j]) - m_use) : 0.f;
    l += e;
    s[j] = __float_as_uint(e);
  }
  red_l[hf][row] = l;
  // P: fp16 hi halves in columns [0, NC/2), lo halves in [NC/2, NC); two cells per 32-bit column
#pragma unroll
  for (int c0 = 0; c0 < SC; c0 += 16) {
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      split2_f16(__uint_as_float(s[c0 + 2 * j]), __uint_as_float(s[c0 + 2 * j + 1]), hi[j], lo[j]);
    tmem_st8(tS + lane_off + (hf * SC + c0) / 2, hi);
    tmem_st8(tS + lane_off + NC / 2 + (hf * SC + c0) / 2, lo);
  }
  (void)l_run;
}

}  // namespace

__global__ void __launch_bounds__(NT, 1)
xattn_union_tc_kernel(naf_xattn_params p, UnionGeom g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t mbar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ int ubox[4];                 // i0, i1, j0, j1
  __shared__ float red_m[2][128];
  __shared__ float red_l[2][128];
  __shared__ __align__(16) uint8_t sMr[RROWS][128];   // row multiplicities, [union row][pixel]
  __shared__ __align__(16) uint8_t sMc[RMAX][128];    // column multiplicities

  const int dq = g.dq, dv = g.dv, dvp = g.dvp;
  const int KC = dq >> 3;                 // 16-byte chunks along the head dim
  const int NG = dvp >> 3;                // channel groups of 8 along the value head
  uint8_t* sKhi = smem;
  uint8_t* sKlo = sKhi + KC * NC * 16;
  uint8_t* sVhi = sKlo + KC * NC * 16;
  uint8_t* sVlo = sVhi + NC * dvp * 2;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifndef NAF_UNION_ELECT
#define NAF_UNION_ELECT 1   // 1: the MMAs are issued under elect.sync (descriptors in uniform registers), 0: by lane 0
#endif
  const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform (see naf_umma.cuh:elect_one_sync)
  const int rowgrp = warp & 3;      // TMEM lane quarter this warp may touch
  const int hf = warp >> 2;         // which half of the columns this thread owns
  const int row = rowgrp * 32 + lane;
  const int K = p.K;

  int bid = blockIdx.x;
  const int head = bid % p.heads;
  bid /= p.heads;
  const int tx0 = (bid % g.tiles_x) * TW;
  bid /= g.tiles_x;
  const int ty0 = (bid % g.tiles_y) * TH;
  const int b = bid / g.tiles_y;

  const int py = ty0 + row / TW, px = tx0 + row % TW;
  const bool valid = py < p.Ho && px < p.Wo;
  const int my_y = min(py, p.Ho - 1), my_x = min(px, p.Wo - 1);

  if (warp == 0) tmem_alloc(&tmem_base_s, g.tmem_cols);
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    fence_mbar_init();
    ubox[0] = INT_MAX; ubox[1] = -1; ubox[2] = INT_MAX; ubox[3] = -1;
  }
  for (int i = tid; i < (RROWS + RMAX) * 128 / 4; i += NT) {
    if (i < RROWS * 32) reinterpret_cast<uint32_t*>(&sMr[0][0])[i] = 0;
    else reinterpret_cast<uint32_t*>(&sMc[0][0])[i - RROWS * 32] = 0;
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();

  // ---- union box of the tile
  for (int i = tid; i < TH * K; i += NT) {
    const int y = ty0 + i / K;
    if (y < p.Ho) {
      const int r = tap_index(p.row_tap, y, i % K, K, g.rh, p.h);
      atomicMin(&ubox[0], r);
      atomicMax(&ubox[1], r);
    }
  }
  for (int i = tid; i < TW * K; i += NT) {
    const int x = tx0 + i / K;
    if (x < p.Wo) {
      const int c = tap_index(p.col_tap, x, i % K, K, g.rw, p.w);
      atomicMin(&ubox[2], c);
      atomicMax(&ubox[3], c);
    }
  }
  __syncthreads();
  const int i0 = ubox[0], j0 = ubox[2];
  const int RH = ubox[1] - i0 + 1, RW = ubox[3] - j0 + 1;
  // Tables built from NATTEN windows and nearest-exact resizing stay far inside this (<= K + 17 cells for
  // K <= 15 at any ratio >= 1); anything else is not a neighbourhood table: fail loudly.
  if (RH > RMAX || RW > RMAX || i0 < 0 || j0 < 0 || ubox[1] >= p.h || ubox[3] >= p.w) __trap();
  const bool wide = RW > 16;
  const int rwp_shift = wide ? 5 : 4;
  const int RPC = NC >> rwp_shift;                 // union rows per chunk: 4 or 8
  const int nchunks = (RH + RPC - 1) / RPC;

  // ---- multiplicity tables: hf 0 counts the row taps of its pixel, hf 1 the column taps
  if (valid) {
    if (hf == 0) {
      for (int t = 0; t < K; ++t) sMr[tap_index(p.row_tap, py, t, K, g.rh, p.h) - i0][row] += 1;
    } else {
      for (int u = 0; u < K; ++u) sMc[tap_index(p.col_tap, px, u, K, g.rw, p.w) - j0][row] += 1;
    }
  }

  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_off = uint32_t(rowgrp * 32) << 16;
  const uint32_t tQ = tmem;                   // Q: fp16 hi in [0, dq/2), lo in [dq/2, dq)
  const uint32_t tS = tmem + dq;              // S (fp32, NC columns) / P (fp16 hi | lo)
  const uint32_t tO = tmem + dq + NC;         // O (fp32, dvp columns)

  // ---- Q -> TMEM (A operand of S = Q K^T): rotate, fold scale * log2(e), split
  {
    const bool rope = p.cos_y != nullptr;
    const int half = dq >> 1, P = dq >> 2;    // P rotation pairs per axis
    const float qscale = p.scale * 1.4426950408889634f;
    const float* qp = p.q + int64_t(b) * p.q_stride_b + int64_t(my_y / p.rep_y) * p.q_stride_y +
                      int64_t(my_x / p.rep_x) * p.q_stride_x + head * dq + P * hf;
    // this thread owns pairs [P*hf, P*hf + P): channels a = P*hf + i and b = a + dq/2; hf 0 rotates by the
    // row angles, hf 1 by the column angles (src/layers/rope.py:139-143)
    const float* ct = hf == 0 ? p.cos_y + int64_t(my_y) * P : p.cos_x + int64_t(my_x) * P;
    const float* st = hf == 0 ? p.sin_y + int64_t(my_y) * P : p.sin_x + int64_t(my_x) * P;
    for (int i = 0; i < P; i += 4) {
      const float4 a4 = __ldg(reinterpret_cast<const float4*>(qp + i));
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(qp + half + i));
      float a[4] = {a4.x, a4.y, a4.z, a4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
      if (rope) {
        const float4 c4 = __ldg(reinterpret_cast<const float4*>(ct + i));
        const float4 s4 = __ldg(reinterpret_cast<const float4*>(st + i));
        const float c[4] = {c4.x, c4.y, c4.z, c4.w}, s[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float ra = a[e] * c[e] - bb[e] * s[e];
          const float rb = bb[e] * c[e] + a[e] * s[e];
          a[e] = ra;
          bb[e] = rb;
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        a[e] *= qscale;
        bb[e] *= qscale;
      }
      uint32_t hi[2], lo[2];
      split4(a, hi, lo);
      const int ca = (P * hf + i) >> 1;             // 32-bit column of channel a (two fp16 per column)
      tmem_st2(tQ + lane_off + ca, hi[0], hi[1]);
      tmem_st2(tQ + lane_off + half + ca, lo[0], lo[1]);
      split4(bb, hi, lo);
      const int cb = (half + P * hf + i) >> 1;
      tmem_st2(tQ + lane_off + cb, hi[0], hi[1]);
      tmem_st2(tQ + lane_off + half + cb, lo[0], lo[1]);
    }
    wait_st();
  }
  __syncthreads();   // multiplicity tables complete

  // this thread's column multiplicities, packed four per register
  uint32_t mcw[8];
#pragma unroll
  for (int w4 = 0; w4 < 8; ++w4) {
    uint32_t v = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) v |= uint32_t(sMc[w4 * 4 + e][row]) << (8 * e);
    mcw[w4] = v;
  }

  const uint32_t idesc_qk = make_idesc_f16(128, NC, false, false);
  const uint32_t idesc_pv = make_idesc_f16(128, dvp, false, true);
  float m_run = -INFINITY, l_run = 0.f;
  uint32_t phase = 0;
  const float* kbase = p.k + int64_t(b) * p.h * p.w * p.D + head * dq;
  const float* vbase = p.v + int64_t(b) * p.h * p.w * p.C + head * dv;

  for (int chunk = 0; chunk < nchunks; ++chunk) {
    const int ci0 = chunk * RPC;      // first union row of the chunk
    // ---- stage the K chunk: canonical K-major [channel chunk c][cell n][16 B]; zero rows for padding
    for (int i = tid; i < NC * KC; i += NT) {
      const int n = i & (NC - 1), c = i >> 7;
      const int ii = ci0 + (n >> rwp_shift), jj = n & ((1 << rwp_shift) - 1);
      uint4 hi = make_uint4(0, 0, 0, 0), lo = hi;
      if (ii < RH && jj < RW) {
        float x[8];
        load8_guard(kbase + (int64_t(i0 + ii) * p.w + j0 + jj) * p.D, c * 8, dq, true, x);
        split2_f16(x[0], x[1], hi.x, lo.x);
        split2_f16(x[2], x[3], hi.y, lo.y);
        split2_f16(x[4], x[5], hi.z, lo.z);
        split2_f16(x[6], x[7], hi.w, lo.w);
      }
      *reinterpret_cast<uint4*>(sKhi + (c * NC + n) * 16) = hi;
      *reinterpret_cast<uint4*>(sKlo + (c * NC + n) * 16) = lo;
    }
    // ---- stage the V chunk: canonical MN-major [cell group][channel group][cell % 8][16 B]
    for (int i = tid; i < NC * NG; i += NT) {
      const int kk = i & 7;
      const int cg = (i >> 3) % NG;
      const int kg = (i >> 3) / NG;
      const int n = kg * 8 + kk;
      const int ii = ci0 + (n >> rwp_shift), jj = n & ((1 << rwp_shift) - 1);
      uint4 hi = make_uint4(0, 0, 0, 0), lo = hi;
      if (ii < RH && jj < RW) {
        float x[8];
        load8_guard(vbase + (int64_t(i0 + ii) * p.w + j0 + jj) * p.C, cg * 8, dv, g.vec_v != 0, x);
        split2_f16(x[0], x[1], hi.x, lo.x);
        split2_f16(x[2], x[3], hi.y, lo.y);
        split2_f16(x[4], x[5], hi.z, lo.z);
        split2_f16(x[6], x[7], hi.w, lo.w);
      }
      const int off = (kg * NG + cg) * 128 + kk * 16;
      *reinterpret_cast<uint4*>(sVhi + off) = hi;
      *reinterpret_cast<uint4*>(sVlo + off) = lo;
    }
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    if (warp_u == 0 && (NAF_UNION_ELECT ? elect_one_sync() : lane == 0)) {
      fence_after_sync();
      // S = Qhi*Khi^T + Qlo*Khi^T + Qhi*Klo^T
      for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a0 = tQ + (pass == 1 ? (dq >> 1) : 0);
        const uint32_t b0 = smem_u32(pass == 2 ? sKlo : sKhi);
        for (int kk = 0; kk < (dq >> 4); ++kk) {
          const uint64_t db = make_desc(b0 + kk * 2 * (NC * 16), NC * 16, 128);
          mma_f16_ts(tS, a0 + kk * 8, db, idesc_qk, (pass | kk) != 0);
        }
      }
      commit(&mbar[0]);
    }
    mbar_wait(&mbar[0], phase);
    fence_after_sync();

    float alpha;
    if (wide) softmax_chunk<32>(tS, lane_off, hf, row, chunk, sMr, mcw, red_m, red_l, m_run, l_run, alpha);
    else softmax_chunk<16>(tS, lane_off, hf, row, chunk, sMr, mcw, red_m, red_l, m_run, l_run, alpha);
    // ---- the running max moved: rescale this thread's 16-column chunks of O (chunks c % 2 == hf)
    if (chunk > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
      for (int c16 = hf; c16 < (dvp >> 4); c16 += 2) {
        uint32_t o[16];
        tmem_ld16(tO + lane_off + c16 * 16, o);
        wait_ld();
#pragma unroll
        for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
        tmem_st16(tO + lane_off + c16 * 16, o);
      }
    }
    wait_st();
    fence_before_sync();
    __syncthreads();
    l_run = l_run * alpha + red_l[0][row] + red_l[1][row];
    if (warp_u == 0 && (NAF_UNION_ELECT ? elect_one_sync() : lane == 0)) {
      fence_after_sync();
      // O (+)= Phi*Vhi + Plo*Vhi + Phi*Vlo
      for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a0 = tS + (pass == 1 ? NC / 2 : 0);
        const uint32_t b0 = smem_u32(pass == 2 ? sVlo : sVhi);
#pragma unroll
        for (int kk = 0; kk < NC / 16; ++kk) {
          const uint64_t db = make_desc(b0 + kk * 2 * NG * 128, NG * 128, 128);
          mma_f16_ts(tO, a0 + kk * 8, db, idesc_pv, (chunk | pass | kk) != 0);
        }
      }
      commit(&mbar[1]);
    }
    mbar_wait(&mbar[1], phase);   // the staging buffers, S and (for the rescale) O are free again
    fence_after_sync();
    phase ^= 1;
  }

  // ---- epilogue: this thread's 16-column chunks of its O row -> normalise -> store
  {
    const float inv_l = 1.f / l_run;
    const int64_t opix = (int64_t(b) * p.Ho + my_y) * p.Wo + my_x;
    const bool vec_o = g.vec_o != 0;
    for (int c16 = hf; c16 < (dvp >> 4); c16 += 2) {
      uint32_t o[16];
      tmem_ld16(tO + lane_off + c16 * 16, o);
      wait_ld();
      if (!valid) continue;
      const int ch0 = c16 * 16;
      if (p.out_dtype == NAF_DTYPE_BF16) {
        __nv_bfloat16* op = static_cast<__nv_bfloat16*>(p.out) + opix * p.C + head * dv + ch0;
#pragma unroll
        for (int e = 0; e < 16; ++e)
          if (ch0 + e < dv) op[e] = __float2bfloat16_rn(__uint_as_float(o[e]) * inv_l);
      } else {
        float* op = static_cast<float*>(p.out) + opix * p.C + head * dv + ch0;
        if (vec_o) {
#pragma unroll
          for (int e = 0; e < 16; e += 4)
            if (ch0 + e < dv)
              stg_stream(op + e, make_float4(__uint_as_float(o[e]) * inv_l, __uint_as_float(o[e + 1]) * inv_l,
                                             __uint_as_float(o[e + 2]) * inv_l, __uint_as_float(o[e + 3]) * inv_l));
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (ch0 + e < dv) op[e] = __uint_as_float(o[e]) * inv_l;
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, g.tmem_cols);
}

// ------------------------------------------------------------------------------------ host side
namespace {

bool fill_geom(const naf_xattn_params& p, UnionGeom& g, size_t& smem_bytes) {
  g.rh = p.Ho / p.h;
  g.rw = p.Wo / p.w;
  g.tiles_y = (p.Ho + TH - 1) / TH;
  g.tiles_x = (p.Wo + TW - 1) / TW;
  g.dq = p.D / p.heads;
  g.dv = p.C / p.heads;
  g.dvp = (g.dv + 15) / 16 * 16;
  const int cols = g.dq + NC + g.dvp;
  g.tmem_cols = cols <= 256 ? 256 : 512;
  g.vec_v = ((g.dv & 3) == 0 && aligned16(p.v)) ? 1 : 0;
  g.vec_o = ((g.dv & 3) == 0 && aligned16(p.out)) ? 1 : 0;
  smem_bytes = size_t(g.dq) * NC * 4 + size_t(g.dvp) * NC * 4;
  return cols <= 512;
}

}  // namespace

bool xattn_union_tc_supported(const naf_xattn_params& p, const char** why) {
  if (p.Kw != 0 && p.Kw != p.K) { *why = "rectangular window"; return false; }
  if (p.q_dtype != NAF_DTYPE_F32 || p.k_dtype != NAF_DTYPE_F32 || p.v_dtype != NAF_DTYPE_F32) {
    *why = "fp32 inputs only";
    return false;
  }
  if (p.scores) { *why = "score output requested (per-tap scores are not formed)"; return false; }
  const int dq = p.D / p.heads, dv = p.C / p.heads;
  if (dq % 16 || dq < 16 || dq > 256) { *why = "head dim must be a multiple of 16 in [16, 256]"; return false; }
  if (dv < 1 || dv > 256) { *why = "value head dim must be in [1, 256]"; return false; }
  if (p.K > 15) { *why = "kernel_size > 15"; return false; }
  UnionGeom g;
  size_t smem = 0;
  if (!fill_geom(p, g, smem)) { *why = "head dims need more than 512 TMEM columns (dq + 128 + dv)"; return false; }
  if (smem > 200 * 1024) { *why = "head dims need more than 200 KB of shared memory"; return false; }
  if (!aligned16(p.q) || !aligned16(p.k) || (p.q_stride_b % 4) || (p.q_stride_y % 4) || (p.q_stride_x % 4)) {
    *why = "q / k pointers or strides not 16-byte aligned";
    return false;
  }
  if (p.cos_y && !(aligned16(p.cos_y) && aligned16(p.sin_y) && aligned16(p.cos_x) && aligned16(p.sin_x))) {
    *why = "rope tables not 16-byte aligned";
    return false;
  }
  if (int64_t(p.B) * p.heads * g.tiles_y * g.tiles_x >= (int64_t(1) << 31)) { *why = "grid too large"; return false; }
  return true;
}

int launch_xattn_union_tc(const naf_xattn_params& p, cudaStream_t st) {
  UnionGeom g;
  size_t smem = 0;
  fill_geom(p, g, smem);
  cudaError_t e = ensure_dyn_smem(xattn_union_tc_kernel, int(smem));
  if (e != cudaSuccess) return fail(NAF_ERR_CUDA, "xattn(union-tc): smem opt-in failed: %s", cudaGetErrorString(e));
  const unsigned grid = unsigned(p.B) * p.heads * g.tiles_y * g.tiles_x;
  xattn_union_tc_kernel<<<grid, NT, smem, st>>>(p, g);
  return check_launch("xattn_union_tc");
}

}  // namespace naf
