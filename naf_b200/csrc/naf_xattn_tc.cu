// Cell kernel, tensor-core path (tcgen05 / TMEM): cross-scale neighbourhood attention for
// integer ratios with 64-wide heads.
//
// One CTA owns one (batch, low-res cell, head).  All rh*rw target pixels of the cell attend over
// the same clamped K x K low-res window, so per 128-pixel tile
//     S[128 x TP]  = Q[128 x 64] * Kwin[TP x 64]^T          (TP = K*K padded to 16)
//     O[128 x DV]  = softmax(S)[128 x TP] * Vwin[TP x DV]
// are two dense GEMMs issued with tcgen05.mma (M=128, fp32 accumulators in TMEM):
//   * 256 threads, TWO threads per pixel row / TMEM lane (warps w and w+4 share lane quarter
//     w%4 and split the columns).  Each thread loads half of its pixel's q vector straight from
//     HBM with 256-bit loads (RoPE on the fly: one half owns the row angles, the other the
//     column angles), writes it to shared memory in the canonical K-major UMMA layout; the
//     softmax runs on the half S row read back with tcgen05.ld (max / sum exchanged through a
//     tiny smem array); P goes back to TMEM over the S columns and feeds the second GEMM as its
//     A operand; the O half rows come back with tcgen05.ld, are normalised and leave as 256-bit
//     streaming stores (each thread writes whole 32 B sectors of its own 128 B lines).
//   * precision: every operand is split fp32 -> fp16 hi + fp16 lo and each GEMM is the three
//     passes hi*hi + lo*hi + hi*lo (~22 mantissa bits, fp32 accumulation), so the result stays
//     within ~1e-5 of the fp32 SIMT kernel while running at kind::f16 rate.
//   * 2 CTAs per SM (96 KB smem, 256 TMEM columns each, 16 warps) overlap one CTA's loads /
//     softmax / epilogue with the other's MMAs; the next tile's q is prefetched into registers.
// HBM traffic = algorithmic minimum (q once, out once; windows from L2).
//
// Reference semantics: src/layers/attentions.py:16-29,53-75; RoPE src/layers/rope.py:137-153.
#include "naf_common.cuh"
#include "naf_umma.cuh"

#ifndef NAF_TC_EXP
#define NAF_TC_EXP 0   // profiling experiments only (scripts/tc_experiments.sh); 0 in every real build
#endif

namespace naf {

using namespace umma;

namespace {

constexpr int DQ = 64;        // query/key head dim this kernel is specialised for
constexpr int KC = DQ / 8;    // 16-byte chunks along the head dim
constexpr int NT = 256;       // threads per CTA

// window side <-> padded tap count is one-to-one for the supported kernels (3,5,7,9,11)
template <int TP>
struct WindowOf { static constexpr int K = TP == 16 ? 3 : TP == 32 ? 5 : TP == 64 ? 7 : TP == 96 ? 9 : 11; };

template <int TP, int DV>
struct TcCfg {
  static constexpr int kSmemK = KC * TP * 16;        // one of hi / lo
  static constexpr int kSmemV = TP * DV * 2;         // one of hi / lo
  static constexpr int kSmemQ = KC * 128 * 16;       // one of hi / lo
  static constexpr int kSmemTotal = 2 * kSmemK + 2 * kSmemV + 2 * kSmemQ;
  static constexpr int kTmemCols = (TP + DV) <= 64 ? 64 : (TP + DV) <= 128 ? 128 : (TP + DV) <= 256 ? 256 : 512;
  static_assert(TP % 16 == 0 && TP <= 256 && DV % 32 == 0 && DV <= 256 && TP + DV <= 512, "bad tile");
};

// 8 fp32 -> one 16-byte chunk of fp16 hi and one of fp16 lo
__device__ __forceinline__ void split8(const float* x, uint4& hi, uint4& lo) {
  split2_f16(x[0], x[1], hi.x, lo.x);
  split2_f16(x[2], x[3], hi.y, lo.y);
  split2_f16(x[4], x[5], hi.z, lo.z);
  split2_f16(x[6], x[7], hi.w, lo.w);
}

}  // namespace

template <int TP, int DV>
__global__ void __launch_bounds__(NT, 2)
xattn_cell_tc_kernel(naf_xattn_params p, int rh, int rw) {
  using Cfg = TcCfg<TP, DV>;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t mbar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float red_m[2][128];
  __shared__ float red_l[2][128];

  uint8_t* sKhi = smem;
  uint8_t* sKlo = sKhi + Cfg::kSmemK;
  uint8_t* sVhi = sKlo + Cfg::kSmemK;
  uint8_t* sVlo = sVhi + Cfg::kSmemV;
  uint8_t* sQhi = sVlo + Cfg::kSmemV;
  uint8_t* sQlo = sQhi + Cfg::kSmemQ;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rowgrp = warp & 3;      // TMEM lane quarter this warp may touch
  const int half = warp >> 2;       // which half of the columns this thread owns
  const int row = rowgrp * 32 + lane;

  constexpr int K = WindowOf<TP>::K, K2 = K * K;
  int bid = blockIdx.x;
  const int head = bid % p.heads;
  bid /= p.heads;
  const int cj = bid % p.w;
  bid /= p.w;
  const int ci = bid % p.h;
  const int b = bid / p.h;
  const int wy0 = window_origin(ci, p.h, K);
  const int wx0 = window_origin(cj, p.w, K);

  if (warp == 0) tmem_alloc(&tmem_base_s, Cfg::kTmemCols);
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    fence_mbar_init();
  }

  // ---- stage the K window: canonical K-major [chunk c][tap n][16 B], zero rows for n >= K2
#pragma unroll 2
  for (int i = tid; i < TP * KC; i += NT) {
    const int n = i % TP, c = i / TP;
    uint4 hi = make_uint4(0, 0, 0, 0), lo = hi;
    if (n < K2) {
      const int t = n / K, u = n - t * K;
      const float* src = p.k + (int64_t(b * p.h + wy0 + t) * p.w + wx0 + u) * p.D + head * DQ + c * 8;
      float x[8];
      ldg8(src, x);
      split8(x, hi, lo);
    }
    *reinterpret_cast<uint4*>(sKhi + (c * TP + n) * 16) = hi;
    *reinterpret_cast<uint4*>(sKlo + (c * TP + n) * 16) = lo;
  }
  // ---- stage the V window: canonical MN-major [tap group][channel group][tap%8][16 B]
  {
    constexpr int NG = DV / 8;
#pragma unroll 4
    for (int i = tid; i < TP * NG; i += NT) {
      const int kk = i & 7;
      const int g = (i >> 3) % NG;
      const int kg = (i >> 3) / NG;
      const int k = kg * 8 + kk;
      uint4 hi = make_uint4(0, 0, 0, 0), lo = hi;
      if (k < K2) {
        const int t = k / K, u = k - t * K;
        const float* src = p.v + (int64_t(b * p.h + wy0 + t) * p.w + wx0 + u) * p.C + head * DV + g * 8;
        float x[8];
        ldg8(src, x);
        split8(x, hi, lo);
      }
      const int off = (kg * NG + g) * 128 + kk * 16;
      *reinterpret_cast<uint4*>(sVhi + off) = hi;
      *reinterpret_cast<uint4*>(sVlo + off) = lo;
    }
  }

  const bool rope = p.cos_y != nullptr;
  constexpr int HALF = DQ / 2, P = DQ / 4;     // P = 16 rotation pairs per axis
  const int npix = rh * rw;
  const int ntiles = (npix + 127) >> 7;
  const int y0 = ci * rh, x0 = cj * rw;
  const float qscale = p.scale * 1.4426950408889634f;  // fold log2(e): softmax via exp2
  // this thread owns rotation pairs [16*half, 16*half+16): channels a = [16h,16h+16), b = a+32
  const float* qbase = p.q + int64_t(b) * p.q_stride_b + head * DQ + P * half;
  float* obase = static_cast<float*>(p.out) + int64_t(b) * p.Ho * p.Wo * p.C + head * DV;

  float qa[P], qb[P];   // raw (un-rotated) prefetched q of the NEXT tile, then rotated in place
  int64_t my_pix = -1;  // linear target pixel of this thread's row, or -1
  int my_y = 0, my_x = 0;

  // issue the q loads of `tile` (no consumer: the latency overlaps the current tile's work)
  auto issue_q = [&](int tile) {
    int pi = tile * 128 + row;
    const bool valid = pi < npix;
    if (!valid) pi = npix - 1;
    const int py = pi / rw;
    my_y = y0 + py;
    my_x = x0 + (pi - py * rw);
    my_pix = valid ? int64_t(my_y) * p.Wo + my_x : -1;
    const float* qp = qbase + int64_t(my_y / p.rep_y) * p.q_stride_y + int64_t(my_x / p.rep_x) * p.q_stride_x;
#if NAF_TC_EXP & 2
    for (int j = 0; j < P; ++j) { qa[j] = float(j + tile) * 0.01f; qb[j] = float(j - row) * 0.02f; }
    (void)qp;
#else
    ldg_stream8(qp, *reinterpret_cast<float(*)[8]>(&qa[0]));
    ldg_stream8(qp + 8, *reinterpret_cast<float(*)[8]>(&qa[8]));
    ldg_stream8(qp + HALF, *reinterpret_cast<float(*)[8]>(&qb[0]));
    ldg_stream8(qp + HALF + 8, *reinterpret_cast<float(*)[8]>(&qb[8]));
#endif
  };
  // rotate + scale the prefetched q in place (tables are L1/L2 resident)
  auto finish_q = [&]() {
    if (rope) {
      // half 0 rotates by the row angles, half 1 by the column angles (src/layers/rope.py:139-143)
      const float* ct = half == 0 ? p.cos_y + int64_t(my_y) * P : p.cos_x + int64_t(my_x) * P;
      const float* st = half == 0 ? p.sin_y + int64_t(my_y) * P : p.sin_x + int64_t(my_x) * P;
      float c[P], s[P];
      ldg8(ct, *reinterpret_cast<float(*)[8]>(&c[0]));
      ldg8(ct + 8, *reinterpret_cast<float(*)[8]>(&c[8]));
      ldg8(st, *reinterpret_cast<float(*)[8]>(&s[0]));
      ldg8(st + 8, *reinterpret_cast<float(*)[8]>(&s[8]));
#pragma unroll
      for (int j = 0; j < P; ++j) {
        const float a = qa[j], bb = qb[j];
        qa[j] = (a * c[j] - bb * s[j]) * qscale;
        qb[j] = (bb * c[j] + a * s[j]) * qscale;
      }
    } else {
#pragma unroll
      for (int j = 0; j < P; ++j) {
        qa[j] *= qscale;
        qb[j] *= qscale;
      }
    }
  };

  issue_q(0);

  // wait for the TMEM base address and make barrier inits visible
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_off = uint32_t(rowgrp * 32) << 16;
  const uint32_t tS = tmem;            // S (fp32, TP cols)  /  P (fp16 hi: TP/2 cols | lo: TP/2 cols)
  const uint32_t tO = tmem + TP;       // O (fp32, DV cols)
  constexpr uint32_t idesc_qk = make_idesc_f16(128, TP, false, false);
  constexpr uint32_t idesc_pv = make_idesc_f16(128, DV, false, true);
  constexpr int SC = TP / 2;           // S columns (taps) per thread
  uint32_t phase = 0;

  for (int tile = 0; tile < ntiles; ++tile) {
    // ================= stage Q: canonical K-major [chunk][row][16 B]; this thread owns chunks
    // 2h, 2h+1 (a part) and 4+2h, 4+2h+1 (b part)
    finish_q();
    const int64_t cur_pix = my_pix;
    {
      uint4 hi, lo;
      split8(&qa[0], hi, lo);
      *reinterpret_cast<uint4*>(sQhi + ((2 * half) * 128 + row) * 16) = hi;
      *reinterpret_cast<uint4*>(sQlo + ((2 * half) * 128 + row) * 16) = lo;
      split8(&qa[8], hi, lo);
      *reinterpret_cast<uint4*>(sQhi + ((2 * half + 1) * 128 + row) * 16) = hi;
      *reinterpret_cast<uint4*>(sQlo + ((2 * half + 1) * 128 + row) * 16) = lo;
      split8(&qb[0], hi, lo);
      *reinterpret_cast<uint4*>(sQhi + ((4 + 2 * half) * 128 + row) * 16) = hi;
      *reinterpret_cast<uint4*>(sQlo + ((4 + 2 * half) * 128 + row) * 16) = lo;
      split8(&qb[8], hi, lo);
      *reinterpret_cast<uint4*>(sQhi + ((5 + 2 * half) * 128 + row) * 16) = hi;
      *reinterpret_cast<uint4*>(sQlo + ((5 + 2 * half) * 128 + row) * 16) = lo;
    }
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      // S = Qhi*Khi^T + Qlo*Khi^T + Qhi*Klo^T
#pragma unroll
      for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a0 = smem_u32(pass == 1 ? sQlo : sQhi);
        const uint32_t b0 = smem_u32(pass == 2 ? sKlo : sKhi);
#pragma unroll
        for (int kk = 0; kk < DQ / 16; ++kk) {
          const uint64_t da = make_desc(a0 + kk * 2 * (128 * 16), 128 * 16, 128);
          const uint64_t db = make_desc(b0 + kk * 2 * (TP * 16), TP * 16, 128);
#if !(NAF_TC_EXP & 4)
          mma_f16_ss(tS, da, db, idesc_qk, (pass | kk) != 0);
#endif
        }
      }
      commit(&mbar[0]);
    }
    // ================= prefetch the next tile's q while the tensor core works
    if (tile + 1 < ntiles) issue_q(tile + 1);

    mbar_wait(&mbar[0], phase);
    fence_after_sync();

    // ================= softmax on this thread's half of the S row
    {
      uint32_t s[SC];
      if constexpr (SC % 16 == 0) {
#pragma unroll
        for (int c0 = 0; c0 < SC; c0 += 16) tmem_ld16(tS + lane_off + half * SC + c0, *reinterpret_cast<uint32_t(*)[16]>(&s[c0]));
      } else {
#pragma unroll
        for (int c0 = 0; c0 < SC; c0 += 8) tmem_ld8(tS + lane_off + half * SC + c0, *reinterpret_cast<uint32_t(*)[8]>(&s[c0]));
      }
      wait_ld();
      const int tap0 = half * SC;
      float m = -INFINITY;
#pragma unroll
      for (int j = 0; j < SC; ++j)
        if (tap0 + j < K2) m = fmaxf(m, __uint_as_float(s[j]));
      red_m[half][row] = m;
      fence_before_sync();   // the S columns were read: P may overwrite them after the barrier
      __syncthreads();
      fence_after_sync();
      m = fmaxf(red_m[0][row], red_m[1][row]);
      float l = 0.f;
#pragma unroll
      for (int j = 0; j < SC; ++j) {
        const float e = (tap0 + j < K2) ? fast_exp2(__uint_as_float(s[j]) - m) : 0.f;
        l += e;
        s[j] = __float_as_uint(e);
      }
      red_l[half][row] = l;
      // P: hi halves in columns [0, TP/2), lo halves in [TP/2, TP); two taps per 32-bit column
      if constexpr (SC % 16 == 0) {
#pragma unroll
        for (int c0 = 0; c0 < SC; c0 += 16) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            split2_f16(__uint_as_float(s[c0 + 2 * j]), __uint_as_float(s[c0 + 2 * j + 1]), hi[j], lo[j]);
          tmem_st8(tS + lane_off + (tap0 + c0) / 2, hi);
          tmem_st8(tS + lane_off + TP / 2 + (tap0 + c0) / 2, lo);
        }
      } else {
#pragma unroll
        for (int c0 = 0; c0 < SC; c0 += 8) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            split2_f16(__uint_as_float(s[c0 + 2 * j]), __uint_as_float(s[c0 + 2 * j + 1]), hi[j], lo[j]);
          tmem_st4(tS + lane_off + (tap0 + c0) / 2, hi);
          tmem_st4(tS + lane_off + TP / 2 + (tap0 + c0) / 2, lo);
        }
      }
      wait_st();
    }
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      // O = Phi*Vhi + Plo*Vhi + Phi*Vlo
#pragma unroll
      for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a0 = tS + (pass == 1 ? TP / 2 : 0);
        const uint32_t b0 = smem_u32(pass == 2 ? sVlo : sVhi);
#pragma unroll
        for (int kk = 0; kk < TP / 16; ++kk) {
          const uint64_t db = make_desc(b0 + kk * 2 * (DV / 8) * 128, (DV / 8) * 128, 128);
#if !(NAF_TC_EXP & 4)
          mma_f16_ts(tO, a0 + kk * 8, db, idesc_pv, (pass | kk) != 0);
#endif
        }
      }
      commit(&mbar[1]);
    }
    const float inv_l = 1.f / (red_l[0][row] + red_l[1][row]);
    mbar_wait(&mbar[1], phase);
    fence_after_sync();

    // ================= epilogue: this thread's 32-column chunks of its O row -> normalise ->
    // 256-bit streaming stores (chunks c with c % 2 == half)
    {
      float* orow = obase + cur_pix * p.C;
      constexpr int NCH = DV / 32;                       // 32-column chunks of the O row
      const int nmine = (NCH + 1 - half) / 2;            // chunks half, half+2, ...
      uint32_t r[2][32];
      if (nmine > 0) tmem_ld32(tO + lane_off + half * 32, r[0]);
#pragma unroll
      for (int i = 0; i < (NCH + 1) / 2; ++i) {
        if (i < nmine) {
          wait_ld();
          // keep one TMEM load in flight while this chunk is normalised and stored
          if (i + 1 < nmine) tmem_ld32(tO + lane_off + (half + 2 * (i + 1)) * 32, r[(i + 1) & 1]);
          if (cur_pix >= 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float o[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(r[i & 1][j + e]) * inv_l;
#if NAF_TC_EXP & 1
              if (o[0] == 123.456f) stg_stream8(orow + (half + 2 * i) * 32 + j, o);
#else
              stg_stream8(orow + (half + 2 * i) * 32 + j, o);
#endif
            }
          }
        }
      }
      wait_ld();
    }
    fence_before_sync();  // O / S columns are rewritten by the next tile's MMAs (after its barriers)
    phase ^= 1;
  }

  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, Cfg::kTmemCols);
}

// ------------------------------------------------------------------------------------ host side
namespace {

int taps_pad(int K) {
  const int k2 = K * K;
  return (k2 + 15) / 16 * 16;
}

template <int TP, int DV>
int launch_tc(const naf_xattn_params& p, cudaStream_t st) {
  using Cfg = TcCfg<TP, DV>;
  auto kern = xattn_cell_tc_kernel<TP, DV>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemTotal);
  if (e != cudaSuccess)
    return fail(NAF_ERR_CUDA, "xattn(cell-tc): smem opt-in failed: %s", cudaGetErrorString(e));
  const unsigned grid = unsigned(p.B) * p.h * p.w * p.heads;
  kern<<<grid, NT, Cfg::kSmemTotal, st>>>(p, p.Ho / p.h, p.Wo / p.w);
  return check_launch("xattn_cell_tc");
}

template <int TP>
int launch_tc_dv(const naf_xattn_params& p, cudaStream_t st) {
  switch (p.C / p.heads) {
    case 32: return launch_tc<TP, 32>(p, st);
    case 64: return launch_tc<TP, 64>(p, st);
    case 96: return launch_tc<TP, 96>(p, st);
    case 128: return launch_tc<TP, 128>(p, st);
    case 192: return launch_tc<TP, 192>(p, st);
    case 256: return launch_tc<TP, 256>(p, st);
    default: return fail(NAF_ERR_UNSUPPORTED, "xattn(cell-tc): value head dim %d", p.C / p.heads);
  }
}

}  // namespace

bool xattn_cell_tc_supported(const naf_xattn_params& p, const char** why) {
  if (p.out_dtype != NAF_DTYPE_F32) { *why = "fp32 output only"; return false; }
  const int dq = p.D / p.heads, dv = p.C / p.heads;
  if (p.row_tap || p.col_tap) { *why = "tap tables given (non-integer ratio path)"; return false; }
  if (p.Ho % p.h || p.Wo % p.w) { *why = "target size is not a multiple of the feature size"; return false; }
  if (p.scores) { *why = "score output requested"; return false; }
  if (dq != DQ) { *why = "head dim must be 64"; return false; }
  if (dv != 32 && dv != 64 && dv != 96 && dv != 128 && dv != 192 && dv != 256) {
    *why = "value head dim must be one of 32/64/96/128/192/256";
    return false;
  }
  const int tp = taps_pad(p.K);
  if (p.K < 3 || (tp != 16 && tp != 32 && tp != 64 && tp != 96 && tp != 128)) {
    *why = "kernel_size must be 3, 5, 7, 9 or 11";
    return false;
  }
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
  if (int64_t(p.B) * p.h * p.w * p.heads >= (int64_t(1) << 31)) { *why = "grid too large"; return false; }
  return true;
}

int launch_xattn_cell_tc(const naf_xattn_params& p, cudaStream_t st) {
  switch (taps_pad(p.K)) {
    case 16: return launch_tc_dv<16>(p, st);
    case 32: return launch_tc_dv<32>(p, st);
    case 64: return launch_tc_dv<64>(p, st);
    case 96: return launch_tc_dv<96>(p, st);
    case 128: return launch_tc_dv<128>(p, st);
    default: return fail(NAF_ERR_UNSUPPORTED, "xattn(cell-tc): kernel_size %d", p.K);
  }
}

}  // namespace naf
