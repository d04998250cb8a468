// Backward of the cross-scale neighbourhood attention on the tensor core (tcgen05 / TMEM), integer ratios,
// 64-wide heads (reference: autograd through legacy_attention, src/layers/attentions.py:16-29,
// as train.py:136 / test/backward_speed.py:51-64 run it).
//
// One CTA owns one (batch, low-res cell, head); all r*r target pixels of the cell share one K x K window, so
// per tile of 128 pixels (one TMEM lane each) the whole backward is five dense GEMMs (TP = K*K padded to 16):
//     S   [128 x TP] = Q  [128 x 64] Kw^T           recomputed scores (nothing is saved by the forward)
//     dP  [128 x TP] = G  [128 x dv] Vw^T           G = dL/dout
//     P = softmax(scale S),  delta = sum_t P dP,  dS = scale P (dP - delta)          two threads per pixel row
//     dQ  [128 x 64] = dS [128 x TP] Kw             -> global (gradient w.r.t. the rotated query)
//     dKw [TP x 64] += dS^T [TP x 128] Q            accumulated in TMEM over ALL tiles of the cell
//     dVw [TP x dv] += P^T  [TP x 128] G            accumulated in TMEM over ALL tiles of the cell
// and the window gradients leave with ONE vector atomic per 4 window elements per cell.  Shared-memory
// operand images are written once and read through two descriptor forms: Q and G are K-major A operands of
// S / dP and MN-major B operands of dKw / dVw; Kw is the K-major B of S and the MN-major B of dQ.  P^T and
// dS^T are MN-major A operands (a pixel writes 8 consecutive taps as one 16-byte chunk).
//
// Precision: every GEMM runs the three-pass split (fp16 hi/lo operands, hi*hi + lo*hi + hi*lo, fp32
// accumulation, ~22 mantissa bits): measured <= 3e-6 of the gradient's largest entry vs the fp64 oracle.  (A
// single-fp16 P^T / dS^T was measured at 2e-4: with sign-alternating upstream gradients the rounding errors
// of the summed pixels do not average out relative to the sum.)  P^T and dS^T share one buffer (P^T is dead
// once the last dVw chunk has been issued and completed).
// The value head is walked in chunks of 64 channels (G and Vw chunks are staged together: one pass over G).
#include "naf_common.cuh"
#include "naf_umma.cuh"

#ifndef NAF_BWD_EXP
#define NAF_BWD_EXP 0   // profiling experiments only (scripts/gpu_bwd_exp.sh); 0 in every real build
#endif

namespace naf {

using namespace umma;

namespace {

constexpr int DQ = 64;        // head dim of q / k
constexpr int KC = DQ / 8;    // 16-byte chunks along a 64-channel operand row
constexpr int NT = 256;
constexpr int DVC = 64;       // value channels per chunk
constexpr int PIXIMG = KC * 128 * 16;   // one fp16 image of a 128 x 64 pixel-major operand (16 KB)
constexpr int TIMG = 128 * 128 * 2;     // P^T / dS^T image: 128 taps x 128 pixels fp16 (32 KB)

template <int TP>
struct BwdCfg {
  static constexpr int kWin = KC * TP * 16;   // one fp16 image of a TP x 64 window operand
  static constexpr int kSmem = 4 * kWin + 4 * PIXIMG + 2 * TIMG;   // K, V chunk (hi, lo); Q, G chunk (hi, lo); T (hi, lo)
  static constexpr int kTA = (TP + 31) / 32 * 32;                   // TMEM regions start on 32-column boundaries
};

__device__ __forceinline__ void red_add4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ void split8(const float (&x)[8], uint4& hi, uint4& lo) {
  split2_f16(x[0], x[1], hi.x, lo.x);
  split2_f16(x[2], x[3], hi.y, lo.y);
  split2_f16(x[4], x[5], hi.z, lo.z);
  split2_f16(x[6], x[7], hi.w, lo.w);
}

}  // namespace

template <int TP>
__global__ void __launch_bounds__(NT, 1)
xattn_bwd_cell_tc_kernel(naf_xattn_bwd_params p, int rh, int rw, int dv, int v_resident) {
  using Cfg = BwdCfg<TP>;
  constexpr int SC = TP / 2;   // S / dP columns per thread
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float red_a[2][128];
  __shared__ float red_b[2][128];

  uint8_t* sKhi = smem;
  uint8_t* sKlo = sKhi + Cfg::kWin;
  // V window: [channel group][tap][16 B]; the whole value head when it fits (staged once per cell), else one
  // 64-channel chunk restaged per tile and chunk
  const int v_img = (v_resident ? (dv >> 3) : KC) * TP * 16;
  uint8_t* sVhi = sKlo + Cfg::kWin;
  uint8_t* sVlo = sVhi + v_img;
  uint8_t* sQhi = sVlo + v_img;
  uint8_t* sQlo = sQhi + PIXIMG;
  uint8_t* sGhi = sQlo + PIXIMG;
  uint8_t* sGlo = sGhi + PIXIMG;
  uint8_t* sThi = sGlo + PIXIMG;      // P^T, then dS^T: MN-major A images [pixel group][tap group][pixel % 8][8 taps]
  uint8_t* sTlo = sThi + TIMG;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rowgrp = warp & 3, hf = warp >> 2;
  const int row = rowgrp * 32 + lane;
  const int K = p.K, K2 = K * K;

  int bid = blockIdx.x;
  const int head = bid % p.heads;
  bid /= p.heads;
  const int cj = bid % p.w;
  bid /= p.w;
  const int ci = bid % p.h;
  const int b = bid / p.h;
  const int wy0 = window_origin(ci, p.h, K), wx0 = window_origin(cj, p.w, K);

  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    mbar_init(&mbar, 1);
    fence_mbar_init();
  }
  // P^T / dS^T images: tap rows >= TP are never written; they must read as zeros (M = 128 always)
  for (int i = tid; i < 2 * TIMG / 16; i += NT) reinterpret_cast<uint4*>(sThi)[i] = make_uint4(0, 0, 0, 0);
  // K window: K-major [channel chunk c][tap n][16 B], zero rows for the padding taps
  const float* kwin = p.k + (int64_t(b * p.h + wy0) * p.w + wx0) * p.D + head * DQ;
  for (int i = tid; i < TP * KC; i += NT) {
    const int n = i % TP, c = i / TP;
    uint4 hi = make_uint4(0, 0, 0, 0), lo = hi;
    if (n < K2) {
      const int t = n / K, u = n - t * K;
      float x[8];
      const float4* src = reinterpret_cast<const float4*>(kwin + (int64_t(t) * p.w + u) * p.D + c * 8);
      const float4 a = __ldg(src), bb = __ldg(src + 1);
      x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = bb.x; x[5] = bb.y; x[6] = bb.z; x[7] = bb.w;
      split8(x, hi, lo);
    }
    *reinterpret_cast<uint4*>(sKhi + (c * TP + n) * 16) = hi;
    *reinterpret_cast<uint4*>(sKlo + (c * TP + n) * 16) = lo;
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();

  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_off = uint32_t(rowgrp * 32) << 16;
  const uint32_t tS = tmem;                  // S, later dS (fp16 hi | lo)
  const uint32_t tdP = tmem + Cfg::kTA;      // dP, later dQ (64 columns)
  const uint32_t tdK = tdP + (Cfg::kTA > DQ ? Cfg::kTA : DQ);   // dKw accumulator (taps x 64)
  const uint32_t tdV = tdK + DQ;             // dVw accumulator (taps x dv)

  const bool rope = p.cos_y != nullptr;
  const float qscale = p.scale * 1.4426950408889634f;
  const int npix = rh * rw;
  const int ntiles = (npix + 127) >> 7;
  const int nchunks = (dv + DVC - 1) / DVC;
  const int y0 = ci * rh, x0 = cj * rw;
  const float* vwin = p.v + (int64_t(b * p.h + wy0) * p.w + wx0) * p.C + head * dv;
  uint32_t phase = 0;
  constexpr uint32_t idesc_s = make_idesc_f16(128, TP, false, false);     // S, dP: A K-major, B K-major
  constexpr uint32_t idesc_dq = make_idesc_f16(128, DQ, false, true);     // dQ: A in TMEM, B MN-major
  constexpr uint32_t idesc_dk = make_idesc_f16(128, DQ, true, true);      // dKw: A MN-major, B MN-major

  // ---- software pipeline over HBM latency: the raw q of the next tile and the upstream-gradient chunk of the
  // next (tile, chunk) step are loaded into registers while the current step converts / multiplies
  struct Px { bool valid; int y, x; int64_t pix; };
  auto px_of = [&](int tile) {
    Px r;
    const int pi = tile * 128 + row;
    r.valid = tile < ntiles && pi < npix;
    const int py = r.valid ? pi / rw : 0;
    r.y = y0 + py;
    r.x = x0 + (r.valid ? pi - py * rw : 0);
    r.pix = (int64_t(b) * p.Ho + r.y) * p.Wo + r.x;
    return r;
  };
  float4 qa4[4], qb4[4];     // hf 0 owns rotation pairs 0..15 (channels 0..15 / 32..47, row angles), hf 1 pairs 16..31
  auto load_q = [&](const Px& px) {
    if (px.valid) {
      const float* qp = p.q + int64_t(b) * p.q_stride_b + int64_t(px.y / p.rep_y) * p.q_stride_y +
                        int64_t(px.x / p.rep_x) * p.q_stride_x + head * DQ + 16 * hf;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        qa4[j] = ldg_stream(qp + 4 * j);
        qb4[j] = ldg_stream(qp + 32 + 4 * j);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) qa4[j] = qb4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  float4 g4[4][2];           // channel groups hf, hf + 2, hf + 4, hf + 6 of this pixel's chunk
  auto load_g = [&](const Px& px, int c) {
    const int ng = min(DVC, dv - c * DVC) >> 3;
    const float* gp = p.dout + px.pix * p.C + head * dv + c * DVC;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int g8 = hf + 2 * k;
      if (px.valid && g8 < ng && !(NAF_BWD_EXP & 16)) {
        g4[k][0] = ldg_stream(gp + g8 * 8);
        g4[k][1] = ldg_stream(gp + g8 * 8 + 4);
      } else {
        g4[k][0] = g4[k][1] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  };
  Px cur = px_of(0);
  load_q(cur);
  load_g(cur, 0);

  for (int tile = 0; tile < ntiles; ++tile) {
    const Px nxt = px_of(tile + 1);
    // ================= stage Q (rotated, unscaled): K-major [chunk][row][16 B]
    {
      float qa[16], qb[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        qa[4 * j] = qa4[j].x; qa[4 * j + 1] = qa4[j].y; qa[4 * j + 2] = qa4[j].z; qa[4 * j + 3] = qa4[j].w;
        qb[4 * j] = qb4[j].x; qb[4 * j + 1] = qb4[j].y; qb[4 * j + 2] = qb4[j].z; qb[4 * j + 3] = qb4[j].w;
      }
      if (rope && cur.valid) {
        const float4* ct = reinterpret_cast<const float4*>(hf == 0 ? p.cos_y + int64_t(cur.y) * 16 : p.cos_x + int64_t(cur.x) * 16);
        const float4* st = reinterpret_cast<const float4*>(hf == 0 ? p.sin_y + int64_t(cur.y) * 16 : p.sin_x + int64_t(cur.x) * 16);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 c4 = __ldg(ct + j), s4 = __ldg(st + j);
          const float c[4] = {c4.x, c4.y, c4.z, c4.w}, sn[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float a = qa[4 * j + e], bb = qb[4 * j + e];
            qa[4 * j + e] = a * c[e] - bb * sn[e];
            qb[4 * j + e] = bb * c[e] + a * sn[e];
          }
        }
      }
      uint4 hi, lo;
      split8(*reinterpret_cast<float(*)[8]>(&qa[0]), hi, lo);
      *reinterpret_cast<uint4*>(sQhi + ((2 * hf) * 128 + row) * 16) = hi;
      *reinterpret_cast<uint4*>(sQlo + ((2 * hf) * 128 + row) * 16) = lo;
      split8(*reinterpret_cast<float(*)[8]>(&qa[8]), hi, lo);
      *reinterpret_cast<uint4*>(sQhi + ((2 * hf + 1) * 128 + row) * 16) = hi;
      *reinterpret_cast<uint4*>(sQlo + ((2 * hf + 1) * 128 + row) * 16) = lo;
      split8(*reinterpret_cast<float(*)[8]>(&qb[0]), hi, lo);
      *reinterpret_cast<uint4*>(sQhi + ((4 + 2 * hf) * 128 + row) * 16) = hi;
      *reinterpret_cast<uint4*>(sQlo + ((4 + 2 * hf) * 128 + row) * 16) = lo;
      split8(*reinterpret_cast<float(*)[8]>(&qb[8]), hi, lo);
      *reinterpret_cast<uint4*>(sQhi + ((5 + 2 * hf) * 128 + row) * 16) = hi;
      *reinterpret_cast<uint4*>(sQlo + ((5 + 2 * hf) * 128 + row) * 16) = lo;
    }
    load_q(nxt);     // in flight during the whole tile
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      // S = Qhi Khi^T + Qlo Khi^T + Qhi Klo^T
#pragma unroll
      for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a0 = smem_u32(pass == 1 ? sQlo : sQhi);
        const uint32_t b0 = smem_u32(pass == 2 ? sKlo : sKhi);
#pragma unroll
        for (int ks = 0; ks < DQ / 16; ++ks)
          mma_f16_ss(tS, make_desc(a0 + ks * 2 * (128 * 16), 128 * 16, 128),
                     make_desc(b0 + ks * 2 * (TP * 16), TP * 16, 128), idesc_s, (pass | ks) != 0);
      }
      commit(&mbar);
    }
    mbar_wait(&mbar, phase);
    phase ^= 1;
    fence_after_sync();

    // ================= P = softmax(scale S) on this thread's half row; P^T (fp16 hi / lo) -> shared memory
    float pr[SC];
    const int tap0 = hf * SC;
    uint8_t* const tdst = sThi + (row >> 3) * 2048 + (tap0 >> 3) * 128 + (row & 7) * 16;
    {
      uint32_t s[SC];
#pragma unroll
      for (int c0 = 0; c0 < SC; c0 += 8) tmem_ld8(tS + lane_off + hf * SC + c0, *reinterpret_cast<uint32_t(*)[8]>(&s[c0]));
      wait_ld();
      float m = -INFINITY;
#pragma unroll
      for (int j = 0; j < SC; ++j)
        if (tap0 + j < K2) m = fmaxf(m, __uint_as_float(s[j]));
      red_a[hf][row] = m;
      fence_before_sync();
      __syncthreads();
      fence_after_sync();
      const float mq = fmaxf(red_a[0][row], red_a[1][row]) * qscale;
      float l = 0.f;
#pragma unroll
      for (int j = 0; j < SC; ++j) {
        const float e = (tap0 + j < K2) ? fast_exp2(fmaf(__uint_as_float(s[j]), qscale, -mq)) : 0.f;
        pr[j] = e;
        l += e;
      }
      red_b[hf][row] = l;
      __syncthreads();
      const float inv = 1.f / (red_b[0][row] + red_b[1][row]);
#pragma unroll
      for (int j = 0; j < SC; ++j) pr[j] *= inv;
#pragma unroll
      for (int g8 = 0; g8 < SC / 8; ++g8) {
        uint4 hi, lo;
        split8(*reinterpret_cast<float(*)[8]>(&pr[8 * g8]), hi, lo);
        *reinterpret_cast<uint4*>(tdst + g8 * 128) = hi;
        *reinterpret_cast<uint4*>(tdst + TIMG + g8 * 128) = lo;
      }
    }

    // ================= value-head chunks: dP += G_c Vw_c^T,  dVw[:, chunk c] += P^T G_c
    for (int c = 0; c < nchunks; ++c) {
      const int wc = min(DVC, dv - c * DVC);          // channels in this chunk (multiple of 16)
      // G chunk (prefetched): K-major [channel group][row][16 B]
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int g8 = hf + 2 * k;
        if (g8 < (wc >> 3)) {
          const float x[8] = {g4[k][0].x, g4[k][0].y, g4[k][0].z, g4[k][0].w, g4[k][1].x, g4[k][1].y, g4[k][1].z, g4[k][1].w};
          uint4 hi, lo;
          split8(x, hi, lo);
          *reinterpret_cast<uint4*>(sGhi + (g8 * 128 + row) * 16) = hi;
          *reinterpret_cast<uint4*>(sGlo + (g8 * 128 + row) * 16) = lo;
        }
      }
      if (c + 1 < nchunks) load_g(cur, c + 1);
      else if (tile + 1 < ntiles) load_g(nxt, 0);
      // V window chunk: K-major [channel group][tap][16 B] (once per cell when the head is one chunk wide)
      uint8_t* const vhi = sVhi + (v_resident ? c * (DVC / 8) * TP * 16 : 0);
      uint8_t* const vlo = sVlo + (v_resident ? c * (DVC / 8) * TP * 16 : 0);
      if (tile == 0 || !v_resident) {
        constexpr int VIT = (TP * 8 + NT - 1) / NT;     // items per thread at the full chunk width
        float4 va[VIT], vb[VIT];
#pragma unroll
        for (int k = 0; k < VIT; ++k) {
          const int i = tid + k * NT;
          const int n = i % TP, g8 = i / TP;
          va[k] = vb[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (g8 < (wc >> 3) && n < K2) {
            const int t = n / K, u = n - t * K;
            const float4* src = reinterpret_cast<const float4*>(vwin + (int64_t(t) * p.w + u) * p.C + c * DVC + g8 * 8);
            va[k] = __ldg(src);
            vb[k] = __ldg(src + 1);
          }
        }
#pragma unroll
        for (int k = 0; k < VIT; ++k) {
          const int i = tid + k * NT;
          const int n = i % TP, g8 = i / TP;
          if (g8 < (wc >> 3)) {
            const float x[8] = {va[k].x, va[k].y, va[k].z, va[k].w, vb[k].x, vb[k].y, vb[k].z, vb[k].w};
            uint4 hi, lo;
            split8(x, hi, lo);
            *reinterpret_cast<uint4*>(vhi + (g8 * TP + n) * 16) = hi;
            *reinterpret_cast<uint4*>(vlo + (g8 * TP + n) * 16) = lo;
          }
        }
      }
      fence_proxy_async_smem();
      fence_before_sync();
      __syncthreads();
      if (tid == 0) {
        fence_after_sync();
        // dP (+)= Ghi Vhi^T + Glo Vhi^T + Ghi Vlo^T
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t a0 = smem_u32(pass == 1 ? sGlo : sGhi);
          const uint32_t b0 = smem_u32(pass == 2 ? vlo : vhi);
          for (int ks = 0; ks < (wc >> 4); ++ks)
            if (!(NAF_BWD_EXP & 4)) mma_f16_ss(tdP, make_desc(a0 + ks * 2 * (128 * 16), 128 * 16, 128),
                       make_desc(b0 + ks * 2 * (TP * 16), TP * 16, 128), idesc_s, (c | pass | ks) != 0);
        }
        // dVw[:, chunk] (+)= P^T_hi Ghi + P^T_lo Ghi + P^T_hi Glo   (A = P^T MN-major, B = G seen MN-major: K = pixels)
        const uint32_t idesc_dv = make_idesc_f16(128, wc, true, true);
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t a0 = smem_u32(pass == 1 ? sTlo : sThi);
          const uint32_t b0 = smem_u32(pass == 2 ? sGlo : sGhi);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            if (!(NAF_BWD_EXP & 1)) mma_f16_ss(tdV + c * DVC, make_desc(a0 + ks * 2 * 2048, 2048, 128),
                       make_desc(b0 + ks * 2 * 128, 128, 128 * 16), idesc_dv, (tile | pass | ks) != 0);
        }
        commit(&mbar);
      }
      mbar_wait(&mbar, phase);
      phase ^= 1;
      fence_after_sync();
    }

    // ================= dS = scale P (dP - delta): TMEM (A operand of dQ, hi | lo) and dS^T -> the T buffer
    {
      uint32_t d[SC];
#pragma unroll
      for (int c0 = 0; c0 < SC; c0 += 8) tmem_ld8(tdP + lane_off + hf * SC + c0, *reinterpret_cast<uint32_t(*)[8]>(&d[c0]));
      wait_ld();
      float dl = 0.f;
#pragma unroll
      for (int j = 0; j < SC; ++j) dl = fmaf(pr[j], __uint_as_float(d[j]), dl);
      red_a[hf][row] = dl;
      fence_before_sync();   // dP was read: dQ may overwrite it after the barriers below
      __syncthreads();
      fence_after_sync();
      const float delta = red_a[0][row] + red_a[1][row];
#pragma unroll
      for (int j = 0; j < SC; ++j) pr[j] = pr[j] * (__uint_as_float(d[j]) - delta) * p.scale;
#pragma unroll
      for (int c0 = 0; c0 < SC; c0 += 8) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split2_f16(pr[c0 + 2 * j], pr[c0 + 2 * j + 1], hi[j], lo[j]);
        tmem_st4(tS + lane_off + (tap0 + c0) / 2, hi);
        tmem_st4(tS + lane_off + TP / 2 + (tap0 + c0) / 2, lo);
        *reinterpret_cast<uint4*>(tdst + (c0 >> 3) * 128) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(tdst + TIMG + (c0 >> 3) * 128) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      wait_st();
    }
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      // dQ = dShi Khi + dSlo Khi + dShi Klo      (B = the K window seen MN-major: N = channels, K = taps)
#pragma unroll
      for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a0 = tS + (pass == 1 ? TP / 2 : 0);
        const uint32_t b0 = smem_u32(pass == 2 ? sKlo : sKhi);
#pragma unroll
        for (int ks = 0; ks < TP / 16; ++ks)
          mma_f16_ts(tdP, a0 + ks * 8, make_desc(b0 + ks * 2 * 128, 128, TP * 16), idesc_dq, (pass | ks) != 0);
      }
      // dKw (+)= dS^T_hi Qhi + dS^T_lo Qhi + dS^T_hi Qlo
#pragma unroll
      for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a0 = smem_u32(pass == 1 ? sTlo : sThi);
        const uint32_t b0 = smem_u32(pass == 2 ? sQlo : sQhi);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          if (!(NAF_BWD_EXP & 2)) mma_f16_ss(tdK, make_desc(a0 + ks * 2 * 2048, 2048, 128), make_desc(b0 + ks * 2 * 128, 128, 128 * 16),
                     idesc_dk, (tile | pass | ks) != 0);
      }
      commit(&mbar);
    }
    mbar_wait(&mbar, phase);
    phase ^= 1;
    fence_after_sync();

    // ================= dQ -> global: this thread's 32 columns of its pixel row
    {
      uint32_t r[32];
      tmem_ld32(tdP + lane_off + hf * 32, r);
      wait_ld();
      if (cur.valid) {
        float* dst = p.dq + cur.pix * p.D + head * DQ + hf * 32;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          stg_stream(dst + 4 * j, make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                              __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])));
      }
    }
    fence_before_sync();   // ordered before the next tile's MMAs by its barriers
    cur = nxt;
  }

  // ================= window gradients: TMEM lane = tap; one vector atomic per 4 elements
  {
    const bool live = row < K2 && !(NAF_BWD_EXP & 8);
    const int t = live ? row / K : 0, u = live ? row - (row / K) * K : 0;
    const int64_t cell = int64_t(b * p.h + wy0 + t) * p.w + wx0 + u;
    uint32_t r[32];
    tmem_ld32(tdK + lane_off + hf * 32, r);
    wait_ld();
    if (live) {
      float* dst = p.dk + cell * p.D + head * DQ + hf * 32;
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        red_add4(dst + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
    }
    for (int c16 = hf; c16 < (dv >> 4); c16 += 2) {
      uint32_t o[16];
      tmem_ld16(tdV + lane_off + c16 * 16, o);
      wait_ld();
      if (live) {
        float* dst = p.dv + cell * p.C + head * dv + c16 * 16;
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          red_add4(dst + j, __uint_as_float(o[j]), __uint_as_float(o[j + 1]), __uint_as_float(o[j + 2]), __uint_as_float(o[j + 3]));
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------ host side
namespace {

int taps_pad(int K) { return (K * K + 15) / 16 * 16; }

template <int TP>
int launch_bwd_tc(const naf_xattn_bwd_params& p, cudaStream_t st) {
  auto kern = xattn_bwd_cell_tc_kernel<TP>;
  const int dv = p.C / p.heads;
  // shared memory: K window + V region + Q, G chunk, T (hi / lo each).  The V region holds the whole value head
  // when that fits beside the rest (K <= 7 always; K = 9 up to dv = 176), else one 64-channel chunk.
  const int fixed = 2 * BwdCfg<TP>::kWin + 4 * PIXIMG + 2 * TIMG;
  const int v_full = 2 * (dv / 8) * TP * 16;
  const int v_resident = fixed + v_full <= 220 * 1024 ? 1 : 0;
  const int smem = fixed + (v_resident ? v_full : 2 * BwdCfg<TP>::kWin);
  cudaError_t e = ensure_dyn_smem(kern, smem);
  if (e != cudaSuccess) return fail(NAF_ERR_CUDA, "xattn_bwd(cell-tc): smem opt-in failed: %s", cudaGetErrorString(e));
  const unsigned grid = unsigned(p.B) * p.h * p.w * p.heads;
  kern<<<grid, NT, smem, st>>>(p, p.Ho / p.h, p.Wo / p.w, dv, v_resident);
  return check_launch("xattn_bwd_cell_tc");
}

}  // namespace

bool xattn_bwd_cell_tc_supported(const naf_xattn_bwd_params& p, const char** why) {
  if (p.row_tap || p.col_tap || p.Ho % p.h || p.Wo % p.w) { *why = "tap tables / non-integer ratio"; return false; }
  const int dq = p.D / p.heads, dv = p.C / p.heads, tp = taps_pad(p.K);
  if (dq != DQ) { *why = "head dim must be 64"; return false; }
  if (dv % 16 || dv < 16 || dv > 256) { *why = "value head dim must be a multiple of 16 in [16, 256]"; return false; }
  if (p.K < 3 || tp > 128) { *why = "kernel_size must be 3..11"; return false; }
  const int tpa = (tp + 31) / 32 * 32;
  if (tpa + (tpa > DQ ? tpa : DQ) + DQ + dv > 512) { *why = "window and value head need more than 512 TMEM columns"; return false; }
  if ((p.Ho / p.h) * (p.Wo / p.w) < 64) { *why = "fewer than 64 pixels per cell"; return false; }
  if (!aligned16(p.q) || !aligned16(p.k) || !aligned16(p.v) || !aligned16(p.dout) || !aligned16(p.dq) ||
      !aligned16(p.dk) || !aligned16(p.dv) || (p.q_stride_b % 4) || (p.q_stride_y % 4) || (p.q_stride_x % 4)) {
    *why = "pointers / strides not 16-byte aligned";
    return false;
  }
  if (p.cos_y && !(aligned16(p.cos_y) && aligned16(p.sin_y) && aligned16(p.cos_x) && aligned16(p.sin_x))) {
    *why = "rope tables not 16-byte aligned";
    return false;
  }
  if (int64_t(p.B) * p.h * p.w * p.heads >= (int64_t(1) << 31)) { *why = "grid too large"; return false; }
  return true;
}

int launch_xattn_bwd_cell_tc(const naf_xattn_bwd_params& p, cudaStream_t st) {
  switch (taps_pad(p.K)) {
    case 16: return launch_bwd_tc<16>(p, st);
    case 32: return launch_bwd_tc<32>(p, st);
    case 64: return launch_bwd_tc<64>(p, st);
    case 96: return launch_bwd_tc<96>(p, st);
    case 128: return launch_bwd_tc<128>(p, st);
    default: return fail(NAF_ERR_UNSUPPORTED, "xattn_bwd(cell-tc): kernel_size %d", p.K);
  }
}

}  // namespace naf
