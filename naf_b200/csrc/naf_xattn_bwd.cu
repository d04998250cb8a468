// Backward of the cross-scale neighbourhood attention.
//
// Reference: autograd through legacy_attention (src/layers/attentions.py:16-29: na2d_qk -> *scale ->
// softmax -> na2d_av; NATTEN provides the two backward functionals) as train.py:136 and
// test/backward_speed.py:51-64 run it, and through KeyEncoder / RoPE (src/model/naf.py:63-69,
// src/layers/rope.py:137-153) for the fused entry.
//
// With  s_t = scale * <q, k_t>,  P = softmax_t(s),  o = sum_t P_t v_t  for the K*K taps t of a pixel and
// g = dL/do:
//     dP_t = <g, v_t>              delta = sum_t P_t dP_t  (= <g, o>)
//     dS_t = P_t (dP_t - delta)
//     dq   = scale * sum_t dS_t k_t
//     dk_{cell(t)} += scale * dS_t q        (scatter-add over the pixels whose windows hold the cell)
//     dv_{cell(t)} += P_t g
// Nothing from the forward is kept (the scores are recomputed from q, k): the activations the reference
// saves for backward (the (B,n,Ho,Wo,K*K) attention tensor and the replicated K, V) never exist here.
//
//   xattn_bwd_generic_kernel : one warp per (pixel, head); tap tables or the integer rule; atomics into
//                              dk / dv.  The path for non-integer ratios (training: 32 <- 13) and ratio 1.
//   xattn_bwd_cell_kernel    : integer ratios: one CTA per (batch, cell, head); the r*r pixels of a cell
//                              share ONE window, so its dK / dV contributions are reduced in registers
//                              over the whole cell first and leave with K*K*(dq+dv) atomics per cell
//                              instead of per pixel.
//   rope_kpool_bwd_kernel    : dx = R^T (dq + sum_{bins containing the pixel} dk_bin / |bin|): gradient of
//                              the key pooling and of the rotation, in one pass (in place over dq).
#include "naf_common.cuh"

namespace naf {

namespace {

constexpr int kBwdWarps = 4;

__device__ __forceinline__ void red_add(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

}  // namespace

__global__ void __launch_bounds__(kBwdWarps * 32)
xattn_bwd_generic_kernel(naf_xattn_bwd_params p, int rh, int rw, int64_t total_items) {
  extern __shared__ float smem_f[];
  const int Kh = p.K, Kw = p.Kw ? p.Kw : p.K;
  const int K2 = Kh * Kw;
  const int dq = p.D / p.heads, dv = p.C / p.heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_warp = 3 * K2 + dq + dv;
  float* sc = smem_f + size_t(warp) * per_warp;           // probabilities              [K2]
  float* ds = sc + K2;                                     // scale * dS                 [K2]
  int* idx = reinterpret_cast<int*>(ds + K2);              // low-res pixel index        [K2]
  float* sq = ds + 2 * K2;                                 // (rotated) query            [dq]
  float* sg = sq + dq;                                     // upstream gradient slice    [dv]

  const bool rope = p.cos_y != nullptr;
  const int half = dq / 2, P = dq / 4;

  for (int64_t item = int64_t(blockIdx.x) * kBwdWarps + warp; item < total_items;
       item += int64_t(gridDim.x) * kBwdWarps) {
    const int head = int(item % p.heads);
    int64_t pix = item / p.heads;
    const int x = int(pix % p.Wo);
    pix /= p.Wo;
    const int y = int(pix % p.Ho);
    const int b = int(pix / p.Ho);

    const float* qp = p.q + int64_t(b) * p.q_stride_b + int64_t(y / p.rep_y) * p.q_stride_y +
                      int64_t(x / p.rep_x) * p.q_stride_x + head * dq;
    if (rope) {
      for (int i = lane; i < half; i += 32) {
        const float a = qp[i], bb = qp[i + half];
        const bool on_y = i < P;
        const int ti = on_y ? i : i - P;
        const float c = on_y ? p.cos_y[int64_t(y) * P + ti] : p.cos_x[int64_t(x) * P + ti];
        const float s = on_y ? p.sin_y[int64_t(y) * P + ti] : p.sin_x[int64_t(x) * P + ti];
        sq[i] = a * c - bb * s;
        sq[i + half] = bb * c + a * s;
      }
    } else {
      for (int i = lane; i < dq; i += 32) sq[i] = qp[i];
    }
    const int64_t pixoff = (int64_t(b) * p.Ho + y) * p.Wo + x;
    const float* gp = p.dout + pixoff * p.C + head * dv;
    for (int c = lane; c < dv; c += 32) sg[c] = gp[c];
    for (int tap = lane; tap < K2; tap += 32) {
      const int t = tap / Kw, u = tap - t * Kw;
      const int r = tap_index(p.row_tap, y, t, Kh, rh, p.h);
      const int c = tap_index(p.col_tap, x, u, Kw, rw, p.w);
      idx[tap] = (b * p.h + r) * p.w + c;
    }
    __syncwarp();

    // ---- recompute the probabilities; dP; lane <-> tap
    float m = -INFINITY;
    for (int tap = lane; tap < K2; tap += 32) {
      const float* kp = p.k + int64_t(idx[tap]) * p.D + head * dq;
      float acc = 0.f;
      for (int d = 0; d < dq; ++d) acc = fmaf(sq[d], __ldg(kp + d), acc);
      acc *= p.scale;
      sc[tap] = acc;
      m = fmaxf(m, acc);
    }
    m = warp_max(m);
    float l = 0.f;
    for (int tap = lane; tap < K2; tap += 32) {
      const float e = expf(sc[tap] - m);
      sc[tap] = e;
      l += e;
    }
    l = warp_sum(l);
    const float inv = 1.f / l;
    float delta = 0.f;
    for (int tap = lane; tap < K2; tap += 32) {
      const float* vp = p.v + int64_t(idx[tap]) * p.C + head * dv;
      float acc = 0.f;
      for (int c = 0; c < dv; ++c) acc = fmaf(sg[c], __ldg(vp + c), acc);
      const float pr = sc[tap] * inv;
      sc[tap] = pr;
      ds[tap] = acc;            // dP for now
      delta = fmaf(pr, acc, delta);
    }
    delta = warp_sum(delta);
    for (int tap = lane; tap < K2; tap += 32) ds[tap] = sc[tap] * (ds[tap] - delta) * p.scale;
    __syncwarp();

    // ---- dq: lane <-> channel
    float* dqp = p.dq + pixoff * p.D + head * dq;
    for (int d = lane; d < dq; d += 32) {
      float acc = 0.f;
      for (int tap = 0; tap < K2; ++tap) acc = fmaf(ds[tap], __ldg(p.k + int64_t(idx[tap]) * p.D + head * dq + d), acc);
      dqp[d] = acc;
    }
    // ---- dk, dv: scatter-add (duplicated taps of non-integer ratios simply add twice)
    for (int tap = 0; tap < K2; ++tap) {
      const float w = ds[tap];
      float* dkp = p.dk + int64_t(idx[tap]) * p.D + head * dq;
      for (int d = lane; d < dq; d += 32) red_add(dkp + d, w * sq[d]);
      const float pr = sc[tap];
      float* dvp = p.dv + int64_t(idx[tap]) * p.C + head * dv;
      for (int c = lane; c < dv; c += 32) red_add(dvp + c, pr * sg[c]);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// Cell kernel.  CTA = (batch, cell, head), NT threads.  Shared memory: the K window [K2][DQ] and the V
// window [K2][dv] of the cell (fp32), and per pixel batch of PB pixels: rotated queries [PB][DQ],
// upstream gradients [PB][dv], probabilities [PB][K2], scaled dS [PB][K2].  Phases per pixel batch:
//   A  thread <-> (pixel, tap):  s = <q, k_t>, then per pixel softmax                (K2*DQ MACs / pixel)
//   B  thread <-> (pixel, tap):  dP = <g, v_t>; delta; dS                            (K2*dv)
//   C  thread <-> (pixel, d)  :  dq = sum_t dS_t k_t[d]  -> global                   (K2*DQ)
//   D  thread <-> fixed (tap, channel) slots for the WHOLE cell, accumulators in registers:
//        dKwin[t][d] += sum_pixels dS[p][t] q[p][d],  dVwin[t][c] += sum_pixels P[p][t] g[p][c]
// and after the last batch the register accumulators leave with one atomic each.
template <int DQ, int PB, int NT, int MAXK, int MAXV>
__global__ void __launch_bounds__(NT)
xattn_bwd_cell_kernel(naf_xattn_bwd_params p, int rh, int rw, int dv, int n_acc_k, int n_acc_v) {
  extern __shared__ __align__(16) float sm[];
  const int K = p.K, K2 = K * K;
  const int tid = threadIdx.x;
  int item = blockIdx.x;
  const int head = item % p.heads;
  item /= p.heads;
  const int cj = item % p.w;
  item /= p.w;
  const int ci = item % p.h;
  const int b = item / p.h;
  const int wy0 = window_origin(ci, p.h, K), wx0 = window_origin(cj, p.w, K);
  const int K2P = K2 + 1;                       // padded rows: consecutive threads read consecutive rows
  constexpr int DQS = DQ + 4;                   // (thread <-> tap), so row strides must not be multiples
  const int DVS = dv + 1;                       // of the 32 banks

  float* sK = sm;                               // [K2][DQS]
  float* sV = sK + K2 * DQS;                    // [K2][DVS]
  float* sQ = sm + (K2 * DQS + K2 * DVS + 3) / 4 * 4;   // [PB][DQ], 16-byte aligned (float4 reads in phase A)
  float* sG = sQ + PB * DQ;                     // [PB][dv]
  float* sP = sG + PB * dv;                     // [PB][K2P]
  float* sD = sP + PB * K2P;                    // [PB][K2P]

  for (int i = tid; i < K2 * (DQ / 4); i += NT) {
    const int t = i / (DQ / 4), d4 = i - t * (DQ / 4);
    const int ty = t / K, tx = t - ty * K;
    const float* src = p.k + (int64_t(b * p.h + wy0 + ty) * p.w + wx0 + tx) * p.D + head * DQ;
    reinterpret_cast<float4*>(sK + t * DQS)[d4] = reinterpret_cast<const float4*>(src)[d4];
  }
  for (int i = tid; i < K2 * dv; i += NT) {
    const int t = i / dv, c = i - t * dv;
    const int ty = t / K, tx = t - ty * K;
    sV[t * DVS + c] = p.v[(int64_t(b * p.h + wy0 + ty) * p.w + wx0 + tx) * p.C + head * dv + c];
  }

  // register accumulators: slot j of this thread is element (tid + j*NT) of dKwin (K2*DQ) / dVwin (K2*dv)
  float acck[MAXK], accv[MAXV];
#pragma unroll
  for (int j = 0; j < MAXK; ++j) acck[j] = 0.f;
#pragma unroll
  for (int j = 0; j < MAXV; ++j) accv[j] = 0.f;

  const bool rope = p.cos_y != nullptr;
  constexpr int HALF = DQ / 2, PQ = DQ / 4;
  const int npix = rh * rw;
  for (int p0 = 0; p0 < npix; p0 += PB) {
    const int nb = min(PB, npix - p0);
    __syncthreads();   // previous batch fully consumed (also orders the window loads before phase A)
    // ---- load q (rotated) and g of the batch
    for (int i = tid; i < PB * HALF; i += NT) {
      const int pp = i / HALF, c = i - pp * HALF;
      float a = 0.f, bb = 0.f;
      if (pp < nb) {
        const int pi = p0 + pp, py = pi / rw;
        const int y = ci * rh + py, x = cj * rw + (pi - py * rw);
        const float* qp = p.q + int64_t(b) * p.q_stride_b + int64_t(y / p.rep_y) * p.q_stride_y +
                          int64_t(x / p.rep_x) * p.q_stride_x + head * DQ;
        a = qp[c];
        bb = qp[c + HALF];
        if (rope) {
          const bool on_y = c < PQ;
          const int ti = on_y ? c : c - PQ;
          const float cs = on_y ? p.cos_y[int64_t(y) * PQ + ti] : p.cos_x[int64_t(x) * PQ + ti];
          const float sn = on_y ? p.sin_y[int64_t(y) * PQ + ti] : p.sin_x[int64_t(x) * PQ + ti];
          const float ra = a * cs - bb * sn, rb = bb * cs + a * sn;
          a = ra;
          bb = rb;
        }
      }
      sQ[pp * DQ + c] = a;
      sQ[pp * DQ + c + HALF] = bb;
    }
    for (int i = tid; i < PB * dv; i += NT) {
      const int pp = i / dv, c = i - pp * dv;
      float g = 0.f;
      if (pp < nb) {
        const int pi = p0 + pp, py = pi / rw;
        const int y = ci * rh + py, x = cj * rw + (pi - py * rw);
        g = p.dout[((int64_t(b) * p.Ho + y) * p.Wo + x) * p.C + head * dv + c];
      }
      sG[i] = g;
    }
    __syncthreads();
    // ---- phase A: scores
    for (int i = tid; i < PB * K2; i += NT) {
      const int pp = i / K2, t = i - pp * K2;
      const float4* q4 = reinterpret_cast<const float4*>(sQ + pp * DQ);
      const float4* k4 = reinterpret_cast<const float4*>(sK + t * DQS);
      float acc = 0.f;
#pragma unroll 4
      for (int d = 0; d < DQ / 4; ++d) {
        const float4 a = q4[d], kk = k4[d];
        acc = fmaf(a.x, kk.x, acc);
        acc = fmaf(a.y, kk.y, acc);
        acc = fmaf(a.z, kk.z, acc);
        acc = fmaf(a.w, kk.w, acc);
      }
      sP[pp * K2P + t] = acc * p.scale;
    }
    __syncthreads();
    // softmax per pixel: one warp per pixel
    for (int pp = tid >> 5; pp < PB; pp += NT / 32) {
      const int lane = tid & 31;
      float m = -INFINITY;
      for (int t = lane; t < K2; t += 32) m = fmaxf(m, sP[pp * K2P + t]);
      m = warp_max(m);
      float l = 0.f;
      for (int t = lane; t < K2; t += 32) {
        const float e = expf(sP[pp * K2P + t] - m);
        sP[pp * K2P + t] = e;
        l += e;
      }
      l = warp_sum(l);
      const float inv = 1.f / l;
      for (int t = lane; t < K2; t += 32) sP[pp * K2P + t] *= inv;
    }
    __syncthreads();
    // ---- phase B: dP, then dS per pixel
    for (int i = tid; i < PB * K2; i += NT) {
      const int pp = i / K2, t = i - pp * K2;
      const float* g = sG + pp * dv;
      const float* vv = sV + t * DVS;
      float acc = 0.f;
      for (int c = 0; c < dv; ++c) acc = fmaf(g[c], vv[c], acc);
      sD[pp * K2P + t] = acc;
    }
    __syncthreads();
    for (int pp = tid >> 5; pp < PB; pp += NT / 32) {
      const int lane = tid & 31;
      float delta = 0.f;
      for (int t = lane; t < K2; t += 32) delta = fmaf(sP[pp * K2P + t], sD[pp * K2P + t], delta);
      delta = warp_sum(delta);
      for (int t = lane; t < K2; t += 32)
        sD[pp * K2P + t] = sP[pp * K2P + t] * (sD[pp * K2P + t] - delta) * p.scale;
    }
    __syncthreads();
    // ---- phase C: dq -> global (gradient w.r.t. the ROTATED query)
    for (int i = tid; i < nb * DQ; i += NT) {
      const int pp = i / DQ, d = i - pp * DQ;
      float acc = 0.f;
      for (int t = 0; t < K2; ++t) acc = fmaf(sD[pp * K2P + t], sK[t * DQS + d], acc);
      const int pi = p0 + pp, py = pi / rw;
      const int y = ci * rh + py, x = cj * rw + (pi - py * rw);
      p.dq[((int64_t(b) * p.Ho + y) * p.Wo + x) * p.D + head * DQ + d] = acc;
    }
    // ---- phase D: register accumulation of the window gradients (padding pixels hold q = g = 0)
#pragma unroll
    for (int j = 0; j < MAXK; ++j) {
      if (j < n_acc_k) {
        const int e = tid + j * NT;
        if (e < K2 * DQ) {
          const int t = e / DQ, d = e - t * DQ;
          float acc = acck[j];
#pragma unroll 4
          for (int pp = 0; pp < PB; ++pp) acc = fmaf(sD[pp * K2P + t], sQ[pp * DQ + d], acc);
          acck[j] = acc;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
      if (j < n_acc_v) {
        const int e = tid + j * NT;
        if (e < K2 * dv) {
          const int t = e / dv, c = e - t * dv;
          float acc = accv[j];
#pragma unroll 4
          for (int pp = 0; pp < PB; ++pp) acc = fmaf(sP[pp * K2P + t], sG[pp * dv + c], acc);
          accv[j] = acc;
        }
      }
    }
  }
  // ---- flush: one atomic per window element
#pragma unroll
  for (int j = 0; j < MAXK; ++j) {
    if (j < n_acc_k) {
      const int e = tid + j * NT;
      if (e < K2 * DQ) {
        const int t = e / DQ, d = e - t * DQ;
        const int ty = t / K, tx = t - ty * K;
        red_add(p.dk + (int64_t(b * p.h + wy0 + ty) * p.w + wx0 + tx) * p.D + head * DQ + d, acck[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    if (j < n_acc_v) {
      const int e = tid + j * NT;
      if (e < K2 * dv) {
        const int t = e / dv, c = e - t * dv;
        const int ty = t / K, tx = t - ty * K;
        red_add(p.dv + (int64_t(b * p.h + wy0 + ty) * p.w + wx0 + tx) * p.C + head * dv + c, accv[j]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// dx = R^T (dq + pooled-key gradient).  One thread per rotation pair (a_i, b_i) of one pixel.
__global__ void __launch_bounds__(256)
rope_kpool_bwd_kernel(naf_kpool_bwd_params p, int64_t total_pairs) {
  const bool rope = p.cos_y != nullptr;
  const int d_head = rope ? p.D / p.rope_heads : p.D;
  const int half = d_head / 2, P = d_head / 4;
  const int pairs = p.D / 2;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total_pairs; i += int64_t(gridDim.x) * blockDim.x) {
    const int pr = int(i % pairs);
    int64_t pix = i / pairs;
    const int x = int(pix % p.Wo);
    pix /= p.Wo;
    const int y = int(pix % p.Ho);
    const int b = int(pix / p.Ho);
    const int head = pr / half, i0 = pr - head * half;
    const int ca = head * d_head + i0;
    const int64_t off = ((int64_t(b) * p.Ho + y) * p.Wo + x) * p.D;
    float ga = p.dq ? p.dq[off + ca] : 0.f, gb = p.dq ? p.dq[off + ca + half] : 0.f;
    if (p.dk) {
      // ATen adaptive pooling bins holding this pixel: [floor(i*Ho/h), ceil((i+1)*Ho/h)); 1 or 2 per axis
      int i_lo = int((int64_t(y) * p.h) / p.Ho), j_lo = int((int64_t(x) * p.w) / p.Wo);
      for (int bi = max(0, i_lo - 1); bi <= min(p.h - 1, i_lo + 1); ++bi) {
        const int ys = int((int64_t(bi) * p.Ho) / p.h), ye = int((int64_t(bi + 1) * p.Ho + p.h - 1) / p.h);
        if (y < ys || y >= ye) continue;
        for (int bj = max(0, j_lo - 1); bj <= min(p.w - 1, j_lo + 1); ++bj) {
          const int xs = int((int64_t(bj) * p.Wo) / p.w), xe = int((int64_t(bj + 1) * p.Wo + p.w - 1) / p.w);
          if (x < xs || x >= xe) continue;
          const float inv = 1.f / float((ye - ys) * (xe - xs));
          const float* dkp = p.dk + ((int64_t(b) * p.h + bi) * p.w + bj) * p.D;
          ga = fmaf(dkp[ca], inv, ga);
          gb = fmaf(dkp[ca + half], inv, gb);
        }
      }
    }
    if (rope) {
      const bool on_y = i0 < P;
      const int ti = on_y ? i0 : i0 - P;
      const float c = on_y ? p.cos_y[int64_t(y) * P + ti] : p.cos_x[int64_t(x) * P + ti];
      const float s = on_y ? p.sin_y[int64_t(y) * P + ti] : p.sin_x[int64_t(x) * P + ti];
      // forward: a' = a c - b s, b' = b c + a s   =>   da = ga c + gb s, db = gb c - ga s
      const float da = ga * c + gb * s, db = gb * c - ga * s;
      ga = da;
      gb = db;
    }
    p.dx[off + ca] = ga;
    p.dx[off + ca + half] = gb;
  }
}

// ------------------------------------------------------------------------------------ host side
namespace {

constexpr int kCellPB = 16;

size_t bwd_cell_smem(int K2, int dq, int dv) {
  const size_t win = size_t(K2) * (dq + 4) + size_t(K2) * (dv + 1);
  return sizeof(float) * ((win + 3) / 4 * 4 + 4 + kCellPB * dq + kCellPB * dv + 2 * kCellPB * (K2 + 1));
}

// accumulator slots per thread: (256 threads, 16 + 40) for windows up to 9x9 / dv 192-ish,
// (512 threads, 16 + 64) for 11x11 with dv up to 256
template <int DQ, int NT, int MAXK, int MAXV>
int launch_bwd_cell_cfg(const naf_xattn_bwd_params& p, cudaStream_t st) {
  const int K2 = p.K * p.K, dv = p.C / p.heads;
  const int n_acc_k = (K2 * DQ + NT - 1) / NT, n_acc_v = (K2 * dv + NT - 1) / NT;
  const size_t smem = bwd_cell_smem(K2, DQ, dv);
  auto kern = xattn_bwd_cell_kernel<DQ, kCellPB, NT, MAXK, MAXV>;
  cudaError_t e = ensure_dyn_smem(kern, int(smem));
  if (e != cudaSuccess) return fail(NAF_ERR_CUDA, "xattn_bwd(cell): smem opt-in failed: %s", cudaGetErrorString(e));
  const unsigned grid = unsigned(p.B) * p.h * p.w * p.heads;
  kern<<<grid, NT, smem, st>>>(p, p.Ho / p.h, p.Wo / p.w, dv, n_acc_k, n_acc_v);
  return check_launch("xattn_bwd_cell");
}

// 0 = unsupported, 256 / 512 = thread count of the configuration that fits
int bwd_cell_config(const naf_xattn_bwd_params& p) {
  if (p.row_tap || p.col_tap || p.Ho % p.h || p.Wo % p.w) return 0;
  const int dq = p.D / p.heads, dv = p.C / p.heads, K2 = p.K * p.K;
  if (dq != 64 && dq != 32) return 0;
  if ((p.Ho / p.h) * (p.Wo / p.w) < 16) return 0;                     // tiny cells: per-pixel atomics are fine
  if (bwd_cell_smem(K2, dq, dv) > 200 * 1024) return 0;
  if (!aligned16(p.k) || p.D % 4) return 0;
  if (int64_t(p.B) * p.h * p.w * p.heads >= (int64_t(1) << 31)) return 0;
  if ((K2 * dq + 255) / 256 <= 16 && (K2 * dv + 255) / 256 <= 40) return 256;
  if ((K2 * dq + 511) / 512 <= 16 && (K2 * dv + 511) / 512 <= 64) return 512;
  return 0;
}

template <int DQ>
int launch_bwd_cell(const naf_xattn_bwd_params& p, int nt, cudaStream_t st) {
  return nt == 256 ? launch_bwd_cell_cfg<DQ, 256, 16, 40>(p, st) : launch_bwd_cell_cfg<DQ, 512, 16, 64>(p, st);
}

}  // namespace

int launch_xattn_bwd(const naf_xattn_bwd_params& p, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(p.dk, 0, sizeof(float) * size_t(p.B) * p.h * p.w * p.D, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(p.dv, 0, sizeof(float) * size_t(p.B) * p.h * p.w * p.C, st);
  if (e != cudaSuccess) return fail(NAF_ERR_CUDA, "xattn_bwd: memset failed: %s", cudaGetErrorString(e));
  const int dq = p.D / p.heads, dv = p.C / p.heads, K2 = p.K * (p.Kw ? p.Kw : p.K);
  const bool rect = p.Kw != 0 && p.Kw != p.K;   // rectangular windows: the generic kernel
  if (rect && p.algo != NAF_ALGO_AUTO && p.algo != NAF_ALGO_GENERIC)
    return fail(NAF_ERR_UNSUPPORTED, "xattn_bwd: rectangular windows (%dx%d) run on the generic kernel only", p.K, p.Kw);
  if (!rect) {
    // integer ratios, 64-wide heads, cells of >= 64 pixels: the five GEMMs on the tensor core (naf_xattn_bwd_tc.cu)
    const char* why = "";
    const bool tc_ok = xattn_bwd_cell_tc_supported(p, &why);
    if (tc_ok && (p.algo == NAF_ALGO_AUTO || p.algo == NAF_ALGO_CELL_TC)) return launch_xattn_bwd_cell_tc(p, st);
    if (p.algo == NAF_ALGO_CELL_TC)
      return fail(NAF_ERR_UNSUPPORTED, "xattn_bwd: the tensor-core cell kernel does not support this request: %s", why);
  }
  const int cell_nt = (p.algo != NAF_ALGO_GENERIC && !rect) ? bwd_cell_config(p) : 0;
  if (cell_nt) return dq == 64 ? launch_bwd_cell<64>(p, cell_nt, st) : launch_bwd_cell<32>(p, cell_nt, st);
  if (p.algo == NAF_ALGO_CELL_SIMT)
    return fail(NAF_ERR_UNSUPPORTED, "xattn_bwd: the cell kernel does not support this request");
  const size_t smem = size_t(kBwdWarps) * (3 * K2 + dq + dv) * sizeof(float);
  NAF_REQUIRE(smem <= 200 * 1024, NAF_ERR_UNSUPPORTED, "xattn_bwd(generic): kernel_size %d / head dims %d, %d need %zu B shared memory",
              p.K, dq, dv, smem);
  NAF_REQUIRE(int64_t(p.B) * p.h * p.w < (int64_t(1) << 31), NAF_ERR_UNSUPPORTED, "xattn_bwd(generic): feature map too large");
  if (smem > 48 * 1024) {
    e = ensure_dyn_smem(xattn_bwd_generic_kernel, int(smem));
    if (e != cudaSuccess) return fail(NAF_ERR_CUDA, "xattn_bwd(generic): smem opt-in failed: %s", cudaGetErrorString(e));
  }
  const int64_t items = int64_t(p.B) * p.Ho * p.Wo * p.heads;
  int64_t blocks = (items + kBwdWarps - 1) / kBwdWarps;
  const int64_t cap = int64_t(device_sm_count()) * 16 * 8;
  if (blocks > cap) blocks = cap;
  xattn_bwd_generic_kernel<<<unsigned(blocks), kBwdWarps * 32, smem, st>>>(p, p.Ho / p.h, p.Wo / p.w, items);
  return check_launch("xattn_bwd_generic");
}

int launch_rope_kpool_bwd(const naf_kpool_bwd_params& p, cudaStream_t st) {
  const int64_t total = int64_t(p.B) * p.Ho * p.Wo * (p.D / 2);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = int64_t(device_sm_count()) * 32;
  if (blocks > cap) blocks = cap;
  prefer_max_shared(rope_kpool_bwd_kernel);
  rope_kpool_bwd_kernel<<<unsigned(blocks), 256, 0, st>>>(p, total);
  return check_launch("rope_kpool_bwd");
}

}  // namespace naf
