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
//
// Roles (384 threads): 8 worker warps, two threads per pixel row (global loads, RoPE, split-fp16 staging of Q / G, the
// two softmax phases, dS, the dQ store, the window-gradient atomics); one MMA warp group whose 128 threads stage the
// K / V windows and whose ONE elected thread issues every MMA.  Twelve mbarriers carry the hand-offs (BwdBars); the
// upstream-gradient chunk is double buffered and, where shared memory and TMEM allow, a second Q image / S region lets
// the next tile's S run under dQ / dKw of the current one.  setmaxnreg moves registers from the MMA group to the workers.
#include "naf_common.cuh"
#include "naf_umma.cuh"

#ifndef NAF_BWD_EXP
#define NAF_BWD_EXP 0   // profiling experiments only (scripts/gpu_bwd_exp.sh); 0 in every real build
#endif

namespace naf {

using namespace umma;

#if NAF_BWD_EXP & 32
// profiling build only: clock stamps of thread 0 of four CTAs at every phase boundary (scripts/trace_bwd.py)
__device__ long long g_bwd_trace[4][128];
#define NAF_TR()                                                        \
  do {                                                                  \
    if (tid == 0 && tr_slot >= 0 && tr_i < 128) g_bwd_trace[tr_slot][tr_i++] = clock64(); \
  } while (0)
#else
#define NAF_TR() do { } while (0)
#endif

namespace {

constexpr int DQ = 64;        // head dim of q / k
constexpr int KC = DQ / 8;    // 16-byte chunks along a 64-channel operand row
constexpr int NW = 256;       // worker threads: two per pixel row (TMEM lane)
constexpr int NT = NW + 128;  // + the MMA warp group (one issuing thread; a whole warp group so that setmaxnreg applies)
constexpr int DVC = 64;       // value channels per chunk
constexpr int PIXIMG = KC * 128 * 16;   // one fp16 image of a 128 x 64 pixel-major operand (16 KB)

template <int TP>
struct BwdCfg {
  static constexpr int kWin = KC * TP * 16;   // one fp16 image of a TP x 64 window operand
  // P^T / dS^T image: TP taps x 128 pixels fp16, [pixel group of 8][tap group of 8][pixel % 8][8 taps]; kPG = bytes
  // per pixel group.  The M = 128 MMAs that read it as their A operand run over tap groups TP/8 .. 15 into whatever
  // follows (the next pixel group, and past the last one the K window, which is why T comes first in shared
  // memory): finite fp16 bits that only reach accumulator lanes >= TP, which nothing ever reads.
  static constexpr int kPG = TP * 16;
  static constexpr int kT = 16 * kPG;
  static constexpr int kTA = (TP + 31) / 32 * 32;                   // TMEM regions start on 32-column boundaries
};

__device__ __forceinline__ void red_add4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// n / d for 0 <= n, n * d < 2^32, with magic = ceil(2^32 / d) computed on the host (d = 1: no magic number fits)
__host__ __device__ __forceinline__ uint32_t magic_of(int d) { return d > 1 ? uint32_t(((1ull << 32) + uint32_t(d) - 1) / uint32_t(d)) : 0u; }
__device__ __forceinline__ int magic_div(int n, int d, uint32_t magic) { return d > 1 ? int(__umulhi(uint32_t(n), magic)) : n; }

// One lane of a converged warp.  Under elect.sync the compiler knows that exactly one thread runs the guarded region and
// keeps the MMA descriptors in uniform registers, where UTCHMMA reads them: back-to-back MMAs in the SASS.  Under
// `lane == 0` every MMA is preceded by R2UR moves and an election loop on one thread's dependent chain (measured ~85
// clocks per N = 64 MMA, more than the tensor core needs).
template <bool kElect>
__device__ __forceinline__ bool elect_one() {
  if constexpr (kElect) {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
  } else {
    return (threadIdx.x & 31) == 0;
  }
}

// 32-byte global load that stays where it is written (register prefetch: the compiler must not sink it to its use)
__device__ __forceinline__ void ldg8_keep(const float* p, float (&r)[8]) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
               : "l"(p));
}

__device__ __forceinline__ void split8(const float (&x)[8], uint4& hi, uint4& lo) {
  split2_f16(x[0], x[1], hi.x, lo.x);
  split2_f16(x[2], x[3], hi.y, lo.y);
  split2_f16(x[4], x[5], hi.z, lo.z);
  split2_f16(x[6], x[7], hi.w, lo.w);
}

}  // namespace

// Barriers of the hand-off protocol (one CTA): "ready" barriers collect one arrival per worker thread, "done"
// barriers are tcgen05.commit targets.  Every waiter waits every phase of its barriers at least once and never
// skips one; a wait may be repeated as long as nothing has committed / arrived on the barrier in between.
struct BwdBars {
  uint64_t q_ready[2];    // Q of a tile staged in Q buffer i            (workers -> MMA thread)
  uint64_t s_done[2];     // S of a tile complete in S region i          (tensor core -> workers)
  uint64_t pt_ready;      // P^T of the tile written                     (workers -> MMA thread)
  uint64_t g_ready[2];    // upstream-gradient chunk staged in G buffer  (workers -> MMA thread)
  uint64_t g_done[2];     // dP / dVw MMAs of a chunk step complete      (tensor core -> workers)
  uint64_t ds_ready;      // dS in TMEM, dS^T in shared memory           (workers -> MMA thread)
  uint64_t dq_done;       // dQ / dKw of the tile complete               (tensor core -> workers)
  uint64_t v_ready;       // resident V window staged                    (warps 9-11 -> MMA thread)
};

__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NW) : "memory"); }

template <int TP>
__global__ void __launch_bounds__(NT, 1)
xattn_bwd_cell_tc_kernel(naf_xattn_bwd_params p, int rh, int rw, int dv, int v_resident, int gbufs, int overlap,
                         uint32_t m_rw, uint32_t m_ry, uint32_t m_rx) {
  using Cfg = BwdCfg<TP>;
  constexpr int SC = TP / 2;   // S / dP columns per thread
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ BwdBars bars;
  __shared__ uint32_t tmem_base_s;
  __shared__ float red_a[2][128];
  __shared__ float red_b[2][128];

  uint8_t* sThi = smem;               // P^T, then dS^T: MN-major A images
  uint8_t* sTlo = sThi + Cfg::kT;
  uint8_t* sKhi = sTlo + Cfg::kT;
  uint8_t* sKlo = sKhi + Cfg::kWin;
  // V window: [channel group][tap][16 B]; the whole value head when it fits (staged once per cell), else one
  // 64-channel chunk restaged per tile and chunk
  const int v_img = (v_resident ? (dv >> 3) : KC) * TP * 16;
  uint8_t* sVhi = sKlo + Cfg::kWin;
  uint8_t* sVlo = sVhi + v_img;
  uint8_t* sQ = sVlo + v_img;                              // (1 + overlap) x (hi, lo) images of the rotated queries
  uint8_t* sG = sQ + (1 + overlap) * (2 * PIXIMG);         // gbufs x (hi, lo) images of the upstream-gradient chunk

  const int tid = threadIdx.x, lane = tid & 31;
  // the warp index through a shuffle: provably warp-uniform, so that the role branches below are uniform branches
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int rowgrp = warp & 3, hf = (warp >> 2) & 1;
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

#if NAF_BWD_EXP & 32
  const int tr_slot = blockIdx.x == gridDim.x / 4 ? 0 : blockIdx.x == gridDim.x / 2 ? 1 : blockIdx.x == gridDim.x / 4 * 3 ? 2 : blockIdx.x == gridDim.x - 200 ? 3 : -1;
  int tr_i = 0;
#endif
  NAF_TR();   // 0: CTA start
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    mbar_init(&bars.q_ready[0], NW);
    mbar_init(&bars.q_ready[1], NW);
    mbar_init(&bars.s_done[0], 1);
    mbar_init(&bars.s_done[1], 1);
    mbar_init(&bars.pt_ready, NW);
    mbar_init(&bars.g_ready[0], NW);
    mbar_init(&bars.g_ready[1], NW);
    mbar_init(&bars.g_done[0], 1);
    mbar_init(&bars.g_done[1], 1);
    mbar_init(&bars.ds_ready, NW);
    mbar_init(&bars.dq_done, 1);
    mbar_init(&bars.v_ready, NT - NW - 32);
    fence_mbar_init();
  }
  const bool rope = p.cos_y != nullptr;
  const float qscale = p.scale * 1.4426950408889634f;
  const int npix = rh * rw;
  const int ntiles = (npix + 127) >> 7;
  const int nchunks = (dv + DVC - 1) / DVC;
  const int y0 = ci * rh, x0 = cj * rw;
  const float* vwin = p.v + (int64_t(b * p.h + wy0) * p.w + wx0) * p.C + head * dv;
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
    const int py = r.valid ? magic_div(pi, rw, m_rw) : 0;
    r.y = y0 + py;
    r.x = x0 + (r.valid ? pi - py * rw : 0);
    r.pix = (int64_t(b) * p.Ho + r.y) * p.Wo + r.x;
    return r;
  };
  // Global loads are laid out for the memory pipeline, not for the TMEM lanes: one warp instruction reads 8 pixels x
  // 128 contiguous bytes (lane = pixel % 8 + 8 * (32-byte piece)), i.e. 8 full lines, where a row-per-thread mapping
  // touches 32 lines for the same bytes (measured: the load / store unit's request rate bounded the whole kernel).
  // The operand images in shared memory are indexed by (channel group, pixel), so any thread may stage any element,
  // and 8 consecutive lanes write 8 consecutive pixels of one group: conflict-free.
  const int lp = lane & 7, lg = lane >> 3;
  auto pix_of = [&](int tile, int P, int& y, int& x) -> bool {     // tile-local pixel P -> target coordinates
    const int pi = tile * 128 + P;
    const bool valid = tile < ntiles && pi < npix;
    const int py = valid ? magic_div(pi, rw, m_rw) : 0;
    y = y0 + py;
    x = x0 + (valid ? pi - py * rw : 0);
    return valid;
  };
  float qa8[2][8], qb8[2][8];     // pixel (k * 8 + warp) * 8 + lp: channels lg * 8 .. + 7 and their rotation partners (+ 32)
  auto load_q = [&](int tile) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      int y, x;
      if (pix_of(tile, (k * 8 + warp) * 8 + lp, y, x)) {
        const float* qp = p.q + int64_t(b) * p.q_stride_b + int64_t(magic_div(y, p.rep_y, m_ry)) * p.q_stride_y +
                          int64_t(magic_div(x, p.rep_x, m_rx)) * p.q_stride_x + head * DQ + lg * 8;
        ldg8_keep(qp, qa8[k]);
        ldg8_keep(qp + 32, qb8[k]);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) qa8[k][e] = qb8[k][e] = 0.f;
      }
    }
  };
  // Upstream-gradient chunks are loaded TWO chunk steps ahead when the registers allow (up to 64 taps): a step is
  // ~1.5 us, about one HBM round trip under load, and with one step of lead every step waited for its loads.
  constexpr bool kDeepG = TP <= 64;
  float g8r[kDeepG ? 2 : 1][4][8];     // [set][block k * 8 + warp of the chunk: 8 pixels x one quad of channel groups]
  auto load_g = [&](int tile, int c, float (&g8r)[4][8]) {
    const int ng = min(DVC, dv - c * DVC) >> 3;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int bi = k * 8 + warp, g = (bi >> 4) * 4 + lg;
      int y, x;
      const bool valid = pix_of(tile, (bi & 15) * 8 + lp, y, x);
      if (valid && g < ng && !(NAF_BWD_EXP & 16)) {
        ldg8_keep(p.dout + ((int64_t(b) * p.Ho + y) * p.Wo + x) * p.C + head * dv + c * DVC + g * 8, g8r[k]);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) g8r[k][e] = 0.f;
      }
    }
  };
  // rotated, unscaled queries of a tile: K-major [chunk][row][16 B] hi / lo images
  auto stage_q = [&](int tile, uint8_t* sQhi) {
    uint8_t* const sQlo = sQhi + PIXIMG;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int P = (k * 8 + warp) * 8 + lp;
      int y, x;
      if (pix_of(tile, P, y, x) && rope) {
        // rotation pair j = lg * 8 + e: row angles for j < 16, column angles above
        const float* ct = lg < 2 ? p.cos_y + int64_t(y) * 16 + lg * 8 : p.cos_x + int64_t(x) * 16 + (lg - 2) * 8;
        const float* st = lg < 2 ? p.sin_y + int64_t(y) * 16 + lg * 8 : p.sin_x + int64_t(x) * 16 + (lg - 2) * 8;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const float4 c4 = __ldg(reinterpret_cast<const float4*>(ct) + j), s4 = __ldg(reinterpret_cast<const float4*>(st) + j);
          const float c[4] = {c4.x, c4.y, c4.z, c4.w}, sn[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float a = qa8[k][4 * j + e], bb = qb8[k][4 * j + e];
            qa8[k][4 * j + e] = a * c[e] - bb * sn[e];
            qb8[k][4 * j + e] = bb * c[e] + a * sn[e];
          }
        }
      }
      uint4 hi, lo;
      split8(qa8[k], hi, lo);
      *reinterpret_cast<uint4*>(sQhi + (lg * 128 + P) * 16) = hi;
      *reinterpret_cast<uint4*>(sQlo + (lg * 128 + P) * 16) = lo;
      split8(qb8[k], hi, lo);
      *reinterpret_cast<uint4*>(sQhi + ((lg + 4) * 128 + P) * 16) = hi;
      *reinterpret_cast<uint4*>(sQlo + ((lg + 4) * 128 + P) * 16) = lo;
    }
  };

  Px cur = px_of(0);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  NAF_TR();   // 1: prologue done

  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_off = uint32_t(rowgrp * 32) << 16;
  // TMEM: S (later dS, fp16 hi | lo) x (1 + overlap) | dP (later dQ) | dKw (taps x 64) | dVw (taps x dv)
  const uint32_t tS0 = tmem;
  const uint32_t tdP = tmem + (1 + overlap) * Cfg::kTA;
  const uint32_t tdK = tdP + (Cfg::kTA > DQ ? Cfg::kTA : DQ);
  const uint32_t tdV = tdK + DQ;
  // buffer / barrier index and wait parity of tile t (Q images, S regions) and of chunk step s (G images)
  auto tb = [&](int t) { return overlap ? (t & 1) : 0; };
  auto tpar = [&](int t) { return uint32_t(overlap ? (t >> 1) & 1 : t & 1); };
  auto gb = [&](int s) { return gbufs == 2 ? (s & 1) : 0; };
  auto gpar = [&](int s) { return uint32_t(gbufs == 2 ? (s >> 1) & 1 : s & 1); };

  if (warp >= NW / 32) {
    // ================================================================ MMA warp group
    // 12 warps start with 168 registers each; this group hands its share to the workers
    asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
    // The group stages the cell's windows while the workers load and stage the first queries and upstream gradients;
    // only the issuing thread depends on them (the workers never read K or a resident V).  All 128 threads stage K;
    // then the issuing warp starts S of the first tile while the other three warps stage the (larger) V window, which
    // the first dP needs a whole softmax later.
    {
      // K window: K-major [channel chunk c][tap n][16 B], zero rows for the padding taps
      const float* kwin = p.k + (int64_t(b * p.h + wy0) * p.w + wx0) * p.D + head * DQ;
      for (int i = tid - NW; i < TP * KC; i += NT - NW) {
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
      fence_proxy_async_smem();
      asm volatile("bar.sync 2, %0;" ::"n"(NT - NW) : "memory");
      // resident value head: the whole V window [channel group][tap][16 B], loads batched four deep
      constexpr int NV = NT - NW - 32;     // threads of warps 9 .. 11
      if (v_resident && warp > NW / 32) {
        const int nv = TP * (dv >> 3);
        for (int i0 = tid - NW - 32; i0 < nv; i0 += 4 * NV) {
          float4 va[4], vb[4];
  #pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int i = i0 + k * NV, n = i % TP, g8 = i / TP;
            va[k] = vb[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < nv && n < K2) {
              const int t = n / K, u = n - t * K;
              const float4* src = reinterpret_cast<const float4*>(vwin + (int64_t(t) * p.w + u) * p.C + g8 * 8);
              va[k] = __ldg(src);
              vb[k] = __ldg(src + 1);
            }
          }
  #pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int i = i0 + k * NV;
            if (i < nv) {
              const float x[8] = {va[k].x, va[k].y, va[k].z, va[k].w, vb[k].x, vb[k].y, vb[k].z, vb[k].w};
              uint4 hi, lo;
              split8(x, hi, lo);
              *reinterpret_cast<uint4*>(sVhi + i * 16) = hi;
              *reinterpret_cast<uint4*>(sVlo + i * 16) = lo;
            }
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(&bars.v_ready);
      }
    }
#if defined(NAF_BWD_LANE0)
    constexpr bool kElect = false;   // profiling: the plain form
#else
    constexpr bool kElect = true;
#endif
    __syncwarp();
    if (warp == NW / 32 && elect_one<kElect>()) {
      fence_after_sync();
      // Descriptors are built once and advanced by adding to their address field (bits 0..13, 16-byte units; no carry
      // can leave it: shared memory ends below 2^18 bytes).
      auto adv = [](uint64_t d, uint32_t bytes) { return d + (bytes >> 4); };
      const uint64_t dKk = make_desc(smem_u32(sKhi), TP * 16, 128);        // K window, K-major (B of S)
      const uint64_t dKm = make_desc(smem_u32(sKhi), 128, TP * 16);        // K window, MN-major (B of dQ)
      const uint64_t dVk = make_desc(smem_u32(sVhi), TP * 16, 128);        // V window, K-major (B of dP)
      const uint64_t dT = make_desc(smem_u32(sThi), Cfg::kPG, 128);        // P^T / dS^T, MN-major (A of dVw, dKw)
      const uint64_t dQk0 = make_desc(smem_u32(sQ), 128 * 16, 128);        // Q image 0, K-major (A of S)
      const uint64_t dQm0 = make_desc(smem_u32(sQ), 128, 128 * 16);        // Q image 0, MN-major (B of dKw)
      const uint64_t dGk0 = make_desc(smem_u32(sG), 128 * 16, 128);        // G image 0, K-major (A of dP)
      const uint64_t dGm0 = make_desc(smem_u32(sG), 128, 128 * 16);        // G image 0, MN-major (B of dVw)
      const uint32_t v_lo = uint32_t(v_img);                               // hi -> lo image of the V window
      auto issue_s = [&](int t) {
        // S = Qhi Khi^T + Qlo Khi^T + Qhi Klo^T
        mbar_wait(&bars.q_ready[tb(t)], tpar(t));
        fence_after_sync();
        const uint64_t qa = adv(dQk0, tb(t) * (2 * PIXIMG));
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const uint64_t a0 = adv(qa, pass == 1 ? PIXIMG : 0);
          const uint64_t b0 = adv(dKk, pass == 2 ? Cfg::kWin : 0);
#pragma unroll
          for (int ks = 0; ks < DQ / 16; ++ks)
            mma_f16_ss(tS0 + tb(t) * Cfg::kTA, adv(a0, ks * 2 * (128 * 16)), adv(b0, ks * 2 * (TP * 16)), idesc_s, (pass | ks) != 0);
        }
        commit(&bars.s_done[tb(t)]);
      };
      issue_s(0);
      int s = 0;
      for (int t = 0; t < ntiles; ++t) {
        for (int c = 0; c < nchunks; ++c, ++s) {
          const int wc = min(DVC, dv - c * DVC);
          const uint64_t gk = adv(dGk0, gb(s) * (2 * PIXIMG)), gm = adv(dGm0, gb(s) * (2 * PIXIMG));
          const uint64_t vk = adv(dVk, v_resident ? c * (DVC / 8) * TP * 16 : 0);
          mbar_wait(&bars.g_ready[gb(s)], gpar(s));
          if (s == 0 && v_resident) mbar_wait(&bars.v_ready, 0);
          fence_after_sync();
          // dP (+)= Ghi Vhi^T + Glo Vhi^T + Ghi Vlo^T
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            const uint64_t a0 = adv(gk, pass == 1 ? PIXIMG : 0);
            const uint64_t b0 = adv(vk, pass == 2 ? v_lo : 0);
#pragma unroll
            for (int ks = 0; ks < DVC / 16; ++ks)
              if (ks < (wc >> 4) && !(NAF_BWD_EXP & 4))
                mma_f16_ss(tdP, adv(a0, ks * 2 * (128 * 16)), adv(b0, ks * 2 * (TP * 16)), idesc_s, (c | pass | ks) != 0);
          }
          if (c == 0) {
            mbar_wait(&bars.pt_ready, t & 1);
            fence_after_sync();
          }
          // dVw[:, chunk] (+)= P^T_hi Ghi + P^T_lo Ghi + P^T_hi Glo   (A = P^T MN-major, B = G seen MN-major: K = pixels)
          const uint32_t idesc_dv = make_idesc_f16(128, wc, true, true);
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            const uint64_t a0 = adv(dT, pass == 1 ? Cfg::kT : 0);
            const uint64_t b0 = adv(gm, pass == 2 ? PIXIMG : 0);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              if (!(NAF_BWD_EXP & 1)) mma_f16_ss(tdV + c * DVC, adv(a0, ks * 2 * Cfg::kPG), adv(b0, ks * 2 * 128), idesc_dv, (t | pass | ks) != 0);
          }
          commit(&bars.g_done[gb(s)]);
        }
        mbar_wait(&bars.ds_ready, t & 1);
        fence_after_sync();
        const uint32_t tS = tS0 + tb(t) * Cfg::kTA;
        // dQ = dShi Khi + dSlo Khi + dShi Klo      (B = the K window seen MN-major: N = channels, K = taps)
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t a0 = tS + (pass == 1 ? TP / 2 : 0);
          const uint64_t b0 = adv(dKm, pass == 2 ? Cfg::kWin : 0);
#pragma unroll
          for (int ks = 0; ks < TP / 16; ++ks)
            mma_f16_ts(tdP, a0 + ks * 8, adv(b0, ks * 2 * 128), idesc_dq, (pass | ks) != 0);
        }
        // dKw (+)= dS^T_hi Qhi + dS^T_lo Qhi + dS^T_hi Qlo
        const uint64_t qm = adv(dQm0, tb(t) * (2 * PIXIMG));
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const uint64_t a0 = adv(dT, pass == 1 ? Cfg::kT : 0);
          const uint64_t b0 = adv(qm, pass == 2 ? PIXIMG : 0);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            if (!(NAF_BWD_EXP & 2)) mma_f16_ss(tdK, adv(a0, ks * 2 * Cfg::kPG), adv(b0, ks * 2 * 128), idesc_dk, (t | pass | ks) != 0);
        }
        commit(&bars.dq_done);
        if (t + 1 < ntiles) issue_s(t + 1);
      }
    }
    __syncwarp();
  } else {
    // ================================================================ workers: two threads per pixel row
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    load_q(0);
    load_g(0, 0, g8r[0]);
    if constexpr (kDeepG) {
      if (nchunks > 1) load_g(0, 1, g8r[1]);
      else load_g(1, 0, g8r[1]);
    }
    const int tap0 = hf * SC;
    uint8_t* const tdst = sThi + (row >> 3) * Cfg::kPG + (tap0 >> 3) * 128 + (row & 7) * 16;
    auto next_q = [&](int t) {      // stage the queries of tile t, start loading those of tile t + 1
      stage_q(t, sQ + tb(t) * (2 * PIXIMG));
      load_q(t + 1);
      fence_proxy_async_smem();
      mbar_arrive(&bars.q_ready[tb(t)]);
    };
    next_q(0);
    int gstep = 0;
    for (int tile = 0; tile < ntiles; ++tile) {
      const Px nxt = px_of(tile + 1);
      float pr[SC];
      // ================= value-head chunks: dP += G_c Vw_c^T,  dVw[:, chunk c] += P^T G_c
      for (int c = 0; c < nchunks; ++c, ++gstep) {
        const int wc = min(DVC, dv - c * DVC);          // channels in this chunk (multiple of 16)
        uint8_t* const sGhi = sG + gb(gstep) * (2 * PIXIMG);
        uint8_t* const sGlo = sGhi + PIXIMG;
        // the MMAs that read this G buffer (step gstep - gbufs) must have completed; with one buffer this also frees
        // the V chunk buffer of a value head that is not resident
        if (gstep >= gbufs) mbar_wait(&bars.g_done[gb(gstep - gbufs)], gpar(gstep - gbufs));
        // G chunk (prefetched): K-major [channel group][row][16 B]; then the loads of the chunk 1 or 2 steps ahead
        auto stage_g = [&](float (&set)[4][8]) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int bi = k * 8 + warp, g8 = (bi >> 4) * 4 + lg;
            if (g8 < (wc >> 3)) {
              uint4 hi, lo;
              split8(set[k], hi, lo);
              *reinterpret_cast<uint4*>(sGhi + (g8 * 128 + (bi & 15) * 8 + lp) * 16) = hi;
              *reinterpret_cast<uint4*>(sGlo + (g8 * 128 + (bi & 15) * 8 + lp) * 16) = lo;
            }
          }
          int t2 = tile, c2 = c + (kDeepG ? 2 : 1);
          while (c2 >= nchunks) c2 -= nchunks, ++t2;
          load_g(t2, c2, set);     // (tiles past the last one load zeros)
        };
        if (kDeepG && (gstep & 1)) stage_g(g8r[kDeepG ? 1 : 0]);
        else stage_g(g8r[0]);
        // V window chunk of a value head that is not resident: K-major [channel group][tap][16 B]
        if (!v_resident) {
          uint8_t* const vhi = sVhi;
          uint8_t* const vlo = sVlo;
          constexpr int VIT = (TP * 8 + NW - 1) / NW;     // items per thread at the full chunk width
          float4 va[VIT], vb[VIT];
#pragma unroll
          for (int k = 0; k < VIT; ++k) {
            const int i = tid + k * NW;
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
            const int i = tid + k * NW;
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
        mbar_arrive(&bars.g_ready[gb(gstep)]);
        NAF_TR();   // D: chunk staged

        if (c == 0) {
          // ================= P = softmax(scale S) on this thread's half row; P^T (fp16 hi / lo) -> shared memory
          // (the tensor core runs dP of chunk 0 meanwhile)
          mbar_wait(&bars.s_done[tb(tile)], tpar(tile));
          fence_after_sync();
          NAF_TR();   // B: S complete
          uint32_t sv[SC];
#pragma unroll
          for (int c0 = 0; c0 < SC; c0 += 8)
            tmem_ld8(tS0 + tb(tile) * Cfg::kTA + lane_off + hf * SC + c0, *reinterpret_cast<uint32_t(*)[8]>(&sv[c0]));
          wait_ld();
          float m = -INFINITY;
#pragma unroll
          for (int j = 0; j < SC; ++j)
            if (tap0 + j < K2) m = fmaxf(m, __uint_as_float(sv[j]));
          red_a[hf][row] = m;
          worker_sync();
          const float mq = fmaxf(red_a[0][row], red_a[1][row]) * qscale;
          float l = 0.f;
#pragma unroll
          for (int j = 0; j < SC; ++j) {
            const float e = (tap0 + j < K2) ? fast_exp2(fmaf(__uint_as_float(sv[j]), qscale, -mq)) : 0.f;
            pr[j] = e;
            l += e;
          }
          red_b[hf][row] = l;
          worker_sync();
          const float inv = 1.f / (red_b[0][row] + red_b[1][row]);
#pragma unroll
          for (int j = 0; j < SC; ++j) pr[j] *= inv;
#pragma unroll
          for (int g8 = 0; g8 < SC / 8; ++g8) {
            uint4 hi, lo;
            split8(*reinterpret_cast<float(*)[8]>(&pr[8 * g8]), hi, lo);
            *reinterpret_cast<uint4*>(tdst + g8 * 128) = hi;
            *reinterpret_cast<uint4*>(tdst + Cfg::kT + g8 * 128) = lo;
          }
          fence_proxy_async_smem();
          mbar_arrive(&bars.pt_ready);
          NAF_TR();   // C: softmax, P^T written
        }
      }
      // dP is complete and P^T may be overwritten once every chunk step of the tile has completed
      if (gbufs == 2 && gstep >= 2) mbar_wait(&bars.g_done[gb(gstep - 2)], gpar(gstep - 2));
      mbar_wait(&bars.g_done[gb(gstep - 1)], gpar(gstep - 1));
      fence_after_sync();
      NAF_TR();   // E: all chunk MMAs complete

      // ================= dS = scale P (dP - delta): TMEM (A operand of dQ, hi | lo) and dS^T -> the T buffer
      {
        const uint32_t tS = tS0 + tb(tile) * Cfg::kTA;
        uint32_t d[SC];
#pragma unroll
        for (int c0 = 0; c0 < SC; c0 += 8) tmem_ld8(tdP + lane_off + hf * SC + c0, *reinterpret_cast<uint32_t(*)[8]>(&d[c0]));
        wait_ld();
        float dl = 0.f;
#pragma unroll
        for (int j = 0; j < SC; ++j) dl = fmaf(pr[j], __uint_as_float(d[j]), dl);
        red_a[hf][row] = dl;
        worker_sync();
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
          *reinterpret_cast<uint4*>(tdst + Cfg::kT + (c0 >> 3) * 128) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        wait_st();
      }
      fence_proxy_async_smem();
      fence_before_sync();   // dP was read and dS written: dQ may overwrite dP, and read dS, after this arrival
      mbar_arrive(&bars.ds_ready);
      NAF_TR();   // F: dS written

      // the queries of the next tile go into the other Q buffer while the tensor core runs dQ / dKw
      if (overlap && tile + 1 < ntiles) next_q(tile + 1);
      mbar_wait(&bars.dq_done, tile & 1);
      fence_after_sync();
      NAF_TR();   // G: dQ / dKw complete

      // ================= dQ -> global: this thread's 32 columns of its pixel row (a line of the output, four 32-byte
      // stores; staging the tile in shared memory for whole-line warp stores was measured no faster)
      {
        uint32_t r[32];
        tmem_ld32(tdP + lane_off + hf * 32, r);
        wait_ld();
        fence_before_sync();   // the next chunk step's arrival lets dP overwrite these columns
        if (cur.valid) {
          float* dst = p.dq + cur.pix * p.D + head * DQ + hf * 32;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[8 * j + e]);
            stg_stream8(dst + 8 * j, v);
          }
        }
      }
      if (!overlap && tile + 1 < ntiles) next_q(tile + 1);
      cur = nxt;
      NAF_TR();   // H: dQ stored
    }
  }
  // every MMA of the cell has completed (the workers waited for the last dq_done)
  fence_before_sync();
  __syncthreads();
  fence_after_sync();

  // ================= window gradients: TMEM lane = tap; one vector atomic per 4 elements
  if (warp < NW / 32) {
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
  NAF_TR();   // Z: window gradients sent
  if (warp == 0) tmem_dealloc(tmem, 512);
}

#if NAF_BWD_EXP & 32
extern "C" NAF_API int naf_debug_bwd_trace(long long* host) {
  return int(cudaMemcpyFromSymbol(host, g_bwd_trace, sizeof(g_bwd_trace)));
}
#endif

// ------------------------------------------------------------------------------------ host side
namespace {

int taps_pad(int K) { return (K * K + 15) / 16 * 16; }

template <int TP>
int launch_bwd_tc(const naf_xattn_bwd_params& p, cudaStream_t st) {
  auto kern = xattn_bwd_cell_tc_kernel<TP>;
  const int dv = p.C / p.heads;
  // Shared memory: T, K window, V region, Q and G images (hi / lo each).  Preferences, as far as they fit:
  // the whole value head resident (staged once per cell; else one 64-channel chunk restaged per tile and chunk),
  // the upstream-gradient chunk double buffered (its MMAs run while the next chunk is staged), and a second Q
  // image + S region ("overlap": the next tile's queries are staged and multiplied while dQ / dKw of this one run).
  using Cfg = BwdCfg<TP>;
  const int img = 2 * PIXIMG;
  const int v_full = 2 * (dv / 8) * TP * 16, v_chunk = 2 * Cfg::kWin;
  const int budget = 224 * 1024;
  const bool tmem2 = 2 * Cfg::kTA + (Cfg::kTA > DQ ? Cfg::kTA : DQ) + DQ + dv <= 512;
  auto need = [&](int vres, int gb, int ov) { return 2 * Cfg::kT + 2 * Cfg::kWin + (vres ? v_full : v_chunk) + (1 + ov) * img + gb * img; };
  int v_resident = 0, gbufs = 1, overlap = 0;
  const int prefs[4][3] = {{1, 2, 1}, {1, 2, 0}, {1, 1, 1}, {1, 1, 0}};
  for (const auto& c : prefs) {
    if ((c[2] && !tmem2) || need(c[0], c[1], c[2]) > budget) continue;
    v_resident = c[0], gbufs = c[1], overlap = c[2];
    break;
  }
#ifdef NAF_BWD_NO_OVERLAP
  overlap = 0;
#endif
  const int smem = need(v_resident, gbufs, overlap);
  cudaError_t e = ensure_dyn_smem(kern, smem);
  if (e != cudaSuccess) return fail(NAF_ERR_CUDA, "xattn_bwd(cell-tc): smem opt-in failed: %s", cudaGetErrorString(e));
  const unsigned grid = unsigned(p.B) * p.h * p.w * p.heads;
  kern<<<grid, NT, smem, st>>>(p, p.Ho / p.h, p.Wo / p.w, dv, v_resident, gbufs, overlap, magic_of(p.Wo / p.w), magic_of(p.rep_y),
                               magic_of(p.rep_x));
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
  if (!aligned32(p.q) || !aligned16(p.k) || !aligned16(p.v) || !aligned32(p.dout) || !aligned32(p.dq) ||
      !aligned16(p.dk) || !aligned16(p.dv) || (p.q_stride_b % 8) || (p.q_stride_y % 8) || (p.q_stride_x % 8)) {
    *why = "pointers / strides not 32-byte (q, dout, dq) / 16-byte aligned";
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
