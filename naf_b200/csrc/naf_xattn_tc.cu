// Cell kernel, tensor-core path (tcgen05 / TMEM): cross-scale neighbourhood attention for
// integer ratios with 64-wide heads.
//
// One CTA (128 threads = 128 TMEM lanes) owns one (batch, low-res cell, head).  All rh*rw target
// pixels of the cell attend over the same clamped K x K low-res window, so per 128-pixel tile
//     S[128 x TP]  = Q[128 x 64] * Kwin[TP x 64]^T          (TP = K*K padded to 16)
//     O[128 x DV]  = softmax(S)[128 x TP] * Vwin[TP x DV]
// are two dense GEMMs issued with tcgen05.mma (M=128, fp32 accumulators in TMEM):
//   * thread <-> pixel row <-> TMEM lane: q row loaded straight from HBM (RoPE on the fly),
//     written to shared memory in the canonical K-major UMMA layout; softmax is thread-local on
//     the S row read back with tcgen05.ld; P goes back to TMEM (over the S columns) and feeds the
//     second GEMM as its A operand; O rows come back with tcgen05.ld, are normalised and leave
//     through a per-warp smem transpose as 128 B coalesced streaming stores.
//   * precision: every operand is split fp32 -> fp16 hi + fp16 lo and each GEMM is the three
//     passes hi*hi + lo*hi + hi*lo (~22 mantissa bits, fp32 accumulation), so the result stays
//     within ~1e-5 of the fp32 SIMT kernel while running at kind::f16 rate.
//   * 2 CTAs per SM (96 KB smem, 256 TMEM columns each) overlap one CTA's loads / softmax /
//     epilogue with the other's MMAs; the next tile's q rows are prefetched into registers.
// HBM traffic = algorithmic minimum (q once, out once; windows from L2).
//
// Reference semantics: src/layers/attentions.py:16-29,53-75; RoPE src/layers/rope.py:137-153.
#include "naf_common.cuh"
#include "naf_umma.cuh"

namespace naf {

using namespace umma;

namespace {

constexpr int DQ = 64;        // query/key head dim this kernel is specialised for
constexpr int KC = DQ / 8;    // 16-byte chunks along the head dim
constexpr int STAGE_STRIDE = 36;  // floats per staged row: 32 columns + 4 pad (bank spread)

template <int TP, int DV>
struct TcCfg {
  static constexpr int kSmemK = KC * TP * 16;        // one of hi / lo
  static constexpr int kSmemV = TP * DV * 2;         // one of hi / lo
  static constexpr int kSmemQ = KC * 128 * 16;       // one of hi / lo
  static constexpr int kSmemTotal = 2 * kSmemK + 2 * kSmemV + 2 * kSmemQ;
  static constexpr int kTmemCols = (TP + DV) <= 64 ? 64 : (TP + DV) <= 128 ? 128 : (TP + DV) <= 256 ? 256 : 512;
  static_assert(TP % 16 == 0 && TP <= 256 && DV % 32 == 0 && DV <= 256 && TP + DV <= 512, "bad tile");
  static_assert(4 * 32 * STAGE_STRIDE * 4 <= 2 * kSmemQ, "staging must fit in the Q tile");
};

__device__ __forceinline__ void split8(const float (&x)[8], uint4& hi, uint4& lo) {
  __half h[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) split_f16(x[i], h[i], l[i]);
  hi = make_uint4(pack_half2(h[0], h[1]), pack_half2(h[2], h[3]), pack_half2(h[4], h[5]), pack_half2(h[6], h[7]));
  lo = make_uint4(pack_half2(l[0], l[1]), pack_half2(l[2], l[3]), pack_half2(l[4], l[5]), pack_half2(l[6], l[7]));
}

}  // namespace

template <int TP, int DV>
__global__ void __launch_bounds__(128)
xattn_cell_tc_kernel(naf_xattn_params p, int rh, int rw) {
  using Cfg = TcCfg<TP, DV>;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t mbar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ uint32_t row_pix[128];

  uint8_t* sKhi = smem;
  uint8_t* sKlo = sKhi + Cfg::kSmemK;
  uint8_t* sVhi = sKlo + Cfg::kSmemK;
  uint8_t* sVlo = sVhi + Cfg::kSmemV;
  uint8_t* sQhi = sVlo + Cfg::kSmemV;
  uint8_t* sQlo = sQhi + Cfg::kSmemQ;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* stage = reinterpret_cast<float*>(sQhi) + warp * 32 * STAGE_STRIDE;  // aliases the Q tile

  const int K = p.K, K2 = K * K;
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
  for (int i = tid; i < TP * KC; i += 128) {
    const int n = i % TP, c = i / TP;
    uint4 hi = make_uint4(0, 0, 0, 0), lo = hi;
    if (n < K2) {
      const int t = n / K, u = n - t * K;
      const float* src = p.k + (int64_t(b * p.h + wy0 + t) * p.w + wx0 + u) * p.D + head * DQ + c * 8;
      const float4 f0 = *reinterpret_cast<const float4*>(src);
      const float4 f1 = *reinterpret_cast<const float4*>(src + 4);
      const float x[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
      split8(x, hi, lo);
    }
    *reinterpret_cast<uint4*>(sKhi + (c * TP + n) * 16) = hi;
    *reinterpret_cast<uint4*>(sKlo + (c * TP + n) * 16) = lo;
  }
  // ---- stage the V window: canonical MN-major [tap group][channel group][tap%8][16 B]
  {
    constexpr int NG = DV / 8;
    for (int i = tid; i < TP * NG; i += 128) {
      const int kk = i & 7;
      const int g = (i >> 3) % NG;
      const int kg = (i >> 3) / NG;
      const int k = kg * 8 + kk;
      uint4 hi = make_uint4(0, 0, 0, 0), lo = hi;
      if (k < K2) {
        const int t = k / K, u = k - t * K;
        const float* src = p.v + (int64_t(b * p.h + wy0 + t) * p.w + wx0 + u) * p.C + head * DV + g * 8;
        const float4 f0 = *reinterpret_cast<const float4*>(src);
        const float4 f1 = *reinterpret_cast<const float4*>(src + 4);
        const float x[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
        split8(x, hi, lo);
      }
      const int off = (kg * NG + g) * 128 + kk * 16;
      *reinterpret_cast<uint4*>(sVhi + off) = hi;
      *reinterpret_cast<uint4*>(sVlo + off) = lo;
    }
  }

  const bool rope = p.cos_y != nullptr;
  constexpr int HALF = DQ / 2, P = DQ / 4;
  const int npix = rh * rw;
  const int ntiles = (npix + 127) >> 7;
  const int y0 = ci * rh, x0 = cj * rw;
  const float qscale = p.scale * 1.4426950408889634f;  // fold log2(e): softmax via exp2
  const float* qbase = p.q + int64_t(b) * p.q_stride_b + head * DQ;
  float* obase = p.out + int64_t(b) * p.Ho * p.Wo * p.C + head * DV;

  float q[DQ];
  uint32_t my_pix = 0xFFFFFFFFu;  // linear target pixel of this thread's row, or invalid

  // q row of pixel `pi` of this cell -> registers (rotated, pre-scaled)
  auto load_q = [&](int tile) {
    int pi = tile * 128 + tid;
    const bool valid = pi < npix;
    if (!valid) pi = npix - 1;
    const int py = pi / rw;
    const int y = y0 + py, x = x0 + (pi - py * rw);
    my_pix = valid ? uint32_t(y * p.Wo + x) : 0xFFFFFFFFu;
    const float* qp = qbase + int64_t(y / p.rep_y) * p.q_stride_y + int64_t(x / p.rep_x) * p.q_stride_x;
#pragma unroll
    for (int i = 0; i < DQ / 4; ++i) {
      const float4 t = ldg_stream(qp + 4 * i);
      q[4 * i] = t.x; q[4 * i + 1] = t.y; q[4 * i + 2] = t.z; q[4 * i + 3] = t.w;
    }
    if (rope) {
#pragma unroll
      for (int i4 = 0; i4 < HALF / 4; ++i4) {
        const bool on_y = (i4 * 4) < P;
        const float* ct = on_y ? p.cos_y + int64_t(y) * P + i4 * 4 : p.cos_x + int64_t(x) * P + (i4 * 4 - P);
        const float* st = on_y ? p.sin_y + int64_t(y) * P + i4 * 4 : p.sin_x + int64_t(x) * P + (i4 * 4 - P);
        const float4 c4 = *reinterpret_cast<const float4*>(ct);
        const float4 s4 = *reinterpret_cast<const float4*>(st);
        const float cc[4] = {c4.x, c4.y, c4.z, c4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float a = q[i4 * 4 + j], bb = q[HALF + i4 * 4 + j];
          q[i4 * 4 + j] = a * cc[j] - bb * ss[j];
          q[HALF + i4 * 4 + j] = bb * cc[j] + a * ss[j];
        }
      }
    }
#pragma unroll
    for (int i = 0; i < DQ; ++i) q[i] *= qscale;
  };

  load_q(0);

  // wait for the TMEM base address and make barrier inits visible
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_off = uint32_t(warp * 32) << 16;
  const uint32_t tS = tmem;            // S (fp32, TP cols)  /  P (fp16 hi: TP/2 cols, lo: TP/2 cols)
  const uint32_t tO = tmem + TP;       // O (fp32, DV cols)
  constexpr uint32_t idesc_qk = make_idesc_f16(128, TP, false, false);
  constexpr uint32_t idesc_pv = make_idesc_f16(128, DV, false, true);
  uint32_t phase = 0;

  for (int tile = 0; tile < ntiles; ++tile) {
    // ================= stage Q (thread <-> row), canonical K-major [chunk][row][16 B]
#pragma unroll
    for (int c = 0; c < KC; ++c) {
      const float x[8] = {q[8 * c], q[8 * c + 1], q[8 * c + 2], q[8 * c + 3],
                          q[8 * c + 4], q[8 * c + 5], q[8 * c + 6], q[8 * c + 7]};
      uint4 hi, lo;
      split8(x, hi, lo);
      *reinterpret_cast<uint4*>(sQhi + (c * 128 + tid) * 16) = hi;
      *reinterpret_cast<uint4*>(sQlo + (c * 128 + tid) * 16) = lo;
    }
    row_pix[tid] = my_pix;
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
          mma_f16_ss(tS, da, db, idesc_qk, (pass | kk) != 0);
        }
      }
      commit(&mbar[0]);
    }
    // ================= prefetch the next tile's q rows while the tensor core works
    if (tile + 1 < ntiles) load_q(tile + 1);

    mbar_wait(&mbar[0], phase);
    fence_after_sync();

    // ================= softmax on this thread's S row; P (hi|lo fp16) back into the S columns
    float inv_l;
    {
      uint32_t s[TP];
#pragma unroll
      for (int c0 = 0; c0 < TP; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tS + lane_off + c0, r);
#pragma unroll
        for (int j = 0; j < 16; ++j) s[c0 + j] = r[j];
      }
      wait_ld();
      float m = -INFINITY;
#pragma unroll
      for (int j = 0; j < TP; ++j)
        if (j < K2) m = fmaxf(m, __uint_as_float(s[j]));
      float l = 0.f;
#pragma unroll
      for (int j = 0; j < TP; ++j) {
        const float e = (j < K2) ? fast_exp2(__uint_as_float(s[j]) - m) : 0.f;
        l += e;
        s[j] = __float_as_uint(e);
      }
      inv_l = 1.f / l;
#pragma unroll
      for (int c0 = 0; c0 < TP; c0 += 16) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          __half h0, l0, h1, l1;
          split_f16(__uint_as_float(s[c0 + 2 * j]), h0, l0);
          split_f16(__uint_as_float(s[c0 + 2 * j + 1]), h1, l1);
          hi[j] = pack_half2(h0, h1);
          lo[j] = pack_half2(l0, l1);
        }
        tmem_st8(tS + lane_off + c0 / 2, hi);
        tmem_st8(tS + lane_off + TP / 2 + c0 / 2, lo);
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
          mma_f16_ts(tO, a0 + kk * 8, db, idesc_pv, (pass | kk) != 0);
        }
      }
      commit(&mbar[1]);
    }
    mbar_wait(&mbar[1], phase);
    fence_after_sync();

    // ================= epilogue: O rows -> normalise -> per-warp smem transpose -> coalesced stores
#pragma unroll 1
    for (int c0 = 0; c0 < DV; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tO + lane_off + c0, r);
      wait_ld();
      float* srow = stage + lane * STAGE_STRIDE;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        *reinterpret_cast<float4*>(srow + j) =
            make_float4(__uint_as_float(r[j]) * inv_l, __uint_as_float(r[j + 1]) * inv_l,
                        __uint_as_float(r[j + 2]) * inv_l, __uint_as_float(r[j + 3]) * inv_l);
      }
      __syncwarp();
      // 4 rows per instruction: lanes 8i..8i+7 carry the 128 B of row (4*it + i)
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int rr = it * 4 + (lane >> 3);
        const uint32_t pix = row_pix[warp * 32 + rr];
        const float4 v = *reinterpret_cast<const float4*>(stage + rr * STAGE_STRIDE + (lane & 7) * 4);
        if (pix != 0xFFFFFFFFu) stg_stream(obase + int64_t(pix) * p.C + c0 + (lane & 7) * 4, v);
      }
      __syncwarp();
    }
    fence_before_sync();
    __syncthreads();  // staging aliases the Q tile; O / S columns are rewritten by the next tile
    phase ^= 1;
  }

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
  kern<<<grid, 128, Cfg::kSmemTotal, st>>>(p, p.Ho / p.h, p.Wo / p.w);
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
  if (tp != 16 && tp != 32 && tp != 64 && tp != 96 && tp != 128) { *why = "kernel_size must be <= 11"; return false; }
  if ((p.Ho / p.h) * (p.Wo / p.w) < 64) { *why = "fewer than 64 pixels per cell"; return false; }
  if (!aligned16(p.q) || !aligned16(p.k) || !aligned16(p.v) || !aligned16(p.out) ||
      (p.q_stride_b % 4) || (p.q_stride_y % 4) || (p.q_stride_x % 4)) {
    *why = "pointers/strides not 16-byte aligned";
    return false;
  }
  if (p.cos_y && !(aligned16(p.cos_y) && aligned16(p.sin_y) && aligned16(p.cos_x) && aligned16(p.sin_x))) {
    *why = "rope tables not 16-byte aligned";
    return false;
  }
  if (int64_t(p.Ho) * p.Wo >= int64_t(0xFFFFFFFFu)) { *why = "target map too large"; return false; }
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
