// Guidance conv encoder on the 5th-gen tensor core.
//
// The reference's encoder (src/layers/convolutions.py:70-95) is a stem Conv2d(3 -> 128, k, reflect)
// followed by EncBlocks  GroupNorm(8) -> SiLU -> Conv2d(128 -> 128, k, reflect)  x2  (k = 1 for the
// "pixel" branch, 3 for the "semantic" branch; src/model/naf.py:26-27).  Run layer by layer every
// activation (822 MB at 8x448x448x128 fp32) crosses HBM five times per GroupNorm+conv pair
// (statistics read, normalise read + write, conv read + write).  Here it crosses twice:
//
//   stem_conv_kernel      image -> y0 (+bias), per-tile GroupNorm partial sums of y0
//   gn_coef_kernel        partial sums -> per (image, channel) scale/shift  (deterministic, double)
//   conv128_tc_kernel     y_in -> [scale/shift + SiLU + fp16 split, reflect halo, in shared memory]
//                         -> implicit GEMM on tcgen05 (accumulator in TMEM) -> +bias -> y_out,
//                         per-tile partial sums of y_out for the NEXT GroupNorm
//
// conv128_tc_kernel: persistent CTAs, one 16x8 pixel tile (M = 128) at a time.
//   warps 0-7  workers : prologue (halo tile HBM -> GN affine -> SiLU -> fp16 -> smem, K-major
//                        canonical layout [channel chunk][halo y][halo x][16 B]) and epilogue
//                        (TMEM -> +bias -> statistics -> per-warp smem transpose -> 128-byte-run stores)
//   warp 8     mma     : for each tap (dy,dx) the A operand is the SAME halo tile addressed through a
//                        shifted descriptor (start += (dy*WX+dx)*16 B, 8-row-group stride = one halo
//                        row): no im2col copy exists anywhere.  B = that tap's 128x128 weights.
//   warp 9     loader  : streams the pre-packed fp16 weights of each tap from L2 into a 2-slot ring
//                        with cp.async.bulk + mbarrier complete_tx.
// PASSES = 1: operands rounded to fp16 (10-bit mantissa, the precision class of the TF32 convolution
//             the reference runs by default on GPU: torch.backends.cudnn.allow_tf32 = True);
//             two CTAs per SM so one tile's prologue/epilogue overlaps the other's MMAs.
// PASSES = 3: fp16 hi/lo split of both operands, hi*hi + lo*hi + hi*lo with fp32 accumulation
//             (~22 mantissa bits: the strict-fp32 class, allow_tf32 = False).
#include "naf_common.cuh"
#include "naf_umma.cuh"

#ifndef NAF_CONV_ITEMPIPE
#define NAF_CONV_ITEMPIPE 1   // producers: 1 = per-item software pipeline (depth NIT/2), 0 = two register batches
#endif
#ifndef NAF_CONV_PF_DIST
#define NAF_CONV_PF_DIST 2   // L2 prefetch distance of the producers, in tiles
#endif
#ifndef NAF_CONV_EXP
#define NAF_CONV_EXP 0   // profiling variants (scripts/conv_experiments.sh): 1 no output stores,
#endif                   // 2 no input loads, 4 no MMAs, 8 weights loaded once, 16 no SiLU

namespace naf {

using namespace umma;

namespace {

constexpr int CC = 128;              // channels in == channels out
constexpr int NCHUNK = CC / 8;       // 16-byte fp16 chunks per pixel
constexpr int TH = 16, TW = 8;       // pixel tile: 16 rows x 8 columns = 128 MMA rows
constexpr int NWORK = 256;
constexpr int NTHREADS = NWORK + 64;
constexpr int MMA_WARP = NWORK / 32, LOAD_WARP = MMA_WARP + 1;
constexpr int W_PLANE = NCHUNK * CC * 16;   // one tap, one plane (hi or lo): 32 KB
constexpr int STAGE_SLOT = 128 + 16;        // epilogue staging slot per worker thread (+ bank pad)

template <int KS, int PASSES>
struct ConvCfg {
  static constexpr int HY = TH + KS - 1, WX = TW + KS - 1, HP = HY * WX;
  static constexpr int CS = HP * 16 + 16;     // chunk stride; the 16 B pad spreads the 16 chunks over banks
  static constexpr int NPLANE = PASSES == 1 ? 1 : 2;
  static constexpr int A_PLANE = NCHUNK * CS;
  static constexpr int STAGE = NWORK * STAGE_SLOT;
  static constexpr int A_RAW = NPLANE * A_PLANE > STAGE ? NPLANE * A_PLANE : STAGE;
  static constexpr int A_BYTES = (A_RAW + 127) / 128 * 128;
  static constexpr int NT = KS * KS;
  static constexpr int W_GRAN = NPLANE * W_PLANE;   // bytes per tap in the ring
  static constexpr int NSLOT = NT == 1 ? 1 : 2;
  static constexpr int SMEM = A_BYTES + NSLOT * W_GRAN;
  static constexpr int CTAS_PER_SM = PASSES == 1 ? 2 : 1;
};

// Activation element type between the layers: fp32, or -- TF32 class only -- fp16 (`IN16` / `OUT16` template
// flags): the 1-pass kernels round their A operand to fp16 anyway, so storing the conv output as fp16 halves
// the HBM bytes of every layer and the load instructions / in-flight registers of the producers.  GroupNorm
// statistics are always taken from the fp32 accumulators.
struct ConvParams {
  const void* in;         // (B,H,W,128) contiguous, fp32 or fp16
  const float2* coef;     // (B,128) {scale, shift} of the GroupNorm applied to `in`
  const uint8_t* wpack;   // [tap][plane hi|lo][chunk 16][n 128][8 x fp16]
  const float* bias;      // (128) or NULL
  void* out;              // pixel-major, out_pix_stride elements per pixel, channels [out_ch_off, +128); fp32 or fp16
  float* part;            // NULL or (B, tiles_y*tiles_x, 8 groups, {sum, sumsq})
  int64_t out_pix_stride;
  int out_ch_off;
  int B, H, W, tiles_y, tiles_x;
};

__device__ __forceinline__ int reflect_clamp(int i, int n) {
  i = i < 0 ? -i : i;
  i = i >= n ? 2 * n - 2 - i : i;
  return i < 0 ? 0 : (i >= n ? n - 1 : i);   // only reached by halo pixels of ragged tiles
}

template <bool FAST>
__device__ __forceinline__ float silu(float t) {
  if constexpr (FAST) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * t));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
    return t * r;
  } else {
    return t / (1.f + expf(-t));
  }
}

// packed fp32x2 helpers (FFMA2 on sm_100): two floats in one 64-bit register
__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

// SiLU of two values, rounded to fp16 (1-pass kernels: the result feeds an fp16 MMA operand).
// silu(t) = t * sigmoid(t) = t * (0.5 + 0.5 * tanh(t / 2)) with ONE MUFU op per element
// (tanh.approx.f32, max relative error 2^-11: below the fp16 rounding that follows) instead of two
// (ex2 + rcp): the producers of the 3x3 kernel were MUFU / issue bound.
#ifndef NAF_CONV_SILU_TANH
#define NAF_CONV_SILU_TANH 1
#endif
__device__ __forceinline__ uint32_t silu2_f16(uint64_t t2) {
  float t0, t1;
  unpack2(t2, t0, t1);
#if NAF_CONV_SILU_TANH
  const uint64_t half2c = pack2(0.5f, 0.5f);
  float h0, h1, th0, th1;
  unpack2(fma2(t2, half2c, 0ull), h0, h1);
  asm("tanh.approx.f32 %0, %1;" : "=f"(th0) : "f"(h0));
  asm("tanh.approx.f32 %0, %1;" : "=f"(th1) : "f"(h1));
  float r0, r1;
  unpack2(fma2(t2, fma2(pack2(th0, th1), half2c, half2c), 0ull), r0, r1);
#else
  float e0, e1, r0, r1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(-1.4426950408889634f * t0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(-1.4426950408889634f * t1));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(1.f + e0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(1.f + e1));
  r0 *= t0;
  r1 *= t1;
#endif
  const __half2 h = __floats2half2_rn(r0, r1);
  return *reinterpret_cast<const uint32_t*>(&h);
}

__device__ __forceinline__ void named_bar_workers() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

}  // namespace

template <int KS, int PASSES>
__global__ void __launch_bounds__(NTHREADS, ConvCfg<KS, PASSES>::CTAS_PER_SM)
conv128_tc_kernel(ConvParams p) {
  using Cfg = ConvCfg<KS, PASSES>;
  constexpr int WX = Cfg::WX, HP = Cfg::HP, CS = Cfg::CS, NT = Cfg::NT;
  constexpr bool kResident = NT == 1;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_a_full, bar_acc_full, bar_w_full[2], bar_w_free[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float2 s_coef[CC];
  __shared__ float s_bias[CC];
  __shared__ float s_part[8][8];

  uint8_t* sA = smem;
  uint8_t* sW = smem + Cfg::A_BYTES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles_per_img = p.tiles_y * p.tiles_x;
  const int total = p.B * tiles_per_img;

  if (warp == MMA_WARP) tmem_alloc(&tmem_base_s, 128);
  if (tid == 0) {
    mbar_init(&bar_a_full, NWORK);
    mbar_init(&bar_acc_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_w_full[s], 1);
      mbar_init(&bar_w_free[s], 1);
    }
    fence_mbar_init();
  }
  if (tid < CC) s_bias[tid] = p.bias ? p.bias[tid] : 0.f;
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  if (warp < MMA_WARP) {
    // ================================================================================ WORKERS
    const int chunk = tid & 15;           // this thread's 8 input channels (prologue)
    const int px0 = tid >> 4;
    const int row = (warp & 3) * 32 + lane, hf = warp >> 2;   // epilogue: TMEM lane, column half
    const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;
    uint8_t* my_stage = sA + tid * STAGE_SLOT;
    const uint8_t* warp_stage = sA + warp * 32 * STAGE_SLOT;
    int prev_tile = -1;
    auto flush_part = [&]() {   // after a workers barrier: the partial sums of the previous tile
      if (p.part && prev_tile >= 0 && tid < 16) {
        const int g = tid >> 1, which = tid & 1, w0 = (g >> 2) * 4, gl = g & 3;
        const float s = (s_part[w0][gl * 2 + which] + s_part[w0 + 1][gl * 2 + which]) +
                        (s_part[w0 + 2][gl * 2 + which] + s_part[w0 + 3][gl * 2 + which]);
        p.part[int64_t(prev_tile) * 16 + tid] = s;
      }
    };
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      const int b = tile / tiles_per_img, rem = tile - b * tiles_per_img;
      const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      const int y0 = ty * TH, x0 = tx * TW;
      if (tid < CC) s_coef[tid] = p.coef[b * CC + tid];
      named_bar_workers();   // staging + s_part of the previous tile are free, s_coef is loaded
      flush_part();
      // ---- prologue: halo tile -> GN affine -> SiLU -> fp16 (hi[, lo]) -> smem
      float sc[8], sh[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 c = s_coef[chunk * 8 + j];
        sc[j] = c.x;
        sh[j] = c.y;
      }
      const float* img = static_cast<const float*>(p.in) + int64_t(b) * p.H * p.W * CC + chunk * 8;
#pragma unroll 4
      for (int px = px0; px < HP; px += 16) {
        const int hy = px / WX, hx = px - hy * WX;
        const int sy = reflect_clamp(y0 + hy - KS / 2, p.H), sx = reflect_clamp(x0 + hx - KS / 2, p.W);
        float v[8];
#if NAF_CONV_EXP & 2
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = float(sy + sx + j);
#else
        ldg_stream8(img + (int64_t(sy) * p.W + sx) * CC, v);
#endif
#if NAF_CONV_EXP & 16
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], sc[j], sh[j]);
#else
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = silu<PASSES == 1>(fmaf(v[j], sc[j], sh[j]));
#endif
        uint4 hi, lo;
        if constexpr (PASSES == 1) {
          const __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
          const __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
          hi.x = *reinterpret_cast<const uint32_t*>(&h0);
          hi.y = *reinterpret_cast<const uint32_t*>(&h1);
          hi.z = *reinterpret_cast<const uint32_t*>(&h2);
          hi.w = *reinterpret_cast<const uint32_t*>(&h3);
        } else {
          split2_f16(v[0], v[1], hi.x, lo.x);
          split2_f16(v[2], v[3], hi.y, lo.y);
          split2_f16(v[4], v[5], hi.z, lo.z);
          split2_f16(v[6], v[7], hi.w, lo.w);
          *reinterpret_cast<uint4*>(sA + Cfg::A_PLANE + chunk * CS + px * 16) = lo;
        }
        *reinterpret_cast<uint4*>(sA + chunk * CS + px * 16) = hi;
      }
      fence_proxy_async_smem();
      fence_before_sync();          // orders this thread's TMEM reads of the previous tile
      mbar_arrive(&bar_a_full);
      // ---- epilogue
      mbar_wait(&bar_acc_full, it & 1);
      fence_after_sync();
      // Each warp owns 32 TMEM lanes (pixels) x one 64-column half.  Per round of 32 columns the
      // warp transposes through its private smem slab so that every store instruction writes four
      // 128-byte runs (4 pixels x 32 channels) instead of 32 scattered 16-byte pieces.
      const bool valid = (y0 + (row >> 3)) < p.H && (x0 + (row & 7)) < p.W;
      const int sub = lane >> 3, piece = lane & 7;
      const int ybase = y0 + (warp & 3) * 4, xbase = x0 + sub;
      float* obase = static_cast<float*>(p.out) + ((int64_t(b) * p.H + ybase) * p.W + xbase) * p.out_pix_stride + p.out_ch_off +
                     hf * 64 + piece * 4;
      float st[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) st[j] = 0.f;
#pragma unroll
      for (int rd = 0; rd < 2; ++rd) {
        uint32_t r[32];
        tmem_ld32(tmem + lane_off + hf * 64 + rd * 32, r);
        wait_ld();
        __syncwarp();   // the previous round's read-back of the slab is complete
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int c = hf * 64 + rd * 32 + j;
          float4 o;
          o.x = __uint_as_float(r[j]) + s_bias[c];
          o.y = __uint_as_float(r[j + 1]) + s_bias[c + 1];
          o.z = __uint_as_float(r[j + 2]) + s_bias[c + 2];
          o.w = __uint_as_float(r[j + 3]) + s_bias[c + 3];
          const int g = rd * 2 + (j >> 4);
          st[g * 2] += (o.x + o.y) + (o.z + o.w);
          st[g * 2 + 1] += (o.x * o.x + o.y * o.y) + (o.z * o.z + o.w * o.w);
          *reinterpret_cast<float4*>(my_stage + j * 4) = o;
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 o = *reinterpret_cast<const float4*>(warp_stage + (i * 4 + sub) * STAGE_SLOT + piece * 16);
          const int yy = ybase + (i >> 1), xx = xbase + (i & 1) * 4;
          if (yy < p.H && xx < p.W && !(NAF_CONV_EXP & 1))
            stg_stream(obase + (int64_t(i >> 1) * p.W + (i & 1) * 4) * p.out_pix_stride + rd * 32, o);
        }
      }
      if (p.part) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float v = valid ? st[j] : 0.f;
          v = warp_sum(v);
          if (lane == 0) s_part[warp][j] = v;
        }
      }
      prev_tile = tile;
    }
    named_bar_workers();
    flush_part();
  } else if (warp == MMA_WARP) {
    // ============================================================================= MMA ISSUER
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(128, CC, false, false);
      const uint32_t a_base = smem_u32(sA), w_base = smem_u32(sW);
      uint32_t n = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
        mbar_wait(&bar_a_full, it & 1);
        fence_after_sync();
#pragma unroll 1
        for (int tap = 0; tap < NT; ++tap, ++n) {
          const int slot = kResident ? 0 : int(n & 1);
#if NAF_CONV_EXP & 8
          if (n < 2) mbar_wait(&bar_w_full[slot], 0);
#else
          if (!kResident) mbar_wait(&bar_w_full[slot], (n >> 1) & 1);
          else if (n == 0) mbar_wait(&bar_w_full[0], 0);
#endif
          const int dy = tap / KS, dx = tap - dy * KS;
          const uint32_t a_tap = a_base + (dy * WX + dx) * 16;
          const uint32_t w_tap = w_base + slot * Cfg::W_GRAN;
#pragma unroll
          for (int pass = 0; pass < PASSES; ++pass) {
            const uint32_t a0 = a_tap + (pass == 1 ? Cfg::A_PLANE : 0);
            const uint32_t b0 = w_tap + (pass == 2 ? W_PLANE : 0);
#pragma unroll
            for (int kk = 0; kk < CC / 16; ++kk) {
              const uint64_t da = make_desc(a0 + kk * 2 * CS, CS, WX * 16);
              const uint64_t db = make_desc(b0 + kk * 2 * (CC * 16), CC * 16, 128);
              if (!(NAF_CONV_EXP & 4)) mma_f16_ss(tmem, da, db, idesc, (tap | pass | kk) != 0);
            }
          }
          if (!kResident && !(NAF_CONV_EXP & 8)) commit(&bar_w_free[slot]);
        }
        commit(&bar_acc_full);
      }
    }
    __syncwarp();
  } else {
    // ================================================================================= LOADER
    if (lane == 0 && blockIdx.x < total) {
      if (kResident) {
        mbar_expect_tx(&bar_w_full[0], Cfg::W_GRAN);
        bulk_load(sW, p.wpack, Cfg::W_GRAN, &bar_w_full[0]);
      } else {
        uint32_t n = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
#pragma unroll 1
          for (int tap = 0; tap < NT; ++tap, ++n) {
            const int slot = int(n & 1);
#if NAF_CONV_EXP & 8
            if (n >= 2) break;
#endif
            if (n >= 2) mbar_wait(&bar_w_free[slot], ((n >> 1) - 1) & 1);
            mbar_expect_tx(&bar_w_full[slot], Cfg::W_GRAN);
            bulk_load(sW + slot * Cfg::W_GRAN, p.wpack + size_t(tap) * 2 * W_PLANE, Cfg::W_GRAN,
                      &bar_w_full[slot]);
          }
        }
      }
    }
    __syncwarp();
  }

  fence_before_sync();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem, 128);
}

// Sum of N per-lane values over the warp with N + log2(32 / N) - 1 shuffles instead of 5 N: at every step a lane
// keeps one half of its values and hands the other half to its partner.  Afterwards the lanes with the low
// log2(32 / N) bits clear hold value `idx` = the high bits of the lane number, in v[0].
template <int N>
__device__ __forceinline__ void warp_sum_transposed(float (&v)[N], int lane) {
  int step = 16;
#pragma unroll
  for (int n = N; n > 1; n >>= 1, step >>= 1) {
    const bool up = (lane & step) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float send = up ? v[i] : v[i + n / 2];
      const float keep = up ? v[i + n / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
    }
  }
#pragma unroll
  for (; step > 0; step >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], step);
}

// Epilogue of one 16x8 tile for the pipelined kernels.  NEPI = 128: 4 warps, warp ew owns TMEM lanes = pixels
// [32 ew, 32 ew + 32) and all 128 channels; NEPI = 256: 8 warps, warp ew owns the lanes of quarter ew & 3 and the
// 64 channels of half ew >> 2.  Accumulator -> +bias -> GroupNorm partial sums -> per-warp transpose slab ->
// stores of four 128-byte (fp16: 64-byte) runs per instruction.  `tmem_acc` = accumulator base + this warp's lane offset.
template <bool OUT16, int NEPI>
__device__ __forceinline__ void ws_drain_tile(const ConvParams& p, uint32_t tmem_acc, uint64_t* bar_acc_free, int b,
                                              int y0, int x0, int tile, int ew, int lane, int et, uint8_t* my_stage,
                                              const uint8_t* warp_stage, const float* s_bias, float (*s_part)[16]) {
  constexpr int NR = NEPI == 256 ? 2 : 4;   // rounds of 32 channels per thread
  const int q = ew & 3, ch = ew >> 2;
  const int row = q * 32 + lane;
  const int sub = lane >> 3, piece = lane & 7;
  const uint64_t one2 = pack2(1.f, 1.f);
  const bool valid = (y0 + (row >> 3)) < p.H && (x0 + (row & 7)) < p.W;
  const int ybase = y0 + q * 4, xbase = x0 + sub;
  const int64_t oelem = ((int64_t(b) * p.H + ybase) * p.W + xbase) * p.out_pix_stride + p.out_ch_off + piece * 4;
  float* obase = static_cast<float*>(p.out) + oelem;          // fp32 view
  __half* obase_h = static_cast<__half*>(p.out) + oelem;      // fp16 view (OUT16)
  uint64_t st2[NR * 4];   // [group][sum | sumsq], each as an (even, odd) channel pair
#pragma unroll
  for (int j = 0; j < NR * 4; ++j) st2[j] = 0ull;
#pragma unroll
  for (int rr = 0; rr < NR; ++rr) {
    const int rd = ch * NR + rr;
    uint32_t r[32];
    tmem_ld32(tmem_acc + rd * 32, r);
    wait_ld();
    if (rr == NR - 1) {   // the accumulator is in registers: the MMA of tile it+2 may overwrite it
      fence_before_sync();
      mbar_arrive(bar_acc_free);
    }
    __syncwarp();   // the previous round's read-back of the slab is complete
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 bv = *reinterpret_cast<const float4*>(&s_bias[rd * 32 + j]);
      const uint64_t o01 = fma2(pack2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), one2, pack2(bv.x, bv.y));
      const uint64_t o23 = fma2(pack2(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])), one2, pack2(bv.z, bv.w));
      const int g = rr * 2 + (j >> 4);
      st2[g * 2] = fma2(o01, one2, st2[g * 2]);
      st2[g * 2] = fma2(o23, one2, st2[g * 2]);
      st2[g * 2 + 1] = fma2(o01, o01, st2[g * 2 + 1]);
      st2[g * 2 + 1] = fma2(o23, o23, st2[g * 2 + 1]);
      if constexpr (OUT16) {
        float a0, a1, a2, a3;
        unpack2(o01, a0, a1);
        unpack2(o23, a2, a3);
        const __half2 h01 = __floats2half2_rn(a0, a1), h23 = __floats2half2_rn(a2, a3);
        *reinterpret_cast<uint2*>(my_stage + j * 2) =
            make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
      } else {
        *reinterpret_cast<uint4*>(my_stage + j * 4) =
            make_uint4(uint32_t(o01), uint32_t(o01 >> 32), uint32_t(o23), uint32_t(o23 >> 32));
      }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int yy = ybase + (i >> 1), xx = xbase + (i & 1) * 4;
      const bool ok = yy < p.H && xx < p.W && !(NAF_CONV_EXP & 1);
      const int64_t off = (int64_t(i >> 1) * p.W + (i & 1) * 4) * p.out_pix_stride + rd * 32;
      if constexpr (OUT16) {
        // 32 fp16 channels = 64 bytes per pixel and round: four 64-byte runs per store instruction
        const uint2 o = *reinterpret_cast<const uint2*>(warp_stage + (i * 4 + sub) * STAGE_SLOT + piece * 8);
        if (ok) asm volatile("st.global.cs.v2.b32 [%0], {%1, %2};" ::"l"(obase_h + off), "r"(o.x), "r"(o.y) : "memory");
      } else {
        const float4 o = *reinterpret_cast<const float4*>(warp_stage + (i * 4 + sub) * STAGE_SLOT + piece * 16);
        if (ok) stg_stream(obase + off, o);
      }
    }
  }
  if (p.part) {
    constexpr int NV = NR * 4;
    float v[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      float v0, v1;
      unpack2(st2[j], v0, v1);
      v[j] = valid ? v0 + v1 : 0.f;
    }
    warp_sum_transposed<NV>(v, lane);
    if ((lane & (32 / NV - 1)) == 0) s_part[q][ch * NV + lane / (32 / NV)] = v[0];
    if constexpr (NEPI == 256) asm volatile("bar.sync 2, 256;" ::: "memory");
    else asm volatile("bar.sync 2, 128;" ::: "memory");
    if (et < 16)
      p.part[int64_t(tile) * 16 + et] = (s_part[0][et] + s_part[1][et]) + (s_part[2][et] + s_part[3][et]);
  }
}

// ---- pipelined variant (1 pass): producer / MMA / epilogue warps work on different tiles ---------
// One persistent CTA per SM.  A (the activated halo tile) and the TMEM accumulator are double
// buffered, so at any time the 8 producer warps convert tile t+1 while the tensor core runs tile t
// and the 4 epilogue warps drain tile t-1:
//   warps 0-7   producers : halo tile HBM -> registers (software-pipelined: the loads of the next
//                           batch of pixels, or of the next tile, are in flight while the current
//                           batch goes through GN affine -> SiLU -> fp16 -> sA[t & 1]); one bulk L2
//                           prefetch per halo row for the tile after next
//   warps 8-11  epilogue  : TMEM acc[t & 1] -> +bias -> statistics -> per-warp transpose -> stores
//   warp 12     mma       : 9 (or 1) taps x 8 k-steps of tcgen05.mma per tile
//   warp 13     loader    : weight ring, as above
namespace {
#ifndef NAF_CONV_NPROD
#define NAF_CONV_NPROD 256
#endif
#ifndef NAF_CONV_NEPI
#define NAF_CONV_NEPI 128   // epilogue threads of the conv layers: 128 (a thread drains a pixel's 128 channels) or
#endif                      // 256 (two threads per pixel, 64 channels each; measured 5 - 12 % slower: 18 warps cap the
                            // kernel at 96 registers and the layer is bound by shared-memory bandwidth, not by this group)
constexpr int WS_NPROD = NAF_CONV_NPROD, WS_NEPI = 128;   // producer threads: 256 or 384; epilogue threads of the stem
constexpr int WS_PXS = WS_NPROD / 16;                      // halo pixels converted concurrently
constexpr int WS_THREADS = WS_NPROD + WS_NEPI + 64;
constexpr int WS_MMA_WARP = (WS_NPROD + WS_NEPI) / 32;
constexpr int WC_NEPI = NAF_CONV_NEPI;                     // conv layers
constexpr int WC_THREADS = WS_NPROD + WC_NEPI + 64;
constexpr int WC_MMA_WARP = (WS_NPROD + WC_NEPI) / 32;

template <int KS>
struct WsConvCfg {
  using Base = ConvCfg<KS, 1>;
  static constexpr int A_BUF = (Base::A_PLANE + 127) / 128 * 128;
  static constexpr int STAGE = WC_NEPI * STAGE_SLOT;
  static constexpr int W_OFF = 2 * A_BUF;
#ifndef NAF_CONV_WSLOTS
#define NAF_CONV_WSLOTS 3
#endif
  static constexpr int NSLOT = Base::NT == 1 ? 1 : NAF_CONV_WSLOTS;   // weight ring depth (taps in flight)
  static constexpr int STAGE_OFF = W_OFF + NSLOT * Base::W_GRAN;
  static constexpr int SMEM = STAGE_OFF + STAGE;
  static constexpr int NIT = (Base::HP + WS_PXS - 1) / WS_PXS;   // halo pixels per producer thread
  static constexpr int BS = NIT / 2 > 4 ? NIT / 4 : NIT / 2;     // pixels per register batch
  static constexpr int NB = NIT / BS;                            // batches per tile (even)
  static_assert(NIT % BS == 0 && NB % 2 == 0, "batching");
};

struct TileCtx {
  const uint8_t* img;   // image base + this thread's channel chunk (bytes)
  const uint8_t* org;   // halo origin of the tile (dereferenced only when interior)
  int b, y0, x0;
  bool interior, valid;
};
}  // namespace

template <int KS, bool IN16, bool OUT16>
__global__ void __launch_bounds__(WC_THREADS, 1)
conv128_ws_kernel(ConvParams p) {
  using Cfg = ConvCfg<KS, 1>;
  constexpr int ES = IN16 ? 2 : 4;              // bytes per input activation element
  const uint8_t* const in_b = static_cast<const uint8_t*>(p.in);
  using Ws = WsConvCfg<KS>;
  constexpr int WX = Cfg::WX, HP = Cfg::HP, CS = Cfg::CS, NT = Cfg::NT;
  constexpr int BS = Ws::BS, NB = Ws::NB;
  constexpr bool kResident = NT == 1;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_a_full[2], bar_a_free[2], bar_acc_full[2], bar_acc_free[2], bar_w_full[4], bar_w_free[4];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_bias[CC];
  __shared__ float s_part[2][4][16];
  __shared__ int s_goff[Ws::NIT * WS_PXS];

  uint8_t* sW = smem + Ws::W_OFF;
  uint8_t* sStage = smem + Ws::STAGE_OFF;
  const int tid = threadIdx.x, lane = tid & 31;
#if NAF_CONV_ELECT_DISABLE_SHFL
  const int warp = tid >> 5;
#else
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform: the role branches are uniform branches
#endif
  const int tiles_per_img = p.tiles_y * p.tiles_x;
  const int total = p.B * tiles_per_img;

  if (warp == WC_MMA_WARP) tmem_alloc(&tmem_base_s, 256);
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_a_full[s], WS_NPROD);
      mbar_init(&bar_a_free[s], 1);
      mbar_init(&bar_acc_full[s], 1);
      mbar_init(&bar_acc_free[s], WC_NEPI);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&bar_w_full[s], 1);
      mbar_init(&bar_w_free[s], 1);
    }
    fence_mbar_init();
  }
  if (tid < CC) s_bias[tid] = p.bias ? p.bias[tid] : 0.f;
  // global element offset of halo pixel px relative to the tile's halo origin (same for every tile)
  for (int px = tid; px < Ws::NIT * WS_PXS; px += WC_THREADS) {
    const int q = px < HP ? px : HP - 1, hy = q / WX, hx = q - hy * WX;
    s_goff[px] = (hy * p.W + hx) * CC;
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  if (warp < WS_NPROD / 32) {
    // ============================================================================== PRODUCERS
    const int chunk = tid & 15, px0 = tid >> 4;
    auto ctx_of = [&](int tile) {
      TileCtx c;
      c.valid = tile < total;
      c.b = 0; c.y0 = 0; c.x0 = 0; c.interior = false; c.img = in_b; c.org = in_b;
      if (c.valid) {
        c.b = tile / tiles_per_img;
        const int rem = tile - c.b * tiles_per_img, ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        c.y0 = ty * TH;
        c.x0 = tx * TW;
        c.interior = c.y0 >= KS / 2 && c.y0 + TH + KS / 2 <= p.H && c.x0 >= KS / 2 && c.x0 + TW + KS / 2 <= p.W;
        c.img = in_b + (int64_t(c.b) * p.H * p.W * CC + chunk * 8) * ES;
        c.org = c.img + (int64_t(c.y0 - KS / 2) * p.W + (c.x0 - KS / 2)) * CC * ES;
      }
      return c;
    };
    auto load_batch = [&](const TileCtx& c, int j, float (&v)[BS][8]) {
      static_assert(NAF_CONV_ITEMPIPE || !IN16, "the batched producer variant reads fp32 activations only");
#pragma unroll
      for (int kk = 0; kk < BS; ++kk) {
        const int px = px0 + WS_PXS * (j * BS + kk);
        if (HP % WS_PXS == 0 || px < HP) {
          const float* src;
          if (c.interior) {
            src = reinterpret_cast<const float*>(c.org) + s_goff[px];
          } else {
            const int hy = px / WX, hx = px - hy * WX;
            src = reinterpret_cast<const float*>(c.img) + (int64_t(reflect_clamp(c.y0 + hy - KS / 2, p.H)) * p.W +
                           reflect_clamp(c.x0 + hx - KS / 2, p.W)) * CC;
          }
#if NAF_CONV_EXP & 2
#pragma unroll
          for (int e = 0; e < 8; ++e) v[kk][e] = float(px + e);
          (void)src;
#else
          ldg_stream8(src, v[kk]);
#endif
        }
      }
    };
    uint64_t sc2[4], sh2[4];   // GroupNorm {scale, shift} of this thread's 8 channels, as pairs
    auto convert_batch = [&](int j, float (&v)[BS][8], uint8_t* sA) {
#pragma unroll
      for (int kk = 0; kk < BS; ++kk) {
        const int k = j * BS + kk;
        if (HP % WS_PXS == 0 || px0 + WS_PXS * k < HP) {
          uint4 hi;
          uint32_t* hp = reinterpret_cast<uint32_t*>(&hi);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint64_t t2 = fma2(pack2(v[kk][2 * e], v[kk][2 * e + 1]), sc2[e], sh2[e]);   // GN affine
            hp[e] = silu2_f16(t2);
          }
          *reinterpret_cast<uint4*>(sA + k * (WS_PXS * 16)) = hi;
        }
      }
    };

#if NAF_CONV_ITEMPIPE
    // Per-item software pipeline of depth DEPTH = NIT / 2: item k is converted DEPTH items after its load
    // was issued (the batched variant gives a load only one batch of conversions to land, less than an L2
    // hit takes); the register budget is the same DEPTH x 8 floats.
    // fp16 activations: an item is 4 registers instead of 8, so the same register budget holds the loads of
    // a WHOLE tile (item k of the next tile is issued as soon as item k of this one has been converted)
    constexpr int NIT = Ws::NIT, DEPTH = IN16 ? NIT : NIT / 2;
    constexpr int IW = IN16 ? 4 : 8;             // 32-bit registers per item (8 channels)
    static_assert(NIT % DEPTH == 0, "item pipeline");
    uint32_t v[DEPTH][IW];
    auto load_item = [&](const TileCtx& c, int k, uint32_t (&dst)[IW]) {
      const int px = px0 + WS_PXS * k;
      if (HP % WS_PXS == 0 || px < HP) {
        const uint8_t* src;
        if (c.interior) {
          src = c.org + int64_t(s_goff[px]) * ES;
        } else {
          const int hy = px / WX, hx = px - hy * WX;
          src = c.img + (int64_t(reflect_clamp(c.y0 + hy - KS / 2, p.H)) * p.W +
                         reflect_clamp(c.x0 + hx - KS / 2, p.W)) * CC * ES;
        }
        if constexpr (IN16) {
          asm("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
              : "=r"(dst[0]), "=r"(dst[1]), "=r"(dst[2]), "=r"(dst[3]) : "l"(src));
        } else {
          asm("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
              : "=r"(dst[0]), "=r"(dst[1]), "=r"(dst[2]), "=r"(dst[3]), "=r"(dst[4]), "=r"(dst[5]), "=r"(dst[6]), "=r"(dst[7])
              : "l"(src));
        }
      }
    };
    auto convert_item = [&](int k, const uint32_t (&src)[IW], uint8_t* sA) {
      if (HP % WS_PXS == 0 || px0 + WS_PXS * k < HP) {
        uint4 hi;
        uint32_t* hp = reinterpret_cast<uint32_t*>(&hi);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          uint64_t x2;
          if constexpr (IN16) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&src[e]));
            x2 = pack2(f.x, f.y);
          } else {
            x2 = pack2(__uint_as_float(src[2 * e]), __uint_as_float(src[2 * e + 1]));
          }
          hp[e] = silu2_f16(fma2(x2, sc2[e], sh2[e]));
        }
        *reinterpret_cast<uint4*>(sA + k * (WS_PXS * 16)) = hi;
      }
    };
    TileCtx cur = ctx_of(blockIdx.x);
    if (cur.valid) {
#pragma unroll
      for (int k = 0; k < DEPTH; ++k) load_item(cur, k, v[k]);
    }
#else
    float va[BS][8], vb[BS][8];
    TileCtx cur = ctx_of(blockIdx.x);
    if (cur.valid) load_batch(cur, 0, va);
#endif
    int b_cur = -1, it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const TileCtx nxt = ctx_of(tile + int(gridDim.x));
      if (cur.b != b_cur) {
        b_cur = cur.b;
        const float4* cp = reinterpret_cast<const float4*>(p.coef + cur.b * CC + chunk * 8);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 c = __ldg(cp + j);   // {scale, shift} of channels 2j, 2j+1
          sc2[j] = pack2(c.x, c.z);
          sh2[j] = pack2(c.y, c.w);
        }
      }
      if (tid < Cfg::HY) {   // L2 prefetch of the tile after next: one bulk prefetch per halo row
        const int t2 = tile + NAF_CONV_PF_DIST * int(gridDim.x);
        if (t2 < total) {
          const int b2 = t2 / tiles_per_img, rem2 = t2 - b2 * tiles_per_img;
          const int ty2 = rem2 / p.tiles_x, tx2 = rem2 - ty2 * p.tiles_x;
          const int yy = reflect_clamp(ty2 * TH - KS / 2 + tid, p.H);
          const int span = WX < p.W ? WX : p.W;
          int xs = tx2 * TW - KS / 2;
          xs = xs < 0 ? 0 : (xs > p.W - span ? p.W - span : xs);
          bulk_prefetch_l2(in_b + ((int64_t(b2) * p.H + yy) * p.W + xs) * CC * ES, uint32_t(span) * CC * ES);
        }
      }
      if (it >= 2) mbar_wait(&bar_a_free[buf], ((it >> 1) - 1) & 1);
      uint8_t* sA = smem + buf * Ws::A_BUF + chunk * CS + px0 * 16;
#if NAF_CONV_ITEMPIPE
#pragma unroll
      for (int k = 0; k < NIT; ++k) {
        convert_item(k, v[k % DEPTH], sA);
        if (k + DEPTH < NIT) load_item(cur, k + DEPTH, v[k % DEPTH]);
        else if (nxt.valid) load_item(nxt, k + DEPTH - NIT, v[k % DEPTH]);
      }
#else
#pragma unroll
      for (int j = 0; j < NB; j += 2) {
        load_batch(cur, j + 1, vb);
        convert_batch(j, va, sA);
        if (j + 2 < NB) load_batch(cur, j + 2, va);
        else if (nxt.valid) load_batch(nxt, 0, va);
        convert_batch(j + 1, vb, sA);
      }
#endif
      fence_proxy_async_smem();
      mbar_arrive(&bar_a_full[buf]);
      cur = nxt;
    }
  } else if (warp < WC_MMA_WARP) {
    // =============================================================================== EPILOGUE
    // warp ew owns TMEM lanes (pixels) [32 (ew & 3), +32) and, with 8 epilogue warps, the 64 channels of half
    // ew >> 2; a thread drains its channels in rounds of 32 columns, each round transposed through the warp's
    // private smem slab so that every store instruction writes four 128-byte runs (4 pixels x 32 channels).
    const int ew = warp - WS_NPROD / 32, et = tid - WS_NPROD;
    const uint32_t lane_off = uint32_t((ew & 3) * 32) << 16;
    uint8_t* my_stage = sStage + et * STAGE_SLOT;
    const uint8_t* warp_stage = sStage + ew * 32 * STAGE_SLOT;
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const int b = tile / tiles_per_img, rem = tile - b * tiles_per_img;
      const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      const int y0 = ty * TH, x0 = tx * TW;
      mbar_wait(&bar_acc_full[buf], (it >> 1) & 1);
      fence_after_sync();
      ws_drain_tile<OUT16, WC_NEPI>(p, tmem + lane_off + buf * CC, &bar_acc_free[buf], b, y0, x0, tile, ew, lane, et, my_stage,
                           warp_stage, s_bias, s_part[buf]);
    }
  } else if (warp == WC_MMA_WARP) {
    // ============================================================================= MMA ISSUER
#ifndef NAF_CONV_ELECT
#define NAF_CONV_ELECT 1   // 1: issue under elect.sync (descriptors in uniform registers), 0: under lane == 0
#endif
    __syncwarp();
    if (NAF_CONV_ELECT ? elect_one_sync() : lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(128, CC, false, false);
      const uint32_t w_base = smem_u32(sW);
      constexpr int NS = Ws::NSLOT;
      uint32_t slot = 0, wphase = 0;   // ring position of the next tap and its mbarrier phase
      uint32_t n = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(&bar_a_full[buf], (it >> 1) & 1);
        if (it >= 2) mbar_wait(&bar_acc_free[buf], ((it >> 1) - 1) & 1);
        fence_after_sync();
        const uint32_t a_base = smem_u32(smem + buf * Ws::A_BUF);
        const uint32_t tD = tmem + buf * CC;
#pragma unroll 1
        for (int tap = 0; tap < NT; ++tap, ++n) {
          if (!kResident) mbar_wait(&bar_w_full[slot], wphase);
          else if (n == 0) mbar_wait(&bar_w_full[0], 0);
          const int dy = tap / KS, dx = tap - dy * KS;
#if NAF_CONV_EXP & 32
          const uint32_t a0 = a_base + 0 * (dy + dx);   // every tap reads the unshifted, 128-byte aligned tile
          constexpr int SBO_A = 128;
#else
          const uint32_t a0 = a_base + (dy * WX + dx) * 16;
          constexpr int SBO_A = WX * 16;
#endif
          const uint32_t b0 = w_base + slot * Cfg::W_GRAN;
#pragma unroll
          for (int kk = 0; kk < CC / 16; ++kk) {
            const uint64_t da = make_desc(a0 + kk * 2 * CS, CS, SBO_A);
            const uint64_t db = make_desc(b0 + kk * 2 * (CC * 16), CC * 16, 128);
            if (!(NAF_CONV_EXP & 4)) mma_f16_ss(tD, da, db, idesc, (tap | kk) != 0);
          }
          if (!kResident) {
            commit(&bar_w_free[slot]);
            if (++slot == NS) { slot = 0; wphase ^= 1; }
          }
        }
        commit(&bar_a_free[buf]);
        commit(&bar_acc_full[buf]);
      }
    }
    __syncwarp();
  } else {
    // ================================================================================= LOADER
    if (lane == 0) {
      if (kResident) {
        mbar_expect_tx(&bar_w_full[0], Cfg::W_GRAN);
        bulk_load(sW, p.wpack, Cfg::W_GRAN, &bar_w_full[0]);
      } else {
        constexpr int NS = Ws::NSLOT;
        uint32_t slot = 0, fphase = 1;   // fphase: parity of the PREVIOUS use's "free" completion
        bool first_round = true;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
#pragma unroll 1
          for (int tap = 0; tap < NT; ++tap) {
            if (!first_round) mbar_wait(&bar_w_free[slot], fphase);
            mbar_expect_tx(&bar_w_full[slot], Cfg::W_GRAN);
            bulk_load(sW + slot * Cfg::W_GRAN, p.wpack + size_t(tap) * 2 * W_PLANE, Cfg::W_GRAN,
                      &bar_w_full[slot]);
            if (++slot == NS) { slot = 0; fphase ^= 1; first_round = false; }
          }
        }
      }
    }
    __syncwarp();
  }

  fence_before_sync();
  __syncthreads();
  if (warp == WC_MMA_WARP) tmem_dealloc(tmem, 256);
}

// ---- stem on the tensor core (1 pass): Conv2d(3 -> 128, KS, reflect) as a 128 x 128 x {16, 32} GEMM ----
// The SIMT stem below is FMA bound (0.40 ms for the 3x3 stem at C2, 3x the time its 822 MB of output
// needs).  Here the producers build the im2col row of every pixel directly in the K-major operand layout
// (K index = tap*3 + channel, 27 of 32 used; 3 of 16 for the 1x1 stem), the weights sit in shared memory
// for the whole kernel, and the epilogue is the one of the pipelined conv kernel (bias, GroupNorm
// partial sums, transpose slab, 128-byte-run stores).  Operands are rounded to fp16 (the TF32 class);
// the strict fp32 class keeps the SIMT kernel.
#ifndef NAF_STEM_CTAS
#define NAF_STEM_CTAS 2   // CTAs per SM of the tensor-core stem (43 KB smem, 256 TMEM columns each)
#endif
template <int KS, bool OUT16>
__global__ void __launch_bounds__(WS_THREADS, NAF_STEM_CTAS)
stem_tc_kernel(ConvParams p, const float* __restrict__ image, int64_t sb, int64_t sc, int64_t sy, int64_t sx,
               const float* __restrict__ w) {
  constexpr int NK = 3 * KS * KS;              // 27 or 3
  constexpr int KDIM = NK > 16 ? 32 : 16;      // MMA K, zero padded
  constexpr int NCH = KDIM / 8;                // 16-byte chunks per row
  constexpr int A_BUF = NCH * 128 * 16;        // one operand tile
  constexpr int STEM_W_OFF = 2 * A_BUF, STEM_STAGE_OFF = STEM_W_OFF + NCH * CC * 16;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_a_full[2], bar_a_free[2], bar_acc_full[2], bar_acc_free[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_bias[CC];
  __shared__ float s_part[2][4][16];

  uint8_t* sW = smem + STEM_W_OFF;
  uint8_t* sStage = smem + STEM_STAGE_OFF;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles_per_img = p.tiles_y * p.tiles_x;
  const int total = p.B * tiles_per_img;

  if (warp == WS_MMA_WARP) tmem_alloc(&tmem_base_s, 256);
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_a_full[s], WS_NPROD);
      mbar_init(&bar_a_free[s], 1);
      mbar_init(&bar_acc_full[s], 1);
      mbar_init(&bar_acc_free[s], WS_NEPI);
    }
    fence_mbar_init();
  }
  if (tid < CC) s_bias[tid] = p.bias ? p.bias[tid] : 0.f;
  // weights (128, 3, KS, KS) fp32 -> K-major fp16 [chunk][n][8], k = tap*3 + ci, zero padded
  for (int i = tid; i < CC * KDIM; i += WS_THREADS) {
    const int n = i / KDIM, k = i - n * KDIM;
    float v = 0.f;
    if (k < NK) {
      const int tap = k / 3, ci = k - tap * 3;
      v = w[(n * 3 + ci) * KS * KS + tap];
    }
    reinterpret_cast<__half*>(sW)[((k >> 3) * CC + n) * 8 + (k & 7)] = __float2half_rn(v);
  }
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  if (warp < WS_NPROD / 32) {
    // ============================================================================== PRODUCERS
    // thread = (pixel pp of the tile, half kh of the K range): 16 K values = two 16-byte chunks
    const int pp = tid & 127, kh = tid >> 7;
    const int py = pp >> 3, px = pp & 7;
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const int b = tile / tiles_per_img, rem = tile - b * tiles_per_img;
      const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      const int y = ty * TH + py, x = tx * TW + px;
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = 0.f;
      if (kh * 16 < NK) {
        const float* img = image + int64_t(b) * sb;
        int64_t roff[KS], coff[KS];
#pragma unroll
        for (int d = 0; d < KS; ++d) {
          roff[d] = int64_t(reflect_clamp(y + d - KS / 2, p.H)) * sy;
          coff[d] = int64_t(reflect_clamp(x + d - KS / 2, p.W)) * sx;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          // k = kh*16 + i; both halves are compiled, the runtime kh selects
          const int k0 = i, k1 = 16 + i;
          if (kh == 0 && k0 < NK) v[i] = __ldg(img + (k0 % 3) * sc + roff[(k0 / 3) / KS] + coff[(k0 / 3) % KS]);
          if (kh == 1 && k1 < NK) v[i] = __ldg(img + (k1 % 3) * sc + roff[(k1 / 3) / KS] + coff[(k1 / 3) % KS]);
        }
      }
      if (it >= 2) mbar_wait(&bar_a_free[buf], ((it >> 1) - 1) & 1);
      if (kh * 2 < NCH) {
        uint8_t* sA = smem + buf * A_BUF + (kh * 2) * (128 * 16) + pp * 16;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint4 hv;
          uint32_t* hp = reinterpret_cast<uint32_t*>(&hv);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const __half2 h = __floats2half2_rn(v[c * 8 + 2 * e], v[c * 8 + 2 * e + 1]);
            hp[e] = *reinterpret_cast<const uint32_t*>(&h);
          }
          *reinterpret_cast<uint4*>(sA + c * (128 * 16)) = hv;
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&bar_a_full[buf]);
    }
  } else if (warp < WS_MMA_WARP) {
    // =============================================================================== EPILOGUE
    const int ew = warp - WS_NPROD / 32, et = tid - WS_NPROD;
    const uint32_t lane_off = uint32_t(ew * 32) << 16;
    uint8_t* my_stage = sStage + et * STAGE_SLOT;
    const uint8_t* warp_stage = sStage + ew * 32 * STAGE_SLOT;
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const int b = tile / tiles_per_img, rem = tile - b * tiles_per_img;
      const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      mbar_wait(&bar_acc_full[buf], (it >> 1) & 1);
      fence_after_sync();
      ws_drain_tile<OUT16, WS_NEPI>(p, tmem + lane_off + buf * CC, &bar_acc_free[buf], b, ty * TH, tx * TW, tile, ew, lane, et,
                           my_stage, warp_stage, s_bias, s_part[buf]);
    }
  } else if (warp == WS_MMA_WARP) {
    // ============================================================================= MMA ISSUER
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(128, CC, false, false);
      const uint32_t w_base = smem_u32(sW);
      int it = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(&bar_a_full[buf], (it >> 1) & 1);
        if (it >= 2) mbar_wait(&bar_acc_free[buf], ((it >> 1) - 1) & 1);
        fence_after_sync();
        const uint32_t a_base = smem_u32(smem + buf * A_BUF);
#pragma unroll
        for (int kk = 0; kk < KDIM / 16; ++kk) {
          const uint64_t da = make_desc(a_base + kk * 2 * (128 * 16), 128 * 16, 128);
          const uint64_t db = make_desc(w_base + kk * 2 * (CC * 16), CC * 16, 128);
          mma_f16_ss(tmem + buf * CC, da, db, idesc, kk != 0);
        }
        commit(&bar_a_free[buf]);
        commit(&bar_acc_full[buf]);
      }
    }
    __syncwarp();
  }

  fence_before_sync();
  __syncthreads();
  if (warp == WS_MMA_WARP) tmem_dealloc(tmem, 256);
}

// ---- stem: Conv2d(3 -> 128, KS, reflect) + bias, pixel-major output, GroupNorm partial sums ------
// CTAs of 128 threads walk 16x8 tiles of one image; thread = 4 output channels (weights in registers,
// loaded once per CTA) x 32 pixels per tile; the 3-channel halo tile sits in shared memory and is
// read with warp-wide broadcasts.
template <int KS, bool OUT16>
__global__ void __launch_bounds__(128)
stem_conv_kernel(const float* __restrict__ image, int64_t sb, int64_t sc, int64_t sy, int64_t sx,
                 const float* __restrict__ w, const float* __restrict__ bias, void* __restrict__ out_v,
                 float* __restrict__ part, int H, int W, int tiles_x, int tiles) {
  float* const out = static_cast<float*>(out_v);
  __half* const out_h = static_cast<__half*>(out_v);
  constexpr int HY = TH + KS - 1, WX = TW + KS - 1, NK = 3 * KS * KS;
  __shared__ __align__(16) float s_in[HY * WX][4];   // [halo pixel][channel (3) + pad]
  __shared__ float s_red[4][8][2];
  const int tid = threadIdx.x, q = tid & 31, slot = tid >> 5;
  const int b = blockIdx.y;
  // weights (128, 3, KS, KS): wr[o][ (dy*KS+dx)*3 + ci ]
  float wr[4][NK];
#pragma unroll
  for (int o = 0; o < 4; ++o)
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
      for (int t = 0; t < KS * KS; ++t) wr[o][t * 3 + ci] = w[((4 * q + o) * 3 + ci) * KS * KS + t];
  const float4 bv = bias ? *reinterpret_cast<const float4*>(bias + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int y0 = ty * TH, x0 = tx * TW;
    __syncthreads();   // the previous tile's readers of s_in / s_red are done
    for (int i = tid; i < HY * WX; i += 128) {
      const int hy = i / WX, hx = i - hy * WX;
      const int yy = reflect_clamp(y0 + hy - KS / 2, H), xx = reflect_clamp(x0 + hx - KS / 2, W);
      const float* src = image + b * sb + yy * sy + xx * sx;
      s_in[i][0] = src[0];
      s_in[i][1] = src[sc];
      s_in[i][2] = src[2 * sc];
      s_in[i][3] = 0.f;
    }
    __syncthreads();
    float s = 0.f, ss = 0.f;
#pragma unroll 2
    for (int pi = slot; pi < TH * TW; pi += 4) {
      const int py = pi >> 3, pxx = pi & 7;
      float4 acc = bv;
#pragma unroll
      for (int dy = 0; dy < KS; ++dy)
#pragma unroll
        for (int dx = 0; dx < KS; ++dx) {
          const float4 in = *reinterpret_cast<const float4*>(&s_in[(py + dy) * WX + pxx + dx][0]);
          const int t = (dy * KS + dx) * 3;
          acc.x = fmaf(in.x, wr[0][t], fmaf(in.y, wr[0][t + 1], fmaf(in.z, wr[0][t + 2], acc.x)));
          acc.y = fmaf(in.x, wr[1][t], fmaf(in.y, wr[1][t + 1], fmaf(in.z, wr[1][t + 2], acc.y)));
          acc.z = fmaf(in.x, wr[2][t], fmaf(in.y, wr[2][t + 1], fmaf(in.z, wr[2][t + 2], acc.z)));
          acc.w = fmaf(in.x, wr[3][t], fmaf(in.y, wr[3][t + 1], fmaf(in.z, wr[3][t + 2], acc.w)));
        }
      const int y = y0 + py, x = x0 + pxx;
      if (y < H && x < W) {
        if constexpr (OUT16) {
          const __half2 h01 = __floats2half2_rn(acc.x, acc.y), h23 = __floats2half2_rn(acc.z, acc.w);
          asm volatile("st.global.cs.v2.b32 [%0], {%1, %2};" ::"l"(out_h + ((int64_t(b) * H + y) * W + x) * CC + 4 * q),
                       "r"(*reinterpret_cast<const uint32_t*>(&h01)), "r"(*reinterpret_cast<const uint32_t*>(&h23))
                       : "memory");
        } else {
          stg_stream(out + ((int64_t(b) * H + y) * W + x) * CC + 4 * q, acc);
        }
        s += (acc.x + acc.y) + (acc.z + acc.w);
        ss += (acc.x * acc.x + acc.y * acc.y) + (acc.z * acc.z + acc.w * acc.w);
      }
    }
    if (part) {
      // 4 adjacent lanes share a group (16 channels)
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      ss += __shfl_xor_sync(0xffffffffu, ss, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      ss += __shfl_xor_sync(0xffffffffu, ss, 2);
      if ((q & 3) == 0) {
        s_red[slot][q >> 2][0] = s;
        s_red[slot][q >> 2][1] = ss;
      }
      __syncthreads();
      if (tid < 16) {
        const int g = tid >> 1, which = tid & 1;
        const float v = (s_red[0][g][which] + s_red[1][g][which]) + (s_red[2][g][which] + s_red[3][g][which]);
        part[(int64_t(b) * tiles + tile) * 16 + tid] = v;
      }
    }
  }
}

// ---- GroupNorm coefficients from the per-tile partial sums (deterministic order, double) --------
// coef[b][c] = { rstd*gamma[c], beta[c] - mean*rstd*gamma[c] }   (biased variance, torch GroupNorm)
// One CTA per (image, group): 256 threads stride over the tiles, then a fixed-order tree reduction.
__global__ void __launch_bounds__(256)
gn_coef_kernel(const float* __restrict__ part, const float* __restrict__ gamma,
               const float* __restrict__ beta, float2* __restrict__ coef, int tiles, double count, float eps) {
  __shared__ double s_s[8], s_ss[8];
  __shared__ float s_mean, s_rstd;
  const int b = blockIdx.x, g = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double ds = 0.0, dss = 0.0;
  const float2* base = reinterpret_cast<const float2*>(part + int64_t(b) * tiles * 16 + g * 2);
  for (int i = threadIdx.x; i < tiles; i += 256) {
    const float2 v = __ldg(base + int64_t(i) * 8);
    ds += double(v.x);
    dss += double(v.y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ds += __shfl_xor_sync(0xffffffffu, ds, o);
    dss += __shfl_xor_sync(0xffffffffu, dss, o);
  }
  if (lane == 0) {
    s_s[warp] = ds;
    s_ss[warp] = dss;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tss = 0.0;
    for (int w = 0; w < 8; ++w) {
      ts += s_s[w];
      tss += s_ss[w];
    }
    const double mean = ts / count;
    double var = tss / count - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    s_mean = float(mean);
    s_rstd = float(1.0 / sqrt(var + double(eps)));
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    const int c = g * 16 + threadIdx.x;
    const float a = s_rstd * gamma[c];
    coef[b * CC + c] = make_float2(a, beta[c] - s_mean * a);
  }
}

// ---- weight packing: (128,128,KS,KS) fp32 -> [tap][hi|lo][chunk][n][8] fp16 ----------------------
__global__ void pack_conv_w_kernel(const float* __restrict__ w, __half* __restrict__ out, int KS) {
  const int NT = KS * KS;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NT * CC * CC) return;
  const int j = i & 7, n = (i >> 3) & 127, c = (i >> 10) & 15, tap = i >> 14;
  const float v = w[(int64_t(n) * CC + c * 8 + j) * NT + tap];
  const __half hi = __float2half_rn(v);
  const __half lo = __float2half_rn(v - __half2float(hi));
  const int64_t o = (int64_t(tap) * 2 * NCHUNK + c) * (CC * 8) + n * 8 + j;
  out[o] = hi;
  out[o + int64_t(NCHUNK) * CC * 8] = lo;
}

// ------------------------------------------------------------------------------------ host side
namespace {

template <int KS, bool IN16, bool OUT16>
int launch_conv_ws(const ConvParams& p, cudaStream_t st) {
  using Ws = WsConvCfg<KS>;
  auto kern = conv128_ws_kernel<KS, IN16, OUT16>;
  cudaError_t e = ensure_dyn_smem(kern, Ws::SMEM);
  if (e != cudaSuccess) return fail(NAF_ERR_CUDA, "enc_conv: smem opt-in failed: %s", cudaGetErrorString(e));
  const int sms = device_sm_count();
  const int64_t total = int64_t(p.B) * p.tiles_y * p.tiles_x;
  const int grid = int(total < sms ? total : sms);
  kern<<<grid, WC_THREADS, Ws::SMEM, st>>>(p);
  return check_launch("enc_conv(ws)");
}

template <int KS, int PASSES>
int launch_conv(const ConvParams& p, cudaStream_t st) {
  using Cfg = ConvCfg<KS, PASSES>;
  auto kern = conv128_tc_kernel<KS, PASSES>;
  cudaError_t e = ensure_dyn_smem(kern, Cfg::SMEM);
  if (e != cudaSuccess) return fail(NAF_ERR_CUDA, "enc_conv: smem opt-in failed: %s", cudaGetErrorString(e));
  const int sms = device_sm_count();
  const int64_t total = int64_t(p.B) * p.tiles_y * p.tiles_x;
  const int64_t cap = int64_t(sms) * Cfg::CTAS_PER_SM;
  const int grid = int(total < cap ? total : cap);
  kern<<<grid, NTHREADS, Cfg::SMEM, st>>>(p);
  return check_launch("enc_conv");
}

template <int KS>
int launch_conv_ws_dt(const ConvParams& p, int in_dtype, int out_dtype, cudaStream_t st) {
  const bool i16 = in_dtype == NAF_DTYPE_F16, o16 = out_dtype == NAF_DTYPE_F16;
  if (i16 && o16) return launch_conv_ws<KS, true, true>(p, st);
  if (i16) return launch_conv_ws<KS, true, false>(p, st);
  if (o16) return launch_conv_ws<KS, false, true>(p, st);
  return launch_conv_ws<KS, false, false>(p, st);
}

}  // namespace

int launch_enc_conv(const void* in, const float* coef, const void* wpack, const float* bias, void* out,
                    int64_t out_pix_stride, int out_ch_off, float* part, int B, int H, int W, int KS,
                    int passes, int in_dtype, int out_dtype, cudaStream_t st) {
  NAF_REQUIRE(KS == 1 || KS == 3, NAF_ERR_UNSUPPORTED, "enc_conv: kernel size %d (1 or 3)", KS);
  NAF_REQUIRE(passes == 1 || passes == 3, NAF_ERR_UNSUPPORTED, "enc_conv: passes must be 1 or 3");
  for (int dt : {in_dtype, out_dtype})
    NAF_REQUIRE(dt == NAF_DTYPE_F32 || dt == NAF_DTYPE_F16, NAF_ERR_UNSUPPORTED, "enc_conv: activation dtype %d (0 = f32, 2 = f16)", dt);
  NAF_REQUIRE(passes == 1 || (in_dtype == NAF_DTYPE_F32 && out_dtype == NAF_DTYPE_F32), NAF_ERR_UNSUPPORTED,
              "enc_conv: fp16 activations belong to the 1-pass (TF32) class only");
  NAF_REQUIRE(KS == 1 || (H >= 2 && W >= 2), NAF_ERR_BAD_SHAPE, "enc_conv: reflect padding needs H, W >= 2");
  NAF_REQUIRE(aligned32(in) && aligned16(wpack) && aligned16(out) && aligned16(coef) &&
                  out_pix_stride % 4 == 0 && out_ch_off % 4 == 0 && out_pix_stride >= out_ch_off + CC,
              NAF_ERR_ALIGNMENT, "enc_conv: alignment / output slab");
  NAF_REQUIRE(int64_t(B) * ((H + TH - 1) / TH) * ((W + TW - 1) / TW) < (int64_t(1) << 27), NAF_ERR_UNSUPPORTED,
              "enc_conv: too many tiles");
  ConvParams p;
  p.in = in;
  p.coef = reinterpret_cast<const float2*>(coef);
  p.wpack = static_cast<const uint8_t*>(wpack);
  p.bias = bias;
  p.out = out;
  p.part = part;
  p.out_pix_stride = out_pix_stride;
  p.out_ch_off = out_ch_off;
  p.B = B;
  p.H = H;
  p.W = W;
  p.tiles_y = (H + TH - 1) / TH;
  p.tiles_x = (W + TW - 1) / TW;
  // passes: 1 = 1-pass pipelined kernel (TF32 class), 3 = split-fp16 kernel (fp32 class)
  if (KS == 1) return passes == 1 ? launch_conv_ws_dt<1>(p, in_dtype, out_dtype, st) : launch_conv<1, 3>(p, st);
  return passes == 1 ? launch_conv_ws_dt<3>(p, in_dtype, out_dtype, st) : launch_conv<3, 3>(p, st);
}

int launch_enc_stem(const float* image, int64_t sb, int64_t sc, int64_t sy, int64_t sx, const float* w,
                    const float* bias, void* out, float* part, int B, int H, int W, int KS, int out_dtype,
                    cudaStream_t st) {
  NAF_REQUIRE(KS == 1 || KS == 3, NAF_ERR_UNSUPPORTED, "enc_stem: kernel size %d (1 or 3)", KS);
  NAF_REQUIRE(out_dtype == NAF_DTYPE_F32 || out_dtype == NAF_DTYPE_F16, NAF_ERR_UNSUPPORTED, "enc_stem: out dtype %d", out_dtype);
  NAF_REQUIRE(KS == 1 || (H >= 2 && W >= 2), NAF_ERR_BAD_SHAPE, "enc_stem: reflect padding needs H, W >= 2");
  NAF_REQUIRE(aligned16(out) && (!bias || aligned16(bias)), NAF_ERR_ALIGNMENT, "enc_stem: 16-byte alignment");
  NAF_REQUIRE(B <= 65535, NAF_ERR_UNSUPPORTED, "enc_stem: batch too large");
  const int tiles_y = (H + TH - 1) / TH, tiles_x = (W + TW - 1) / TW, tiles = tiles_y * tiles_x;
  const int sms = device_sm_count();
  int per_img = (sms * 4 + B - 1) / B;   // ~4 CTAs of 128 threads per SM over the whole batch
  if (per_img > tiles) per_img = tiles;
  const dim3 grid(unsigned(per_img), unsigned(B), 1u);
  auto go = [&](auto kern) {
    prefer_max_shared(kern);
    kern<<<grid, 128, 0, st>>>(image, sb, sc, sy, sx, w, bias, out, part, H, W, tiles_x, tiles);
  };
  const bool o16 = out_dtype == NAF_DTYPE_F16;
  if (KS == 1) { if (o16) go(stem_conv_kernel<1, true>); else go(stem_conv_kernel<1, false>); }
  else { if (o16) go(stem_conv_kernel<3, true>); else go(stem_conv_kernel<3, false>); }
  return check_launch("enc_stem");
}

int launch_enc_stem_tc(const float* image, int64_t sb, int64_t sc, int64_t sy, int64_t sx, const float* w,
                       const float* bias, void* out, float* part, int B, int H, int W, int KS, int out_dtype,
                       cudaStream_t st) {
  NAF_REQUIRE(KS == 1 || KS == 3, NAF_ERR_UNSUPPORTED, "enc_stem_tc: kernel size %d (1 or 3)", KS);
  NAF_REQUIRE(out_dtype == NAF_DTYPE_F32 || out_dtype == NAF_DTYPE_F16, NAF_ERR_UNSUPPORTED, "enc_stem_tc: out dtype %d", out_dtype);
  NAF_REQUIRE(KS == 1 || (H >= 2 && W >= 2), NAF_ERR_BAD_SHAPE, "enc_stem_tc: reflect padding needs H, W >= 2");
  NAF_REQUIRE(aligned16(out) && (!bias || aligned16(bias)), NAF_ERR_ALIGNMENT, "enc_stem_tc: 16-byte alignment");
  ConvParams p;
  p.in = nullptr;
  p.coef = nullptr;
  p.wpack = nullptr;
  p.bias = bias;
  p.out = out;
  p.part = part;
  p.out_pix_stride = CC;
  p.out_ch_off = 0;
  p.B = B;
  p.H = H;
  p.W = W;
  p.tiles_y = (H + TH - 1) / TH;
  p.tiles_x = (W + TW - 1) / TW;
  NAF_REQUIRE(int64_t(B) * p.tiles_y * p.tiles_x < (int64_t(1) << 27), NAF_ERR_UNSUPPORTED, "enc_stem_tc: too many tiles");
  const int sms = device_sm_count();
  const int64_t total = int64_t(B) * p.tiles_y * p.tiles_x;
  const int64_t cap = int64_t(sms) * NAF_STEM_CTAS;
  const int grid = int(total < cap ? total : cap);
  const int nch = KS == 3 ? 4 : 2;
  const int smem = 2 * nch * 128 * 16 + nch * CC * 16 + WS_NEPI * STAGE_SLOT;
  cudaError_t e = cudaSuccess;
  auto go = [&](auto kern) {
    prefer_max_shared(kern);
    e = ensure_dyn_smem(kern, smem);
    if (e == cudaSuccess) kern<<<grid, WS_THREADS, smem, st>>>(p, image, sb, sc, sy, sx, w);
  };
  const bool o16 = out_dtype == NAF_DTYPE_F16;
  if (KS == 3) { if (o16) go(stem_tc_kernel<3, true>); else go(stem_tc_kernel<3, false>); }
  else { if (o16) go(stem_tc_kernel<1, true>); else go(stem_tc_kernel<1, false>); }
  if (e != cudaSuccess) return fail(NAF_ERR_CUDA, "enc_stem_tc: smem opt-in failed: %s", cudaGetErrorString(e));
  return check_launch("enc_stem_tc");
}

int launch_enc_gn_coef(const float* part, const float* gamma, const float* beta, float* coef, int B, int H,
                       int W, float eps, cudaStream_t st) {
  NAF_REQUIRE(aligned16(coef), NAF_ERR_ALIGNMENT, "enc_gn_coef: 16-byte alignment");
  const int tiles = ((H + TH - 1) / TH) * ((W + TW - 1) / TW);
  NAF_REQUIRE(B <= 65535, NAF_ERR_UNSUPPORTED, "enc_gn_coef: batch too large");
  prefer_max_shared(gn_coef_kernel);
  gn_coef_kernel<<<dim3(unsigned(B), 8u, 1u), 256, 0, st>>>(part, gamma, beta, reinterpret_cast<float2*>(coef), tiles,
                                                          double(H) * W * 16.0, eps);
  return check_launch("enc_gn_coef");
}

int launch_enc_conv_pack(const float* w, void* packed, int KS, cudaStream_t st) {
  NAF_REQUIRE(KS == 1 || KS == 3, NAF_ERR_UNSUPPORTED, "enc_conv_pack: kernel size %d (1 or 3)", KS);
  const int n = KS * KS * CC * CC;
  pack_conv_w_kernel<<<(n + 255) / 256, 256, 0, st>>>(w, static_cast<__half*>(packed), KS);
  return check_launch("enc_conv_pack");
}

}  // namespace naf
