// Guidance-encoder glue kernels: GroupNorm statistics and the fused
// bias + GroupNorm + SiLU (+ reflect pad) pass, on pixel-major (NHWC) activations.
//
// The reference's EncBlock (src/layers/convolutions.py:55-67) runs GroupNorm -> SiLU -> Conv2d
// (reflect padding) as ~6 ATen kernels per conv (a 64-CTA Welford reduction that takes 3 ms per
// call at 8x128x448x448, a normalise pass, a SiLU pass, a reflection_pad2d pass, a bias add and two
// cuDNN layout conversions).  Here one HBM-bound reduction produces the group moments and ONE
// elementwise pass applies  silu((y + conv_bias - mean) * rstd * gamma + beta)  and writes the
// result directly into the reflect-padded NHWC tensor the next convolution consumes.
#include "naf_common.cuh"

namespace naf {

// ---- moments: sums[b][g] = { sum(y+bias), sum((y+bias)^2) } over H*W*(C/G) elements, in double --
// block = 256 threads = 8 pixel slots x 32 channel quads (C == 128), or generic C via loop.
__global__ void __launch_bounds__(256)
gn_stats_kernel(const float* __restrict__ y, const float* __restrict__ bias, double* __restrict__ sums,
                int64_t HW, int C, int G, int chunks) {
  extern __shared__ float red[];  // [256][2]
  const int b = blockIdx.y;
  const int quads = C >> 2;                 // float4 per pixel
  const int slots = 256 / quads;            // pixels processed concurrently
  const int q = threadIdx.x % quads, slot = threadIdx.x / quads;
  const bool active = slot < slots;
  const int64_t per = (HW + chunks - 1) / chunks;
  const int64_t p0 = int64_t(blockIdx.x) * per;
  const int64_t p1 = p0 + per < HW ? p0 + per : HW;
  float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bias && active) bv = *reinterpret_cast<const float4*>(bias + 4 * q);
  float s = 0.f, ss = 0.f;
  if (active) {
    const float* base = y + (int64_t(b) * HW) * C + 4 * q;
#pragma unroll 4
    for (int64_t p = p0 + slot; p < p1; p += slots) {
      const float4 v = ldg_stream(base + p * C);
      const float a0 = v.x + bv.x, a1 = v.y + bv.y, a2 = v.z + bv.z, a3 = v.w + bv.w;
      s += (a0 + a1) + (a2 + a3);
      ss += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
    }
  }
  red[threadIdx.x * 2] = s;
  red[threadIdx.x * 2 + 1] = ss;
  __syncthreads();
  // one thread per group gathers the quads of its group over all slots
  const int qpg = quads / G;  // quads per group
  if (threadIdx.x < G) {
    double ds = 0.0, dss = 0.0;
    for (int sl = 0; sl < slots; ++sl)
      for (int j = 0; j < qpg; ++j) {
        const int t = sl * quads + threadIdx.x * qpg + j;
        ds += double(red[t * 2]);
        dss += double(red[t * 2 + 1]);
      }
    atomicAdd(&sums[(int64_t(b) * G + threadIdx.x) * 2], ds);
    atomicAdd(&sums[(int64_t(b) * G + threadIdx.x) * 2 + 1], dss);
  }
}

// ---- apply: out = silu((y + bias - mean) * rstd * gamma + beta), optional reflect pad of 1 ------
__global__ void __launch_bounds__(256)
gn_silu_apply_kernel(const float* __restrict__ y, const float* __restrict__ bias,
                     const float* __restrict__ gamma, const float* __restrict__ beta,
                     const double* __restrict__ sums, float* __restrict__ out, int H, int W, int C,
                     int G, float eps, int pad) {
  __shared__ float s_mean[64], s_rstd[64];
  const int b = blockIdx.y;
  const int quads = C >> 2;
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  if (threadIdx.x < G) {
    const double cnt = double(H) * W * (C / G);
    const double su = sums[(int64_t(b) * G + threadIdx.x) * 2], sq = sums[(int64_t(b) * G + threadIdx.x) * 2 + 1];
    const double mean_d = su / cnt;
    double var_d = sq / cnt - mean_d * mean_d;   // biased variance, as torch.nn.GroupNorm
    var_d = var_d < 0.0 ? 0.0 : var_d;
    s_mean[threadIdx.x] = float(mean_d);
    s_rstd[threadIdx.x] = float(1.0 / sqrt(var_d + double(eps)));
  }
  __syncthreads();
  // one CTA per (padded output row, batch image): 32-bit index math, fully coalesced float4 traffic
  const int yo = blockIdx.x;
  int ys = yo - pad;                             // reflect (no edge repeat), pad <= 1
  ys = ys < 0 ? -ys : (ys >= H ? 2 * H - 2 - ys : ys);
  const float* yrow = y + (int64_t(b) * H + ys) * int64_t(W) * C;
  float* orow = out + (int64_t(b) * Hp + yo) * int64_t(Wp) * C;
  const int n = Wp * quads;
  const int cpg = C / G;
#pragma unroll 2
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int xo = i / quads, q = i - xo * quads;
    int xs = xo - pad;
    xs = xs < 0 ? -xs : (xs >= W ? 2 * W - 2 - xs : xs);
    const int g = (4 * q) / cpg;
    const float mean = s_mean[g], rstd = s_rstd[g];
    const float4 v = ldg_stream(yrow + int64_t(xs) * C + 4 * q);
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) bv = *reinterpret_cast<const float4*>(bias + 4 * q);
    const float4 ga = *reinterpret_cast<const float4*>(gamma + 4 * q);
    const float4 be = *reinterpret_cast<const float4*>(beta + 4 * q);
    float r[4] = {v.x + bv.x, v.y + bv.y, v.z + bv.z, v.w + bv.w};
    const float gg[4] = {ga.x, ga.y, ga.z, ga.w}, bb[4] = {be.x, be.y, be.z, be.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float t = (r[j] - mean) * rstd * gg[j] + bb[j];
      r[j] = t / (1.f + expf(-t));
    }
    *reinterpret_cast<float4*>(orow + int64_t(xo) * C + 4 * q) = make_float4(r[0], r[1], r[2], r[3]);
  }
}

int launch_gn_stats(const float* y, const float* bias, double* sums, int B, int64_t HW, int C, int G,
                    cudaStream_t st) {
  NAF_REQUIRE(C % 4 == 0 && C <= 1024 && (C / 4) % G == 0 && 256 % (C / 4) == 0, NAF_ERR_UNSUPPORTED,
              "gn_stats: C=%d G=%d not supported (need C/4 | 256 and G | C/4)", C, G);
  NAF_REQUIRE(aligned16(y) && (!bias || aligned16(bias)), NAF_ERR_ALIGNMENT, "gn_stats: 16-byte alignment");
  cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * 2 * size_t(B) * G, st);
  if (e != cudaSuccess) return fail(NAF_ERR_CUDA, "gn_stats: memset failed: %s", cudaGetErrorString(e));
  int chunks = int((148 * 8 + B - 1) / B);
  const int64_t max_chunks = (HW + 63) / 64;
  if (chunks > max_chunks) chunks = int(max_chunks);
  if (chunks < 1) chunks = 1;
  const dim3 grid = dim3(unsigned(chunks), unsigned(B), 1u);
  gn_stats_kernel<<<grid, 256, 256 * 2 * sizeof(float), st>>>(y, bias, sums, HW, C, G, chunks);
  return check_launch("gn_stats");
}

int launch_gn_silu_apply(const float* y, const float* bias, const float* gamma, const float* beta,
                         const double* sums, float* out, int B, int H, int W, int C, int G, float eps,
                         int pad, cudaStream_t st) {
  NAF_REQUIRE(C % 4 == 0 && (C / G) % 4 == 0 && G <= 64, NAF_ERR_UNSUPPORTED, "gn_silu_apply: C=%d G=%d not supported", C, G);
  NAF_REQUIRE(pad == 0 || (pad == 1 && H >= 2 && W >= 2), NAF_ERR_UNSUPPORTED, "gn_silu_apply: pad must be 0 or 1");
  NAF_REQUIRE(aligned16(y) && aligned16(out) && aligned16(gamma) && aligned16(beta) && (!bias || aligned16(bias)),
              NAF_ERR_ALIGNMENT, "gn_silu_apply: 16-byte alignment");
  NAF_REQUIRE(B <= 65535, NAF_ERR_UNSUPPORTED, "gn_silu_apply: batch too large");
  const dim3 grid = dim3(unsigned(H + 2 * pad), unsigned(B), 1u);
  gn_silu_apply_kernel<<<grid, 256, 0, st>>>(y, bias, gamma, beta, sums, out, H, W, C, G, eps, pad);
  return check_launch("gn_silu_apply");
}

}  // namespace naf
