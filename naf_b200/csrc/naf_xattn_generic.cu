// Generic cross-scale neighbourhood attention: any head dims, any odd K, tap tables (non-integer
// ratios, duplicated taps) or the integer-ratio rule, optional pre-softmax score output.
// Correctness-first path: one warp per (pixel, head); K and V are read through L1/L2 (the
// low-res maps are a few MB and stay cache resident).  The cell kernels are the fast paths.
//
// Reference semantics: src/layers/attentions.py:16-29 (QK -> *scale -> softmax -> AV),
// :53-75 (dilation, layouts); tap order t_h*Kw + t_w (NATTEN); rectangular windows kernel_size=(Kh, Kw).
#include <cuda_bf16.h>

#include "naf_common.cuh"

namespace naf {

constexpr int kGenericWarps = 4;

__global__ void __launch_bounds__(kGenericWarps * 32)
xattn_generic_kernel(naf_xattn_params p, int rh, int rw, int64_t total_items) {
  extern __shared__ float smem_f[];
  const int Kh = p.K, Kw = p.Kw ? p.Kw : p.K;   // window height / width (rectangular: NATTEN kernel_size=(Kh, Kw))
  const int K2 = Kh * Kw;
  const int dq = p.D / p.heads, dv = p.C / p.heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_warp = 2 * K2 + dq;
  float* sc = smem_f + size_t(warp) * per_warp;          // scores / probabilities [K2]
  int* idx = reinterpret_cast<int*>(sc + K2);             // low-res pixel index    [K2]
  float* sq = sc + 2 * K2;                                // (rotated) query        [dq]

  const bool rope = p.cos_y != nullptr;
  const int half = dq / 2, P = dq / 4;

  for (int64_t item = int64_t(blockIdx.x) * kGenericWarps + warp; item < total_items;
       item += int64_t(gridDim.x) * kGenericWarps) {
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
    for (int tap = lane; tap < K2; tap += 32) {
      const int t = tap / Kw, u = tap - t * Kw;
      const int r = tap_index(p.row_tap, y, t, Kh, rh, p.h);
      const int c = tap_index(p.col_tap, x, u, Kw, rw, p.w);
      idx[tap] = (b * p.h + r) * p.w + c;
    }
    __syncwarp();

    // ---- scores: lane <-> tap
    float m = -INFINITY;
    for (int tap = lane; tap < K2; tap += 32) {
      const float* kp = p.k + int64_t(idx[tap]) * p.D + head * dq;
      float acc = 0.f;
      for (int d = 0; d < dq; ++d) acc = fmaf(sq[d], __ldg(kp + d), acc);
      acc *= p.scale;
      sc[tap] = acc;
      m = fmaxf(m, acc);
    }
    if (p.scores) {
      float* so = p.scores + ((((int64_t(b) * p.heads + head) * p.Ho + y) * p.Wo + x)) * K2;
      for (int tap = lane; tap < K2; tap += 32) so[tap] = sc[tap];
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
    __syncwarp();

    // ---- aggregation: lane <-> value channel
    const int64_t ooff = ((int64_t(b) * p.Ho + y) * p.Wo + x) * p.C + head * dv;
    for (int c0 = 0; c0 < dv; c0 += 32) {
      const int c = c0 + lane;
      float acc = 0.f;
      if (c < dv) {
        for (int tap = 0; tap < K2; ++tap)
          acc = fmaf(sc[tap], __ldg(p.v + int64_t(idx[tap]) * p.C + head * dv + c), acc);
        if (p.out_dtype == NAF_DTYPE_BF16) static_cast<__nv_bfloat16*>(p.out)[ooff + c] = __float2bfloat16_rn(acc * inv);
        else static_cast<float*>(p.out)[ooff + c] = acc * inv;
      }
    }
    __syncwarp();
  }
}

int launch_xattn_generic(const naf_xattn_params& p, cudaStream_t st) {
  const int K2 = p.K * (p.Kw ? p.Kw : p.K);
  const int dq = p.D / p.heads;
  const size_t smem = size_t(kGenericWarps) * (2 * K2 + dq) * sizeof(float);
  NAF_REQUIRE(smem <= 200 * 1024, NAF_ERR_UNSUPPORTED,
              "xattn(generic): kernel_size %d / head dim %d need %zu B shared memory", p.K, dq,
              smem);
  NAF_REQUIRE(int64_t(p.B) * p.h * p.w < (int64_t(1) << 31), NAF_ERR_UNSUPPORTED,
              "xattn(generic): feature map too large");
  if (smem > 48 * 1024) {   // per (kernel, device) opt-in, cached in ensure_dyn_smem
    cudaError_t e = ensure_dyn_smem(xattn_generic_kernel, int(smem));
    if (e != cudaSuccess) return fail(NAF_ERR_CUDA, "xattn(generic): smem opt-in failed: %s", cudaGetErrorString(e));
  }
  const int64_t items = int64_t(p.B) * p.Ho * p.Wo * p.heads;
  int64_t blocks = (items + kGenericWarps - 1) / kGenericWarps;
  const int64_t cap = int64_t(148) * 16 * 8;
  if (blocks > cap) blocks = cap;
  xattn_generic_kernel<<<unsigned(blocks), kGenericWarps * 32, smem, st>>>(
      p, p.Ho / p.h, p.Wo / p.w, items);
  return check_launch("xattn_generic");
}

}  // namespace naf
