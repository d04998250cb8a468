// RoPE + key pooling pre-pass:  q = RoPE(x) ; k = adaptive_avg_pool2d(q, (h, w)).
// Reference: src/layers/rope.py:137-174 (rotation), src/model/naf.py:63-69 (key pooling).
//
// One CTA per (batch, pooling bin).  A thread owns VEC consecutive rotation pairs (a_i, b_i) =
// (x[c], x[c + D_head/2]) of one rope head, so the rotation needs no data exchange; the G pixel
// groups of a CTA stride over the pixels of the bin and are reduced through shared memory.
// HBM-bound: reads x once (1 KB per pixel at D=256), writes k (tiny) and optionally q.
#include "naf_common.cuh"

namespace naf {

template <int VEC>
struct Vec;
template <>
struct Vec<4> { using T = float4; };
template <>
struct Vec<1> { using T = float; };

template <int VEC>
__device__ __forceinline__ void load_vec(float (&d)[VEC], const float* p) {
  if constexpr (VEC == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    d[0] = t.x; d[1] = t.y; d[2] = t.z; d[3] = t.w;
  } else {
    d[0] = *p;
  }
}
template <int VEC>
__device__ __forceinline__ void load_vec_stream(float (&d)[VEC], const float* p) {
  if constexpr (VEC == 4) {
    const float4 t = ldg_stream(p);
    d[0] = t.x; d[1] = t.y; d[2] = t.z; d[3] = t.w;
  } else {
    d[0] = __ldg(p);
  }
}
template <int VEC>
__device__ __forceinline__ void store_vec(float* p, const float (&d)[VEC]) {
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(d[0], d[1], d[2], d[3]);
  } else {
    *p = d[0];
  }
}

template <int VEC>
__global__ void __launch_bounds__(256)
rope_kpool_kernel(naf_kpool_params p, int lanes, int groups, int bins_h, int bins_w) {
  extern __shared__ float red[];  // [groups][lanes][2*VEC]
  const int lane = threadIdx.x % lanes;
  const int group = threadIdx.x / lanes;
  const bool active = group < groups;

  const int b = blockIdx.z;
  const int bi = blockIdx.y, bj = blockIdx.x;
  // ATen adaptive pooling bin: [floor(i*In/Out), ceil((i+1)*In/Out))
  const int ys = int((int64_t(bi) * p.Ho) / bins_h);
  const int ye = int((int64_t(bi + 1) * p.Ho + bins_h - 1) / bins_h);
  const int xs = int((int64_t(bj) * p.Wo) / bins_w);
  const int xe = int((int64_t(bj + 1) * p.Wo + bins_w - 1) / bins_w);
  const int bw = xe - xs;
  const int npix = (ye - ys) * bw;

  const bool rope = p.cos_y != nullptr;
  const int d_head = rope ? p.D / p.rope_heads : p.D;
  const int half = d_head / 2;
  const int P = d_head / 4;
  // pair index of this thread inside the whole D vector
  const int pair0 = lane * VEC;            // in [0, D/2)
  const int head = pair0 / half;
  const int i0 = pair0 - head * half;      // pair index inside the head, [0, half)
  const int ca = head * d_head + i0;       // channel of a_i ; b_i is ca + half
  const bool on_y = i0 < P;
  const int ti = on_y ? i0 : i0 - P;

  float sa[VEC], sb[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) sa[v] = sb[v] = 0.f;

  if (active) {
    const float* xb = p.x + int64_t(b) * p.x_stride_b;
    float* qb = p.q_out ? p.q_out + int64_t(b) * p.Ho * p.Wo * p.D : nullptr;
    // pixel (yy, xx) of this group walks the bin row-major with stride `groups`: advance the
    // coordinates incrementally instead of dividing per pixel
    int yy = ys + group / bw, xx = xs + group % bw;
    const int step_y = groups / bw, step_x = groups % bw;
    const bool no_rep = p.rep_y == 1 && p.rep_x == 1;
#pragma unroll 4
    for (int pi = group; pi < npix; pi += groups) {
      const float* px = no_rep ? xb + int64_t(yy) * p.x_stride_y + int64_t(xx) * p.x_stride_x
                               : xb + int64_t(yy / p.rep_y) * p.x_stride_y + int64_t(xx / p.rep_x) * p.x_stride_x;
      float a[VEC], bb[VEC];
      load_vec_stream<VEC>(a, px + ca);
      load_vec_stream<VEC>(bb, px + ca + half);
      if (rope) {
        float c[VEC], s[VEC];
        const float* ct = on_y ? p.cos_y + int64_t(yy) * P + ti : p.cos_x + int64_t(xx) * P + ti;
        const float* st = on_y ? p.sin_y + int64_t(yy) * P + ti : p.sin_x + int64_t(xx) * P + ti;
        load_vec<VEC>(c, ct);
        load_vec<VEC>(s, st);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          // rope_apply: x*cos + rotate_half(x)*sin  (src/layers/rope.py:22-34)
          const float ra = a[v] * c[v] - bb[v] * s[v];
          const float rb = bb[v] * c[v] + a[v] * s[v];
          a[v] = ra;
          bb[v] = rb;
        }
      }
      if (qb) {
        float* pq = qb + (int64_t(yy) * p.Wo + xx) * p.D;
        store_vec<VEC>(pq + ca, a);
        store_vec<VEC>(pq + ca + half, bb);
      }
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        sa[v] += a[v];
        sb[v] += bb[v];
      }
      yy += step_y;
      xx += step_x;
      if (xx >= xe) {
        xx -= bw;
        ++yy;
      }
    }
  }
  if (!p.k_out) return;

  if (active) {
    float* r = red + (size_t(group) * lanes + lane) * 2 * VEC;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      r[v] = sa[v];
      r[VEC + v] = sb[v];
    }
  }
  __syncthreads();
  if (active && group == 0) {
    for (int g = 1; g < groups; ++g) {
      const float* r = red + (size_t(g) * lanes + lane) * 2 * VEC;
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        sa[v] += r[v];
        sb[v] += r[VEC + v];
      }
    }
    const float inv = 1.f / float(npix);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      sa[v] *= inv;
      sb[v] *= inv;
    }
    float* ko = p.k_out + ((int64_t(b) * p.h + bi) * p.w + bj) * p.D;
    store_vec<VEC>(ko + ca, sa);
    store_vec<VEC>(ko + ca + half, sb);
  }
}

int launch_rope_kpool(const naf_kpool_params& p, cudaStream_t st) {
  const bool rope = p.cos_y != nullptr;
  const int d_head = rope ? p.D / p.rope_heads : p.D;
  NAF_REQUIRE(d_head % 2 == 0, NAF_ERR_BAD_SHAPE, "rope_kpool: head dim must be even");
  const int half = d_head / 2, P = d_head / 4;
  // vector path needs 16 B alignment of everything a float4 touches
  bool vec4 = (half % 4 == 0) && (!rope || P % 4 == 0) && aligned16(p.x) &&
              (p.x_stride_b % 4 == 0) && (p.x_stride_y % 4 == 0) && (p.x_stride_x % 4 == 0) &&
              (p.D % 4 == 0) && (!p.k_out || aligned16(p.k_out)) &&
              (!p.q_out || aligned16(p.q_out));
  if (rope)
    vec4 = vec4 && aligned16(p.cos_y) && aligned16(p.sin_y) && aligned16(p.cos_x) &&
           aligned16(p.sin_x);
  const int VEC = vec4 ? 4 : 1;
  const int lanes = p.D / (2 * VEC);
  NAF_REQUIRE(lanes >= 1 && lanes <= 256, NAF_ERR_UNSUPPORTED,
              "rope_kpool: embed_dim %d not supported", p.D);
  int groups = 256 / lanes;
  // pooling grid: the key map, or a pseudo grid when only q_out is wanted
  const int bins_h = p.k_out ? p.h : (p.Ho + 7) / 8;
  const int bins_w = p.k_out ? p.w : (p.Wo + 7) / 8;
  const int64_t max_bin = (int64_t(p.Ho + bins_h - 1) / bins_h + 1) * (int64_t(p.Wo + bins_w - 1) / bins_w + 1);
  while (groups > 1 && groups > max_bin) groups >>= 1;
  NAF_REQUIRE(bins_h <= 65535 && p.B <= 65535, NAF_ERR_UNSUPPORTED, "rope_kpool: grid too large");
  dim3 grid(unsigned(bins_w), unsigned(bins_h), unsigned(p.B));
  const int threads = ((lanes * groups + 31) / 32) * 32;
  const size_t smem = size_t(groups) * lanes * 2 * VEC * sizeof(float);
  prefer_max_shared(rope_kpool_kernel<4>);
  prefer_max_shared(rope_kpool_kernel<1>);
  if (VEC == 4)
    rope_kpool_kernel<4><<<grid, threads, smem, st>>>(p, lanes, groups, bins_h, bins_w);
  else
    rope_kpool_kernel<1><<<grid, threads, smem, st>>>(p, lanes, groups, bins_h, bins_w);
  return check_launch("rope_kpool");
}

}  // namespace naf
