// Layout packing (B,C,H,W; any strides) -> pixel-major (B,H,W,C), and the tap-index dump hook.
#include "naf_common.cuh"

namespace naf {

// 32 channels x 32 pixels per tile; reads run along pixels (the contiguous axis of NCHW), writes
// run along channels (the contiguous axis of the packed layout).  Both sides are 128 B coalesced.
__global__ void __launch_bounds__(256)
pack_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int H, int W,
                 int64_t sb, int64_t sc, int64_t sh, int64_t sw, int dstC, int dst_off) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32;
  const int64_t p0 = int64_t(blockIdx.x) * 32;
  const int64_t HW = int64_t(H) * W;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int64_t p = p0 + tx;
  int y = 0, x = 0;
  if (p < HW) {
    y = int(p / W);
    x = int(p - int64_t(y) * W);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i;
    float val = 0.f;
    if (c < C && p < HW) val = src[b * sb + c * sc + y * sh + x * sw];
    tile[ty + 8 * i][tx] = val;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t pp = p0 + ty + 8 * i;
    const int c = c0 + tx;
    if (c < C && pp < HW) dst[(int64_t(b) * HW + pp) * dstC + dst_off + c] = tile[tx][ty + 8 * i];
  }
}

int launch_pack_nhwc(const float* src, float* dst, int B, int C, int H, int W, int64_t sb,
                     int64_t sc, int64_t sh, int64_t sw, int dstC, int dst_off, cudaStream_t st) {
  const int64_t HW = int64_t(H) * W;
  const int64_t ptiles = (HW + 31) / 32;
  NAF_REQUIRE(ptiles < (int64_t(1) << 31) && B <= 65535 && (C + 31) / 32 <= 65535,
              NAF_ERR_UNSUPPORTED, "pack_nhwc: tensor too large for one launch");
  dim3 grid(unsigned(ptiles), unsigned((C + 31) / 32), unsigned(B));
  prefer_max_shared(pack_nhwc_kernel);
  pack_nhwc_kernel<<<grid, 256, 0, st>>>(src, dst, C, H, W, sb, sc, sh, sw, dstC, dst_off);
  return check_launch("pack_nhwc");
}

// out[p][0:Ca] = a[p][:] + bias_a ; out[p][Ca:Ca+Cb] = b[p][:] + bias_b   (pixel-major, float4)
__global__ void __launch_bounds__(256)
concat_bias_kernel(const float* __restrict__ a, const float* __restrict__ bias_a, int Ca,
                   const float* __restrict__ b, const float* __restrict__ bias_b, int Cb,
                   float* __restrict__ out, int64_t npix) {
  const int qa = Ca >> 2, qb = Cb >> 2, qt = qa + qb;
  const int64_t total = npix * qt;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int q = int(i % qt);
    const int64_t p = i / qt;
    float4 v, bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < qa) {
      v = ldg_stream(a + p * Ca + 4 * q);
      if (bias_a) bv = *reinterpret_cast<const float4*>(bias_a + 4 * q);
    } else {
      v = ldg_stream(b + p * Cb + 4 * (q - qa));
      if (bias_b) bv = *reinterpret_cast<const float4*>(bias_b + 4 * (q - qa));
    }
    *reinterpret_cast<float4*>(out + p * (Ca + Cb) + 4 * q) =
        make_float4(v.x + bv.x, v.y + bv.y, v.z + bv.z, v.w + bv.w);
  }
}

int launch_concat_bias(const float* a, const float* bias_a, int Ca, const float* b, const float* bias_b,
                       int Cb, float* out, int64_t npix, cudaStream_t st) {
  NAF_REQUIRE(Ca % 4 == 0 && Cb % 4 == 0, NAF_ERR_UNSUPPORTED, "concat_bias: channel counts must be multiples of 4");
  NAF_REQUIRE(aligned16(a) && aligned16(b) && aligned16(out) && (!bias_a || aligned16(bias_a)) &&
                  (!bias_b || aligned16(bias_b)),
              NAF_ERR_ALIGNMENT, "concat_bias: 16-byte alignment");
  const int64_t total = npix * ((Ca + Cb) / 4);
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  prefer_max_shared(concat_bias_kernel);
  concat_bias_kernel<<<unsigned(blocks), 256, 0, st>>>(a, bias_a, Ca, b, bias_b, Cb, out, npix);
  return check_launch("concat_bias");
}

// One thread per (pixel, tap): writes the linear low-res cell index the kernels gather from.
__global__ void dump_taps_kernel(int32_t* __restrict__ out, const int32_t* __restrict__ row_tap,
                                 const int32_t* __restrict__ col_tap, int Ho, int Wo, int h, int w,
                                 int K) {
  const int64_t total = int64_t(Ho) * Wo * K * K;
  const int rh = Ho / h, rw = Wo / w;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int tap = int(i % (K * K));
    const int64_t pix = i / (K * K);
    const int x = int(pix % Wo), y = int(pix / Wo);
    const int t = tap / K, u = tap % K;
    const int r = tap_index(row_tap, y, t, K, rh, h);
    const int c = tap_index(col_tap, x, u, K, rw, w);
    out[i] = r * w + c;
  }
}

int launch_dump_taps(int32_t* idx_out, const int32_t* row_tap, const int32_t* col_tap, int Ho,
                     int Wo, int h, int w, int K, cudaStream_t st) {
  const int64_t total = int64_t(Ho) * Wo * K * K;
  const int blocks = int((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
  dump_taps_kernel<<<blocks, 256, 0, st>>>(idx_out, row_tap, col_tap, Ho, Wo, h, w, K);
  return check_launch("dump_taps");
}

}  // namespace naf
