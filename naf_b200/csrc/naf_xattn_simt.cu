// Cell kernel, fp32 SIMT: cross-scale neighbourhood attention for integer ratios.
//
// For Ho = rh*h, Wo = rw*w every target pixel of one low-res cell (rh x rw pixels) attends over
// the SAME clamped K x K low-res window.  One CTA owns one (batch, cell,
// head): it stages that window of K (K*K x dq) and V (K*K x dv) in shared memory once, then its
// warps stream the cell's pixels in chunks of 32:
//   phase A (lane <-> pixel)   : q row in registers (RoPE applied on the fly), K*K dot products
//                                against broadcast K rows, softmax statistics per thread,
//                                un-normalised probabilities parked in a per-warp smem tile;
//   phase B (lane <-> channels): P (8 pixels x taps, broadcast) times V (taps x 64-channel slabs),
//                                fp32x2 FMAs, 256 B coalesced streaming stores of the output.
// HBM traffic is the algorithmic minimum: q read once, out written once, windows hit L2.
// Bound: fp32 FMA pipe (2*K^2*(dq+dv) flop per pixel-head), not HBM -- the tensor-core cell
// kernel (naf_xattn_tc.cu) lifts that.
//
// Reference semantics: src/layers/attentions.py:16-29,53-75; RoPE src/layers/rope.py:137-153.
#include "naf_common.cuh"

namespace naf {

constexpr int kCellMaxWarps = 8;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// DQ: query/key head dim.  NJ: number of 64-channel slabs of the value head dim (dv <= 64*NJ).
template <int DQ, int NJ>
__global__ void __launch_bounds__(kCellMaxWarps * 32, 2)
xattn_cell_simt_kernel(naf_xattn_params p, int rh, int rw, int dv) {
  extern __shared__ __align__(16) float smem[];
  const int K = p.K, K2 = K * K;
  const int nwarps = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int VSTR = NJ * 64;                 // V row stride (zero padded past dv)
  float* Kwin = smem;                           // [K2][DQ]
  float* Vwin = Kwin + K2 * DQ;                 // [K2][VSTR]
  float* Swarp = Vwin + K2 * VSTR + warp * (K2 * 32 + 32);  // [K2][32] + inv_l[32]
  float* inv_l = Swarp + K2 * 32;

  // ---- which (batch, cell, head)
  int bid = blockIdx.x;
  const int head = bid % p.heads;
  bid /= p.heads;
  const int cj = bid % p.w;
  bid /= p.w;
  const int ci = bid % p.h;
  const int b = bid / p.h;
  const int wy0 = window_origin(ci, p.h, K);
  const int wx0 = window_origin(cj, p.w, K);

  // ---- stage the K and V windows (cp.async, 16 B granules)
  {
    constexpr int KQ4 = DQ / 4;
    for (int i = threadIdx.x; i < K2 * KQ4; i += blockDim.x) {
      const int tap = i / KQ4, d4 = i - tap * KQ4;
      const int t = tap / K, u = tap - t * K;
      const float* src = p.k + (int64_t(b * p.h + wy0 + t) * p.w + wx0 + u) * p.D + head * DQ + d4 * 4;
      cp_async16(Kwin + tap * DQ + d4 * 4, src);
    }
    const int V4 = dv / 4;  // dv % 4 == 0 guaranteed by the launcher
    constexpr int VS4 = VSTR / 4;
    for (int i = threadIdx.x; i < K2 * VS4; i += blockDim.x) {
      const int tap = i / VS4, c4 = i - tap * VS4;
      float* dst = Vwin + tap * VSTR + c4 * 4;
      if (c4 < V4) {
        const int t = tap / K, u = tap - t * K;
        const float* src = p.v + (int64_t(b * p.h + wy0 + t) * p.w + wx0 + u) * p.C + head * dv + c4 * 4;
        cp_async16(dst, src);
      } else {
        *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    cp_async_wait_all();
    __syncthreads();
  }

  const bool rope = p.cos_y != nullptr;
  constexpr int HALF = DQ / 2, P = DQ / 4;
  const int npix = rh * rw;
  const int nchunks = (npix + 31) >> 5;
  const int y0 = ci * rh, x0 = cj * rw;
  const float qscale = p.scale * 1.4426950408889634f;  // fold log2(e): softmax via exp2
  const float* qbase = p.q + int64_t(b) * p.q_stride_b + head * DQ;
  float* obase = static_cast<float*>(p.out) + int64_t(b) * p.Ho * p.Wo * p.C + head * dv;

  for (int ch = warp; ch < nchunks; ch += nwarps) {
    // ================= phase A: lane <-> pixel =================
    {
      int pi = ch * 32 + lane;
      if (pi >= npix) pi = npix - 1;  // duplicate the last pixel; its stores are masked later
      const int py = pi / rw;
      const int y = y0 + py, x = x0 + (pi - py * rw);
      const float* qp = qbase + int64_t(y / p.rep_y) * p.q_stride_y + int64_t(x / p.rep_x) * p.q_stride_x;
      float2 q[DQ / 2];
#pragma unroll
      for (int i = 0; i < DQ / 4; ++i) {
        const float4 t = ldg_stream(qp + 4 * i);
        q[2 * i] = make_float2(t.x, t.y);
        q[2 * i + 1] = make_float2(t.z, t.w);
      }
      if (rope) {
        // pair (i, i+HALF); angle index i: rows for i < P, columns otherwise
#pragma unroll
        for (int i4 = 0; i4 < HALF / 4; ++i4) {
          const bool on_y = (i4 * 4) < P;
          const float* ct = on_y ? p.cos_y + int64_t(y) * P + i4 * 4 : p.cos_x + int64_t(x) * P + (i4 * 4 - P);
          const float* st = on_y ? p.sin_y + int64_t(y) * P + i4 * 4 : p.sin_x + int64_t(x) * P + (i4 * 4 - P);
          const float4 c = *reinterpret_cast<const float4*>(ct);
          const float4 s = *reinterpret_cast<const float4*>(st);
          float2& a0 = q[2 * i4];
          float2& a1 = q[2 * i4 + 1];
          float2& b0 = q[HALF / 2 + 2 * i4];
          float2& b1 = q[HALF / 2 + 2 * i4 + 1];
          float ra, rb;
          ra = a0.x * c.x - b0.x * s.x; rb = b0.x * c.x + a0.x * s.x; a0.x = ra; b0.x = rb;
          ra = a0.y * c.y - b0.y * s.y; rb = b0.y * c.y + a0.y * s.y; a0.y = ra; b0.y = rb;
          ra = a1.x * c.z - b1.x * s.z; rb = b1.x * c.z + a1.x * s.z; a1.x = ra; b1.x = rb;
          ra = a1.y * c.w - b1.y * s.w; rb = b1.y * c.w + a1.y * s.w; a1.y = ra; b1.y = rb;
        }
      }
#pragma unroll
      for (int i = 0; i < DQ / 2; ++i) {
        q[i].x *= qscale;
        q[i].y *= qscale;
      }
      float m = -INFINITY;
      int tap = 0;
      for (; tap + 1 < K2; tap += 2) {
        const float4* k0 = reinterpret_cast<const float4*>(Kwin + tap * DQ);
        const float4* k1 = reinterpret_cast<const float4*>(Kwin + (tap + 1) * DQ);
        float2 a0 = make_float2(0.f, 0.f), a1 = a0, c0 = a0, c1 = a0;
#pragma unroll
        for (int i = 0; i < DQ / 4; ++i) {
          const float4 ka = k0[i];
          const float4 kb = k1[i];
          ffma2(a0, q[2 * i], make_float2(ka.x, ka.y));
          ffma2(c0, q[2 * i], make_float2(kb.x, kb.y));
          ffma2(a1, q[2 * i + 1], make_float2(ka.z, ka.w));
          ffma2(c1, q[2 * i + 1], make_float2(kb.z, kb.w));
        }
        const float s0 = (a0.x + a0.y) + (a1.x + a1.y);
        const float s1 = (c0.x + c0.y) + (c1.x + c1.y);
        Swarp[tap * 32 + lane] = s0;
        Swarp[(tap + 1) * 32 + lane] = s1;
        m = fmaxf(m, fmaxf(s0, s1));
      }
      if (tap < K2) {
        const float4* k0 = reinterpret_cast<const float4*>(Kwin + tap * DQ);
        float2 a0 = make_float2(0.f, 0.f), a1 = a0;
#pragma unroll
        for (int i = 0; i < DQ / 4; ++i) {
          const float4 ka = k0[i];
          ffma2(a0, q[2 * i], make_float2(ka.x, ka.y));
          ffma2(a1, q[2 * i + 1], make_float2(ka.z, ka.w));
        }
        const float s0 = (a0.x + a0.y) + (a1.x + a1.y);
        Swarp[tap * 32 + lane] = s0;
        m = fmaxf(m, s0);
      }
      float l = 0.f;
#pragma unroll 7
      for (int t2 = 0; t2 < K2; ++t2) {
        const float e = fast_exp2(Swarp[t2 * 32 + lane] - m);
        Swarp[t2 * 32 + lane] = e;
        l += e;
      }
      inv_l[lane] = 1.f / l;
    }
    __syncwarp();

    // ================= phase B: lane <-> channel pairs =================
#pragma unroll 1
    for (int sub = 0; sub < 4; ++sub) {
      const int pbase = ch * 32 + sub * 8;
      if (pbase >= npix) break;
      float2 acc[8][NJ];
#pragma unroll
      for (int px = 0; px < 8; ++px)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[px][j] = make_float2(0.f, 0.f);
      const float* sp = Swarp + sub * 8;
      const float* vp = Vwin + 2 * lane;
#pragma unroll 2
      for (int tap = 0; tap < K2; ++tap) {
        const float4 pa = *reinterpret_cast<const float4*>(sp + tap * 32);
        const float4 pb = *reinterpret_cast<const float4*>(sp + tap * 32 + 4);
        float2 v[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) v[j] = *reinterpret_cast<const float2*>(vp + tap * VSTR + 64 * j);
        const float pr[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
#pragma unroll
        for (int px = 0; px < 8; ++px) {
          const float2 pp = make_float2(pr[px], pr[px]);
#pragma unroll
          for (int j = 0; j < NJ; ++j) ffma2(acc[px][j], pp, v[j]);
        }
      }
#pragma unroll
      for (int px = 0; px < 8; ++px) {
        const int pi = pbase + px;
        if (pi < npix) {
          const float il = inv_l[sub * 8 + px];
          const int py = pi / rw;
          const int y = y0 + py, x = x0 + (pi - py * rw);
          float* op = obase + (int64_t(y) * p.Wo + x) * p.C + 2 * lane;
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            if (2 * lane + 64 * j < dv)
              stg_stream2(op + 64 * j, make_float2(acc[px][j].x * il, acc[px][j].y * il));
          }
        }
      }
    }
    __syncwarp();
  }
}

static size_t cell_simt_smem(int K, int dq, int nj, int warps) {
  const size_t K2 = size_t(K) * K;
  return (K2 * dq + K2 * nj * 64 + size_t(warps) * (K2 * 32 + 32)) * sizeof(float);
}

static int cell_simt_warps(int K, int dq, int nj) {
  for (int wps = kCellMaxWarps; wps >= 2; wps >>= 1)
    if (cell_simt_smem(K, dq, nj, wps) <= 227 * 1024) return wps;
  return 0;
}

bool xattn_cell_simt_supported(const naf_xattn_params& p, const char** why) {
  if (p.Kw != 0 && p.Kw != p.K) { *why = "rectangular window"; return false; }
  if (p.out_dtype != NAF_DTYPE_F32) { *why = "fp32 output only"; return false; }
  const int dq = p.D / p.heads, dv = p.C / p.heads;
  if (p.row_tap || p.col_tap) { *why = "tap tables given (non-integer ratio path)"; return false; }
  if (p.Ho % p.h || p.Wo % p.w) { *why = "target size is not a multiple of the feature size"; return false; }
  if (p.scores) { *why = "score output requested"; return false; }
  if (dq != 64 && dq != 32 && dq != 16) { *why = "head dim must be 16, 32 or 64"; return false; }
  if (dv % 4 != 0 || dv > 256) { *why = "value head dim must be a multiple of 4 and <= 256"; return false; }
  if ((p.Ho / p.h) * (p.Wo / p.w) < 32) { *why = "fewer than 32 pixels per cell"; return false; }
  if (!aligned16(p.q) || !aligned16(p.k) || !aligned16(p.v) || !aligned16(p.out) ||
      (p.q_stride_b % 4) || (p.q_stride_y % 4) || (p.q_stride_x % 4)) {
    *why = "pointers/strides not 16-byte aligned";
    return false;
  }
  if (p.cos_y && !(aligned16(p.cos_y) && aligned16(p.sin_y) && aligned16(p.cos_x) && aligned16(p.sin_x) && (dq % 16 == 0))) {
    *why = "rope tables not 16-byte aligned or head dim % 16 != 0";
    return false;
  }
  const int nj = (dv + 63) / 64;
  if (cell_simt_warps(p.K, dq, nj) == 0) { *why = "window does not fit in shared memory"; return false; }
  if (int64_t(p.B) * p.h * p.w * p.heads >= (int64_t(1) << 31)) { *why = "grid too large"; return false; }
  return true;
}

template <int DQ, int NJ>
static int launch_cell(const naf_xattn_params& p, cudaStream_t st) {
  const int dv = p.C / p.heads;
  const int warps = cell_simt_warps(p.K, DQ, NJ);
  const size_t smem = cell_simt_smem(p.K, DQ, NJ, warps);
  cudaError_t e = cudaFuncSetAttribute(xattn_cell_simt_kernel<DQ, NJ>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (e != cudaSuccess)
    return fail(NAF_ERR_CUDA, "xattn(cell-simt): smem opt-in failed: %s", cudaGetErrorString(e));
  const unsigned grid = unsigned(p.B) * p.h * p.w * p.heads;
  xattn_cell_simt_kernel<DQ, NJ><<<grid, warps * 32, smem, st>>>(p, p.Ho / p.h, p.Wo / p.w, dv);
  return check_launch("xattn_cell_simt");
}

template <int DQ>
static int launch_cell_dq(const naf_xattn_params& p, cudaStream_t st) {
  const int nj = (p.C / p.heads + 63) / 64;
  switch (nj) {
    case 1: return launch_cell<DQ, 1>(p, st);
    case 2: return launch_cell<DQ, 2>(p, st);
    case 3: return launch_cell<DQ, 3>(p, st);
    default: return launch_cell<DQ, 4>(p, st);
  }
}

int launch_xattn_cell_simt(const naf_xattn_params& p, cudaStream_t st) {
  switch (p.D / p.heads) {
    case 64: return launch_cell_dq<64>(p, st);
    case 32: return launch_cell_dq<32>(p, st);
    default: return launch_cell_dq<16>(p, st);
  }
}

}  // namespace naf
