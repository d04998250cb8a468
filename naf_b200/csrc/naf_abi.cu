// extern "C" entry points of libnaf_b200.so: argument validation, algorithm selection, error
// reporting.  See include/naf_b200.h for the contract of every function.
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <utility>

#include "naf_common.cuh"

namespace naf {

char* last_error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(naf_status code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buffer(), 512, fmt, ap);
  va_end(ap);
  return static_cast<int>(code);
}

// kernel launches performed by this library since it was loaded, per kernel family (the `what` of
// check_launch); read through naf_launch_count().  Counting lives here, not in the Python layer.
static std::mutex g_count_mu;
static std::map<std::string, unsigned long long> g_counts;

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return fail(NAF_ERR_CUDA, "%s: CUDA error %d (%s)", what, int(e), cudaGetErrorString(e));
  {
    std::lock_guard<std::mutex> lock(g_count_mu);
    ++g_counts[what];
  }
  return NAF_OK;
}

unsigned long long launch_count(const char* prefix) {
  std::lock_guard<std::mutex> lock(g_count_mu);
  unsigned long long n = 0;
  const size_t len = prefix ? strlen(prefix) : 0;
  for (const auto& kv : g_counts)
    if (len == 0 || kv.first.compare(0, len, prefix) == 0) n += kv.second;
  return n;
}

int device_sm_count() {
  constexpr int kMaxDev = 64;
  static int cache[kMaxDev] = {0};   // benign race: every writer stores the same value
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < kMaxDev && cache[dev] > 0) return cache[dev];
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (dev >= 0 && dev < kMaxDev) cache[dev] = sms;
  return sms;
}

cudaError_t ensure_dyn_smem_impl(const void* func, int bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, int> configured;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  int& have = configured[std::make_pair(func, dev)];
  if (bytes <= have) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) have = bytes;
  return e;
}

void prefer_max_shared_impl(const void* func) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, bool> done;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  bool& d = done[std::make_pair(func, dev)];
  if (d) return;
  cudaFuncSetAttribute(func, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  d = true;
}

static int validate_xattn(const naf_xattn_params& p) {
  NAF_REQUIRE(p.q && p.k && p.v && p.out, NAF_ERR_NULL, "xattn: q, k, v and out must be non-NULL");
  NAF_REQUIRE(p.B > 0 && p.D > 0 && p.C > 0 && p.heads > 0 && p.Ho > 0 && p.Wo > 0 && p.h > 0 &&
                  p.w > 0,
              NAF_ERR_BAD_SHAPE, "xattn: sizes must be positive");
  NAF_REQUIRE(p.D % p.heads == 0, NAF_ERR_BAD_SHAPE, "dim must be divisible by num_heads");
  NAF_REQUIRE(p.C % p.heads == 0, NAF_ERR_BAD_SHAPE,
              "value channels (%d) must be divisible by num_heads (%d)", p.C, p.heads);
  NAF_REQUIRE(p.K >= 1 && (p.K & 1), NAF_ERR_WINDOW, "kernel_size must be odd and >= 1, got %d",
              p.K);
  NAF_REQUIRE(p.Kw == 0 || (p.Kw >= 1 && (p.Kw & 1)), NAF_ERR_WINDOW,
              "kernel_size must be odd and >= 1, got (%d, %d)", p.K, p.Kw);
  const int kw_ = p.Kw ? p.Kw : p.K;
  const bool have_tables = p.row_tap && p.col_tap;
  NAF_REQUIRE(have_tables || (!p.row_tap && !p.col_tap), NAF_ERR_NULL,
              "xattn: row_tap and col_tap must both be given or both be NULL");
  if (!have_tables) {
    NAF_REQUIRE(p.Ho % p.h == 0 && p.Wo % p.w == 0, NAF_ERR_BAD_SHAPE,
                "xattn: tap tables are required when the target size is not a multiple of the "
                "feature size");
  }
  // NATTEN's constraint kernel_size * dilation <= axis length (dilation = Ho // h)
  NAF_REQUIRE(p.Ho / p.h >= 1 && p.Wo / p.w >= 1 && int64_t(p.K) * (p.Ho / p.h) <= p.Ho &&
                  int64_t(kw_) * (p.Wo / p.w) <= p.Wo,
              NAF_ERR_WINDOW,
              "kernel_size*dilation exceeds the target size (K=%dx%d, target %dx%d, features %dx%d)",
              p.K, kw_, p.Ho, p.Wo, p.h, p.w);
  const int nrope = (p.cos_y != nullptr) + (p.sin_y != nullptr) + (p.cos_x != nullptr) +
                    (p.sin_x != nullptr);
  NAF_REQUIRE(nrope == 0 || nrope == 4, NAF_ERR_NULL,
              "xattn: rope tables must be all given or all NULL");
  if (nrope == 4)
    NAF_REQUIRE((p.D / p.heads) % 4 == 0, NAF_ERR_BAD_SHAPE,
                "xattn: on-the-fly RoPE needs head dim %% 4 == 0");
  NAF_REQUIRE(p.q_stride_x >= p.D && p.q_stride_y >= 0 && p.q_stride_b >= 0, NAF_ERR_BAD_SHAPE,
              "xattn: bad q strides");
  NAF_REQUIRE(p.rep_y >= 1 && p.rep_x >= 1 && p.Ho % p.rep_y == 0 && p.Wo % p.rep_x == 0,
              NAF_ERR_BAD_SHAPE, "xattn: replication factors (%d,%d) must divide the target size",
              p.rep_y, p.rep_x);
  NAF_REQUIRE(p.out_dtype == NAF_DTYPE_F32 || p.out_dtype == NAF_DTYPE_BF16, NAF_ERR_UNSUPPORTED,
              "xattn: out_dtype %d (0 = f32, 1 = bf16)", p.out_dtype);
  for (int dt : {p.q_dtype, p.k_dtype, p.v_dtype})
    NAF_REQUIRE(dt == NAF_DTYPE_F32 || dt == NAF_DTYPE_BF16, NAF_ERR_UNSUPPORTED, "xattn: input dtype %d (0 = f32, 1 = bf16)", dt);
  return NAF_OK;
}

static int select_algo(const naf_xattn_params& p, bool explain) {
  const char* why = "";
  if (p.q_dtype != NAF_DTYPE_F32 || p.k_dtype != NAF_DTYPE_F32 || p.v_dtype != NAF_DTYPE_F32) {
    // bf16 inputs: only the TMA kernel reads them natively
    if ((p.algo == NAF_ALGO_AUTO || p.algo == NAF_ALGO_CELL_TMA) && xattn_cell_tma_supported(p, &why)) return NAF_ALGO_CELL_TMA;
    return -fail(NAF_ERR_UNSUPPORTED, "xattn: bf16 inputs are read natively by the TMA cell kernel only (%s); pass fp32 "
                                       "tensors to the other kernels", why);
  }
  if (p.algo == NAF_ALGO_GENERIC) return NAF_ALGO_GENERIC;
  if (p.Kw != 0 && p.Kw != p.K) {
    // rectangular windows (NATTEN kernel_size=(kh, kw)): the generic kernel
    if (p.algo == NAF_ALGO_AUTO) return NAF_ALGO_GENERIC;
    return -fail(NAF_ERR_UNSUPPORTED, "xattn: rectangular windows (%dx%d) run on the generic kernel only", p.K, p.Kw);
  }
  if (p.algo == NAF_ALGO_CELL_TMA) {
    if (xattn_cell_tma_supported(p, &why)) return NAF_ALGO_CELL_TMA;
    return -fail(NAF_ERR_UNSUPPORTED, "xattn: TMA tensor-core cell kernel unsupported: %s", why);
  }
  if (p.algo == NAF_ALGO_CELL_TCWS) {
    if (xattn_cell_tcws_supported(p, &why)) return NAF_ALGO_CELL_TCWS;
    return -fail(NAF_ERR_UNSUPPORTED, "xattn: pipelined tensor-core cell kernel unsupported: %s", why);
  }
  if (p.algo == NAF_ALGO_CELL_TC)
    return -fail(NAF_ERR_UNSUPPORTED, "xattn: the non-pipelined tensor-core kernel (algo 3) was removed in ABI v3; "
                                       "use NAF_ALGO_CELL_TCWS");
  if (p.algo == NAF_ALGO_CELL_SIMT) {
    if (xattn_cell_simt_supported(p, &why)) return NAF_ALGO_CELL_SIMT;
    return -fail(NAF_ERR_UNSUPPORTED, "xattn: SIMT cell kernel unsupported: %s", why);
  }
  if (p.algo == NAF_ALGO_UNION_TC) {
    if (xattn_union_tc_supported(p, &why)) return NAF_ALGO_UNION_TC;
    return -fail(NAF_ERR_UNSUPPORTED, "xattn: union-window tensor-core kernel unsupported: %s", why);
  }
  if (p.algo != NAF_ALGO_AUTO) return -fail(NAF_ERR_UNSUPPORTED, "xattn: unknown algo %d", p.algo);
  (void)explain;
  const bool tma_ok = xattn_cell_tma_supported(p, &why);
  const bool ws_ok = xattn_cell_tcws_supported(p, &why);
  if (tma_ok && ws_ok) {
    // Both pipelined tensor-core kernels take the request: measured choice (profiles/).  The TMA
    // kernel wins where windows turn over quickly and for wide value heads (C2 4.41 vs 4.62 ms, C3 5.8 vs 8.5 ms);
    // the in-kernel-conversion kernel keeps 2-5 % on very large cells, where a window serves >= 16 tiles and its
    // per-thread bulk-copy epilogue streams better (C4 2.11 vs 2.17 ms, C5 10.6 vs 11.1 ms), and on latency-bound
    // launches that cannot amortise the plane pre-pass (C1: 0.067 vs 0.078 ms).
    const int64_t cell_px = int64_t(p.Ho / p.h) * (p.Wo / p.w);
    const int64_t items = int64_t(p.B) * p.h * p.w * p.heads;
    if (cell_px >= 2048 || items < 2048) return NAF_ALGO_CELL_TCWS;
    return NAF_ALGO_CELL_TMA;
  }
  if (tma_ok) return NAF_ALGO_CELL_TMA;
  if (ws_ok) return NAF_ALGO_CELL_TCWS;
  // Everything the pipelined cell kernels refuse: tap tables (non-integer ratios), ratio 1, cells of a few pixels,
  // head dims other than 64, K > 11.  Measured (scripts/check_union.py, profiles/r2_union_tc.txt): the union-window
  // tensor-core kernel is 7-36x faster than the warp-per-pixel kernel and 2.5x faster than the fp32 cell kernel
  // at K = 13; the fp32 cell kernel keeps narrow heads (dq <= 32: 0.81 vs 0.92 ms at r = 14).
  const bool simt_ok = xattn_cell_simt_supported(p, &why);
  const bool union_ok = xattn_union_tc_supported(p, &why);
  if (simt_ok && (!union_ok || p.D / p.heads <= 32)) return NAF_ALGO_CELL_SIMT;
  if (union_ok) return NAF_ALGO_UNION_TC;
  return NAF_ALGO_GENERIC;
}

}  // namespace naf

using namespace naf;

extern "C" {

int naf_abi_version(void) { return NAF_ABI_VERSION; }

const char* naf_last_error(void) { return last_error_buffer(); }

unsigned long long naf_launch_count(const char* prefix) { return launch_count(prefix); }

int naf_has_tensor_path(void) {
#ifdef NAF_WITH_TC
  return 1;
#else
  return 0;
#endif
}

int naf_pack_nhwc_f32(const float* src, float* dst, int B, int C, int H, int W, int64_t stride_b,
                      int64_t stride_c, int64_t stride_h, int64_t stride_w, void* stream) {
  NAF_REQUIRE(src && dst, NAF_ERR_NULL, "pack_nhwc: NULL pointer");
  NAF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, NAF_ERR_BAD_SHAPE, "pack_nhwc: sizes must be positive");
  return launch_pack_nhwc(src, dst, B, C, H, W, stride_b, stride_c, stride_h, stride_w, C, 0,
                          static_cast<cudaStream_t>(stream));
}

int naf_pack_nhwc_slab_f32(const float* src, float* dst, int B, int C, int H, int W,
                           int64_t stride_b, int64_t stride_c, int64_t stride_h, int64_t stride_w,
                           int dst_channels, int dst_channel_offset, void* stream) {
  NAF_REQUIRE(src && dst, NAF_ERR_NULL, "pack_nhwc_slab: NULL pointer");
  NAF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, NAF_ERR_BAD_SHAPE, "pack_nhwc_slab: sizes must be positive");
  NAF_REQUIRE(dst_channel_offset >= 0 && dst_channel_offset + C <= dst_channels, NAF_ERR_BAD_SHAPE,
              "pack_nhwc_slab: slab [%d, %d) does not fit %d channels", dst_channel_offset,
              dst_channel_offset + C, dst_channels);
  return launch_pack_nhwc(src, dst, B, C, H, W, stride_b, stride_c, stride_h, stride_w,
                          dst_channels, dst_channel_offset, static_cast<cudaStream_t>(stream));
}

int naf_rope_kpool_f32(const naf_kpool_params* pp, void* stream) {
  NAF_REQUIRE(pp, NAF_ERR_NULL, "rope_kpool: NULL params");
  const naf_kpool_params& p = *pp;
  NAF_REQUIRE(p.x && (p.k_out || p.q_out), NAF_ERR_NULL, "rope_kpool: x and one output required");
  NAF_REQUIRE(p.B > 0 && p.D > 0 && p.Ho > 0 && p.Wo > 0, NAF_ERR_BAD_SHAPE,
              "rope_kpool: sizes must be positive");
  NAF_REQUIRE(!p.k_out || (p.h > 0 && p.w > 0), NAF_ERR_BAD_SHAPE, "rope_kpool: bad pooled size");
  const int nrope = (p.cos_y != nullptr) + (p.sin_y != nullptr) + (p.cos_x != nullptr) +
                    (p.sin_x != nullptr);
  NAF_REQUIRE(nrope == 0 || nrope == 4, NAF_ERR_NULL,
              "rope_kpool: rope tables must be all given or all NULL");
  if (nrope == 4)
    NAF_REQUIRE(p.rope_heads > 0 && p.D % (4 * p.rope_heads) == 0, NAF_ERR_BAD_SHAPE,
                "rope_kpool: embed_dim %% (4*num_heads) must be 0");
  NAF_REQUIRE(p.x_stride_x >= p.D, NAF_ERR_BAD_SHAPE, "rope_kpool: bad x strides");
  NAF_REQUIRE(p.rep_y >= 1 && p.rep_x >= 1 && p.Ho % p.rep_y == 0 && p.Wo % p.rep_x == 0,
              NAF_ERR_BAD_SHAPE, "rope_kpool: replication factors (%d,%d) must divide the map size",
              p.rep_y, p.rep_x);
  return launch_rope_kpool(p, static_cast<cudaStream_t>(stream));
}

int naf_xattn_bwd_f32(const naf_xattn_bwd_params* pp, void* stream) {
  NAF_REQUIRE(pp, NAF_ERR_NULL, "xattn_bwd: NULL params");
  const naf_xattn_bwd_params& p = *pp;
  NAF_REQUIRE(p.q && p.k && p.v && p.dout && p.dq && p.dk && p.dv, NAF_ERR_NULL,
              "xattn_bwd: q, k, v, dout, dq, dk and dv must be non-NULL");
  // the forward's own validation, on the forward view of these parameters
  naf_xattn_params f;
  memset(&f, 0, sizeof(f));
  f.q = p.q; f.k = p.k; f.v = p.v; f.out = p.dq;
  f.row_tap = p.row_tap; f.col_tap = p.col_tap;
  f.cos_y = p.cos_y; f.sin_y = p.sin_y; f.cos_x = p.cos_x; f.sin_x = p.sin_x;
  f.B = p.B; f.D = p.D; f.C = p.C; f.heads = p.heads; f.Ho = p.Ho; f.Wo = p.Wo; f.h = p.h; f.w = p.w; f.K = p.K;
  f.Kw = p.Kw;
  f.scale = p.scale;
  f.q_stride_b = p.q_stride_b; f.q_stride_y = p.q_stride_y; f.q_stride_x = p.q_stride_x;
  f.rep_y = p.rep_y; f.rep_x = p.rep_x;
  int rc = validate_xattn(f);
  if (rc != NAF_OK) return rc;
  NAF_REQUIRE(p.algo == NAF_ALGO_AUTO || p.algo == NAF_ALGO_GENERIC || p.algo == NAF_ALGO_CELL_SIMT || p.algo == NAF_ALGO_CELL_TC,
              NAF_ERR_UNSUPPORTED, "xattn_bwd: algo %d (auto, generic, cell_simt or cell_tc)", p.algo);
  return launch_xattn_bwd(p, static_cast<cudaStream_t>(stream));
}

int naf_rope_kpool_bwd_f32(const naf_kpool_bwd_params* pp, void* stream) {
  NAF_REQUIRE(pp, NAF_ERR_NULL, "rope_kpool_bwd: NULL params");
  const naf_kpool_bwd_params& p = *pp;
  NAF_REQUIRE(p.dx && (p.dq || p.dk), NAF_ERR_NULL, "rope_kpool_bwd: dx and one of dq / dk required");
  NAF_REQUIRE(p.B > 0 && p.D > 0 && p.Ho > 0 && p.Wo > 0 && p.D % 2 == 0, NAF_ERR_BAD_SHAPE,
              "rope_kpool_bwd: sizes must be positive, D even");
  NAF_REQUIRE(!p.dk || (p.h > 0 && p.w > 0), NAF_ERR_BAD_SHAPE, "rope_kpool_bwd: bad pooled size");
  const int nrope = (p.cos_y != nullptr) + (p.sin_y != nullptr) + (p.cos_x != nullptr) + (p.sin_x != nullptr);
  NAF_REQUIRE(nrope == 0 || nrope == 4, NAF_ERR_NULL, "rope_kpool_bwd: rope tables must be all given or all NULL");
  if (nrope == 4)
    NAF_REQUIRE(p.rope_heads > 0 && p.D % (4 * p.rope_heads) == 0, NAF_ERR_BAD_SHAPE,
                "rope_kpool_bwd: embed_dim %% (4*num_heads) must be 0");
  return launch_rope_kpool_bwd(p, static_cast<cudaStream_t>(stream));
}

int naf_xattn_select_algo(const naf_xattn_params* pp) {
  if (!pp) return -fail(NAF_ERR_NULL, "xattn: NULL params");
  int rc = validate_xattn(*pp);
  if (rc != NAF_OK) return -rc;
  return select_algo(*pp, true);
}

size_t naf_xattn_workspace_bytes(const naf_xattn_params* pp) {
  if (!pp || validate_xattn(*pp) != NAF_OK) return 0;
  // ask "would the TMA kernel take this if it had its workspace?"
  naf_xattn_params q = *pp;
  q.workspace = reinterpret_cast<void*>(uintptr_t(256));
  q.workspace_bytes = INT64_MAX;
  const char* why = "";
  if ((q.algo == NAF_ALGO_AUTO || q.algo == NAF_ALGO_CELL_TMA) && xattn_cell_tma_supported(q, &why))
    return xattn_cell_tma_workspace(q);
  return 0;
}

int naf_xattn_fwd_f32(const naf_xattn_params* pp, void* stream) {
  NAF_REQUIRE(pp, NAF_ERR_NULL, "xattn: NULL params");
  const naf_xattn_params& p = *pp;
  int rc = validate_xattn(p);
  if (rc != NAF_OK) return rc;
  const int algo = select_algo(p, false);
  if (algo < 0) return -algo;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (algo) {
    case NAF_ALGO_CELL_TMA:
      return launch_xattn_cell_tma(p, st);
    case NAF_ALGO_CELL_TCWS:
      return launch_xattn_cell_tcws(p, st);
    case NAF_ALGO_CELL_SIMT:
      return launch_xattn_cell_simt(p, st);
    case NAF_ALGO_UNION_TC:
      return launch_xattn_union_tc(p, st);
    default:
      return launch_xattn_generic(p, st);
  }
}

int naf_concat_bias_nhwc_f32(const float* a, const float* bias_a, int Ca, const float* b,
                             const float* bias_b, int Cb, float* out, int64_t npix, void* stream) {
  NAF_REQUIRE(a && b && out, NAF_ERR_NULL, "concat_bias: NULL pointer");
  NAF_REQUIRE(Ca > 0 && Cb > 0 && npix > 0, NAF_ERR_BAD_SHAPE, "concat_bias: bad sizes");
  return launch_concat_bias(a, bias_a, Ca, b, bias_b, Cb, out, npix, static_cast<cudaStream_t>(stream));
}

int naf_gn_stats_f32(const float* y, const float* bias, double* sums, int B, int64_t HW, int C, int G,
                     void* stream) {
  NAF_REQUIRE(y && sums, NAF_ERR_NULL, "gn_stats: NULL pointer");
  NAF_REQUIRE(B > 0 && HW > 0 && C > 0 && G > 0 && C % G == 0, NAF_ERR_BAD_SHAPE, "gn_stats: bad sizes");
  return launch_gn_stats(y, bias, sums, B, HW, C, G, static_cast<cudaStream_t>(stream));
}

int naf_gn_silu_apply_f32(const float* y, const float* bias, const float* gamma, const float* beta,
                          const double* sums, float* out, int B, int H, int W, int C, int G, float eps,
                          int pad, void* stream) {
  NAF_REQUIRE(y && gamma && beta && sums && out, NAF_ERR_NULL, "gn_silu_apply: NULL pointer");
  NAF_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && G > 0 && C % G == 0, NAF_ERR_BAD_SHAPE,
              "gn_silu_apply: bad sizes");
  return launch_gn_silu_apply(y, bias, gamma, beta, sums, out, B, H, W, C, G, eps, pad,
                              static_cast<cudaStream_t>(stream));
}

int naf_enc_stem_f32(const float* image, int64_t stride_b, int64_t stride_c, int64_t stride_h,
                     int64_t stride_w, const float* weight, const float* bias, float* out, float* part,
                     int B, int H, int W, int KS, void* stream) {
  NAF_REQUIRE(image && weight && out, NAF_ERR_NULL, "enc_stem: NULL pointer");
  NAF_REQUIRE(B > 0 && H > 0 && W > 0, NAF_ERR_BAD_SHAPE, "enc_stem: sizes must be positive");
  return launch_enc_stem(image, stride_b, stride_c, stride_h, stride_w, weight, bias, out, part, B, H, W,
                         KS, NAF_DTYPE_F32, static_cast<cudaStream_t>(stream));
}

int naf_enc_stem_tc_f32(const float* image, int64_t stride_b, int64_t stride_c, int64_t stride_h,
                        int64_t stride_w, const float* weight, const float* bias, float* out, float* part,
                        int B, int H, int W, int KS, void* stream) {
  NAF_REQUIRE(image && weight && out, NAF_ERR_NULL, "enc_stem_tc: NULL pointer");
  NAF_REQUIRE(B > 0 && H > 0 && W > 0, NAF_ERR_BAD_SHAPE, "enc_stem_tc: sizes must be positive");
#ifndef NAF_WITH_TC
  return fail(NAF_ERR_UNSUPPORTED, "enc_stem_tc: library built without the tensor-core path");
#else
  return launch_enc_stem_tc(image, stride_b, stride_c, stride_h, stride_w, weight, bias, out, part, B, H, W,
                            KS, NAF_DTYPE_F32, static_cast<cudaStream_t>(stream));
#endif
}

int naf_enc_gn_coef_f32(const float* part, const float* gamma, const float* beta, float* coef, int B,
                        int H, int W, float eps, void* stream) {
  NAF_REQUIRE(part && gamma && beta && coef, NAF_ERR_NULL, "enc_gn_coef: NULL pointer");
  NAF_REQUIRE(B > 0 && H > 0 && W > 0, NAF_ERR_BAD_SHAPE, "enc_gn_coef: sizes must be positive");
  return launch_enc_gn_coef(part, gamma, beta, coef, B, H, W, eps, static_cast<cudaStream_t>(stream));
}

int naf_enc_conv_pack_f32(const float* weight, void* packed, int KS, void* stream) {
  NAF_REQUIRE(weight && packed, NAF_ERR_NULL, "enc_conv_pack: NULL pointer");
  return launch_enc_conv_pack(weight, packed, KS, static_cast<cudaStream_t>(stream));
}

int naf_enc_conv_f32(const float* in, const float* coef, const void* wpacked, const float* bias,
                     float* out, int64_t out_pix_stride, int out_channel_offset, float* part, int B,
                     int H, int W, int KS, int passes, void* stream) {
  NAF_REQUIRE(in && coef && wpacked && out, NAF_ERR_NULL, "enc_conv: NULL pointer");
  NAF_REQUIRE(B > 0 && H > 0 && W > 0, NAF_ERR_BAD_SHAPE, "enc_conv: sizes must be positive");
#ifndef NAF_WITH_TC
  return fail(NAF_ERR_UNSUPPORTED, "enc_conv: library built without the tensor-core path");
#else
  return launch_enc_conv(in, coef, wpacked, bias, out, out_pix_stride, out_channel_offset, part, B, H, W,
                         KS, passes, NAF_DTYPE_F32, NAF_DTYPE_F32, static_cast<cudaStream_t>(stream));
#endif
}

int naf_enc_stem_ex(const float* image, int64_t stride_b, int64_t stride_c, int64_t stride_h, int64_t stride_w,
                    const float* weight, const float* bias, void* out, float* part, int B, int H, int W, int KS,
                    int tensor_core, int out_dtype, void* stream) {
  NAF_REQUIRE(image && weight && out, NAF_ERR_NULL, "enc_stem_ex: NULL pointer");
  NAF_REQUIRE(B > 0 && H > 0 && W > 0, NAF_ERR_BAD_SHAPE, "enc_stem_ex: sizes must be positive");
#ifndef NAF_WITH_TC
  if (tensor_core) return fail(NAF_ERR_UNSUPPORTED, "enc_stem_ex: library built without the tensor-core path");
#endif
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return tensor_core ? launch_enc_stem_tc(image, stride_b, stride_c, stride_h, stride_w, weight, bias, out, part, B, H, W, KS, out_dtype, st)
                     : launch_enc_stem(image, stride_b, stride_c, stride_h, stride_w, weight, bias, out, part, B, H, W, KS, out_dtype, st);
}

int naf_enc_conv_ex(const void* in, const float* coef, const void* wpacked, const float* bias, void* out,
                    int64_t out_pix_stride, int out_channel_offset, float* part, int B, int H, int W, int KS,
                    int passes, int in_dtype, int out_dtype, void* stream) {
  NAF_REQUIRE(in && coef && wpacked && out, NAF_ERR_NULL, "enc_conv_ex: NULL pointer");
  NAF_REQUIRE(B > 0 && H > 0 && W > 0, NAF_ERR_BAD_SHAPE, "enc_conv_ex: sizes must be positive");
#ifndef NAF_WITH_TC
  return fail(NAF_ERR_UNSUPPORTED, "enc_conv_ex: library built without the tensor-core path");
#else
  return launch_enc_conv(in, coef, wpacked, bias, out, out_pix_stride, out_channel_offset, part, B, H, W, KS, passes,
                         in_dtype, out_dtype, static_cast<cudaStream_t>(stream));
#endif
}

int naf_xattn_dump_taps_i32(int32_t* idx_out, const int32_t* row_tap, const int32_t* col_tap,
                            int Ho, int Wo, int h, int w, int K, void* stream) {
  NAF_REQUIRE(idx_out, NAF_ERR_NULL, "dump_taps: NULL output");
  NAF_REQUIRE(Ho > 0 && Wo > 0 && h > 0 && w > 0 && K >= 1 && (K & 1), NAF_ERR_BAD_SHAPE,
              "dump_taps: bad sizes");
  const bool have = row_tap && col_tap;
  NAF_REQUIRE(have || (Ho % h == 0 && Wo % w == 0 && !row_tap && !col_tap), NAF_ERR_BAD_SHAPE,
              "dump_taps: tables required for non-integer ratios");
  return launch_dump_taps(idx_out, row_tap, col_tap, Ho, Wo, h, w, K,
                          static_cast<cudaStream_t>(stream));
}

}  // extern "C"
