// Shared host/device helpers for the naf_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "naf_b200.h"

namespace naf {

// ---- host-side error plumbing (thread-local message, C status codes; no exceptions) ----------
char* last_error_buffer();
int fail(naf_status code, const char* fmt, ...);
int check_launch(const char* what);
unsigned long long launch_count(const char* prefix);

#define NAF_REQUIRE(cond, code, ...)          \
  do {                                        \
    if (!(cond)) return ::naf::fail((code), __VA_ARGS__); \
  } while (0)

// Ask for the maximum shared-memory carveout for a kernel that itself needs little shared memory.  Every
// hot kernel of the step runs with ~200 KB of dynamic shared memory; a small kernel launched between two
// of them with the default (L1-heavy) carveout makes the SMs reconfigure their L1/shared split twice,
// which drains the machine (measured: ~1 ms of idle GPU per C2 step over 21 launches).
void prefer_max_shared_impl(const void* func);   // once per (kernel, device)
template <class Kernel>
inline void prefer_max_shared(Kernel k) {
  prefer_max_shared_impl(reinterpret_cast<const void*>(k));
}

// ---- launch-time device queries, cached (the latency-bound shapes pay for every runtime call) ----
// SM count of the current device (queried once per device).
int device_sm_count();
// One-time dynamic shared-memory opt-in per (kernel, device); raises the limit when a later call
// needs more.  Function attributes are per device, so the cache is keyed on both.
cudaError_t ensure_dyn_smem_impl(const void* func, int bytes);
template <class Kernel>
inline cudaError_t ensure_dyn_smem(Kernel k, int bytes) {
  return ensure_dyn_smem_impl(reinterpret_cast<const void*>(k), bytes);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- device-side index rule shared by every kernel -------------------------------------------
// Integer-ratio window origin along one axis: first low-res index of the K-tap window of the
// low-res cell `cell` (NATTEN's shifted window composed with nearest-exact,
// for Ho = r*h): clamp(cell - K/2, 0, len - K).
__host__ __device__ __forceinline__ int window_origin(int cell, int len, int K) {
  int o = cell - (K >> 1);
  o = o < 0 ? 0 : o;
  int hi = len - K;
  return o > hi ? hi : o;
}

// Low-res index of tap t of target index i: table lookup when a tap table is given, else the
// integer-ratio rule.
__device__ __forceinline__ int tap_index(const int32_t* __restrict__ table, int i, int t, int K,
                                         int ratio, int len) {
  if (table) return table[i * K + t];
  return window_origin(i / ratio, len, K) + t;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 2^x on the SFU (ex2.approx, max rel. error 2^-22); x <= 0 in every use (softmax numerators)
__device__ __forceinline__ float fast_exp2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// Packed fp32x2 FMA (Blackwell FFMA2): d = a*b + d on both halves of a 64-bit register pair.
__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
  uint64_t dd = *reinterpret_cast<uint64_t*>(&d);
  const uint64_t aa = *reinterpret_cast<const uint64_t*>(&a);
  const uint64_t bb = *reinterpret_cast<const uint64_t*>(&b);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
  d = *reinterpret_cast<float2*>(&dd);
}

// streaming (evict-first) 16-byte global load / store for data touched exactly once
__device__ __forceinline__ float4 ldg_stream(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(float* p, const float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
// 256-bit (sm_100+) streaming global load / store: one full 32-byte sector per thread
__device__ __forceinline__ void ldg_stream8(const float* p, float (&r)[8]) {
  asm("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]),
                 "=f"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void ldg_hint8(const float* p, float (&r)[8], uint64_t policy) {
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]),
                 "=f"(r[7])
               : "l"(p), "l"(policy));
}
__device__ __forceinline__ void ldg8(const float* p, float (&r)[8]) {
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]),
                 "=f"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void stg_stream8(float* p, const float (&r)[8]) {
  asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(r[0]), "f"(r[1]),
               "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7])
               : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
inline bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; }
__device__ __forceinline__ void stg_stream2(float* p, const float2 v) {
  asm volatile("st.global.cs.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}

// ---- kernel launchers implemented in the individual .cu files --------------------------------
int launch_pack_nhwc(const float* src, float* dst, int B, int C, int H, int W, int64_t sb,
                     int64_t sc, int64_t sh, int64_t sw, int dst_channels, int dst_offset,
                     cudaStream_t st);
int launch_rope_kpool(const naf_kpool_params& p, cudaStream_t st);
int launch_xattn_generic(const naf_xattn_params& p, cudaStream_t st);
int launch_xattn_bwd(const naf_xattn_bwd_params& p, cudaStream_t st);
int launch_rope_kpool_bwd(const naf_kpool_bwd_params& p, cudaStream_t st);
bool xattn_bwd_cell_tc_supported(const naf_xattn_bwd_params& p, const char** why);
int launch_xattn_bwd_cell_tc(const naf_xattn_bwd_params& p, cudaStream_t st);
bool xattn_cell_simt_supported(const naf_xattn_params& p, const char** why);
int launch_xattn_cell_simt(const naf_xattn_params& p, cudaStream_t st);
bool xattn_cell_tcws_supported(const naf_xattn_params& p, const char** why);
int launch_xattn_cell_tcws(const naf_xattn_params& p, cudaStream_t st);
bool xattn_cell_tma_supported(const naf_xattn_params& p, const char** why);
int launch_xattn_cell_tma(const naf_xattn_params& p, cudaStream_t st);
size_t xattn_cell_tma_workspace(const naf_xattn_params& p);
bool xattn_union_tc_supported(const naf_xattn_params& p, const char** why);
int launch_xattn_union_tc(const naf_xattn_params& p, cudaStream_t st);
int launch_concat_bias(const float* a, const float* bias_a, int Ca, const float* b, const float* bias_b,
                       int Cb, float* out, int64_t npix, cudaStream_t st);
int launch_gn_stats(const float* y, const float* bias, double* sums, int B, int64_t HW, int C, int G,
                    cudaStream_t st);
int launch_gn_silu_apply(const float* y, const float* bias, const float* gamma, const float* beta,
                         const double* sums, float* out, int B, int H, int W, int C, int G, float eps,
                         int pad, cudaStream_t st);
int launch_enc_conv(const void* in, const float* coef, const void* wpack, const float* bias, void* out,
                    int64_t out_pix_stride, int out_ch_off, float* part, int B, int H, int W, int KS,
                    int passes, int in_dtype, int out_dtype, cudaStream_t st);
int launch_enc_stem(const float* image, int64_t sb, int64_t sc, int64_t sy, int64_t sx, const float* w,
                    const float* bias, void* out, float* part, int B, int H, int W, int KS, int out_dtype,
                    cudaStream_t st);
int launch_enc_stem_tc(const float* image, int64_t sb, int64_t sc, int64_t sy, int64_t sx, const float* w,
                       const float* bias, void* out, float* part, int B, int H, int W, int KS, int out_dtype,
                       cudaStream_t st);
int launch_enc_gn_coef(const float* part, const float* gamma, const float* beta, float* coef, int B, int H,
                       int W, float eps, cudaStream_t st);
int launch_enc_conv_pack(const float* w, void* packed, int KS, cudaStream_t st);
int launch_dump_taps(int32_t* idx_out, const int32_t* row_tap, const int32_t* col_tap, int Ho,
                     int Wo, int h, int w, int K, cudaStream_t st);

}  // namespace naf
