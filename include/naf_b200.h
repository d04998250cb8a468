/*
 * naf_b200 -- C ABI of the B200-native cross-scale neighbourhood-attention forward.
 *
 * This is the drop-in boundary for the hot path of valeoai/NAF.  The reference binds this path
 * through Python (PyTorch + the third-party NATTEN extension); every entry point below names the
 * reference interface it replaces.  All pointers are DEVICE pointers unless stated otherwise,
 * all tensors are fp32, all strides are in ELEMENTS.  No entry point allocates, frees or
 * synchronises; work is enqueued on `stream` (a cudaStream_t passed as void*).  Every function
 * returns a naf_status; a human-readable message for the last failure on the calling thread is
 * available from naf_last_error().
 *
 * Canonical device layouts ("pixel-major", channels innermost):
 *   guidance  x / q : (B, Ho, Wo, D)   channel stride 1, pixel strides free (>= D, 16 B aligned)
 *   keys      k     : (B, h,  w,  D)   contiguous
 *   values    v     : (B, h,  w,  C)   contiguous
 *   output    out   : (B, Ho, Wo, C)   contiguous, fp32 or bf16 (naf_xattn_params.out_dtype)  (the reference also returns pixel-major
 *                                       storage: src/layers/attentions.py:75 is a permuted view)
 *   scores          : (B, n, Ho, Wo, K*K) contiguous, tap order t_h*K + t_w
 */
#ifndef NAF_B200_H_
#define NAF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NAF_ABI_VERSION 4

#if defined(__GNUC__)
#define NAF_API __attribute__((visibility("default")))
#else
#define NAF_API
#endif

typedef enum naf_status {
  NAF_OK = 0,
  NAF_ERR_BAD_SHAPE = 1,      /* non-positive size, C or D not divisible by heads, ...          */
  NAF_ERR_UNSUPPORTED = 2,    /* valid request this build has no kernel for                      */
  NAF_ERR_WINDOW = 3,         /* kernel_size even / < 1, or kernel_size*dilation > target axis  */
  NAF_ERR_ALIGNMENT = 4,      /* pointer or stride not aligned as the layout above requires     */
  NAF_ERR_NULL = 5,           /* required pointer is NULL                                        */
  NAF_ERR_CUDA = 6            /* CUDA runtime error at launch (message in naf_last_error)       */
} naf_status;

/* ABI version of the loaded library (== NAF_ABI_VERSION of the header it was built from). */
NAF_API int naf_abi_version(void);

/* Message for the last non-OK status returned on this host thread ("" if none). */
NAF_API const char* naf_last_error(void);

/* Number of kernels this library has launched since it was loaded whose family name starts with
 * `prefix` (NULL or "" = all): "xattn", "enc_conv", "rope_kpool", "pack_nhwc", ...  Host-side
 * counter incremented after every successful launch; used by bench.py's `gpu_launches`. */
NAF_API unsigned long long naf_launch_count(const char* prefix);

/* 1 if the library was compiled with the tcgen05/TMEM tensor-core path, 0 if SIMT only. */
NAF_API int naf_has_tensor_path(void);

/* ------------------------------------------------------------------------------------------
 * Layout packing: (B, C, H, W) with arbitrary element strides -> (B, H, W, C) contiguous.
 * Replaces the `rearrange(x, "b (n d) h w -> b h w n d")` / `.to(dtype)` layout steps of
 * reference src/layers/attentions.py:49-51,59 (without the nearest-exact replication: K and V
 * stay at low resolution).
 * ---------------------------------------------------------------------------------------- */
NAF_API int naf_pack_nhwc_f32(const float* src, float* dst, int B, int C, int H, int W,
                      int64_t stride_b, int64_t stride_c, int64_t stride_h, int64_t stride_w,
                      void* stream);

/* Same, writing the C channels as the slab [dst_channel_offset, dst_channel_offset + C) of a
 * pixel-major tensor with dst_channels channels per pixel.  Two calls concatenate the two encoder
 * branches (reference src/model/naf.py:33, torch.cat) while packing them. */
NAF_API int naf_pack_nhwc_slab_f32(const float* src, float* dst, int B, int C, int H, int W,
                      int64_t stride_b, int64_t stride_c, int64_t stride_h, int64_t stride_w,
                      int dst_channels, int dst_channel_offset, void* stream);

/* ------------------------------------------------------------------------------------------
 * RoPE + key pooling pre-pass.
 *   q      = RoPE(x)                      reference src/layers/rope.py:155-174 (eval branch)
 *   k_out  = adaptive_avg_pool2d(q,(h,w)) reference src/model/naf.py:63-69,108
 * x is pixel-major with channel stride 1.  The four tables are what RoPE.rotate builds
 * (src/layers/rope.py:139-146), factored per axis: cos/sin of 2*pi*coord/period for the rows
 * (`*_y`, shape (Ho, P)) and the columns (`*_x`, shape (Wo, P)), P = (D/rope_heads)/4.
 * q_out may be NULL (keys only); when given it receives the rotated map, contiguous
 * (B,Ho,Wo,D).  If all four tables are NULL, x is taken as already rotated.
 * Pooling bins follow ATen: rows [floor(i*Ho/h), ceil((i+1)*Ho/h)).
 *
 * Replicated guidance (rep_y, rep_x >= 1): when the conv encoder ran at a resolution (Ho/rep_y,
 * Wo/rep_x) below the target, the reference's adaptive_avg_pool2d to the target size
 * (src/model/naf.py:34) is pure replication; x then holds the SOURCE map (B, Ho/rep_y,
 * Wo/rep_x, D) and target pixel (y, x) reads source pixel (y/rep_y, x/rep_x).  RoPE always uses
 * the target coordinates.  The strides describe the source map.
 * ---------------------------------------------------------------------------------------- */
typedef struct naf_kpool_params {
  const float* x;        /* (B,Ho,Wo,D) */
  float* k_out;          /* (B,h,w,D) contiguous */
  float* q_out;          /* NULL or (B,Ho,Wo,D) contiguous */
  const float* cos_y;    /* (Ho,P) */
  const float* sin_y;
  const float* cos_x;    /* (Wo,P) */
  const float* sin_x;
  int32_t B, D, Ho, Wo, h, w;
  int32_t rope_heads;    /* D % (4*rope_heads) == 0 */
  int64_t x_stride_b, x_stride_y, x_stride_x; /* elements */
  int32_t rep_y, rep_x;  /* replication factors of x (1 = x is at target resolution) */
} naf_kpool_params;

NAF_API int naf_rope_kpool_f32(const naf_kpool_params* p, void* stream);

/* ------------------------------------------------------------------------------------------
 * Cross-scale neighbourhood attention forward.
 * Replaces reference CrossAttention.forward (src/layers/attentions.py:53-75) together with
 * legacy_attention (:16-29) and the NATTEN functionals it calls (na2d_qk :20, na2d_av :24, or
 * fused na2d :72):
 *   s[t,u] = scale * <q[b,y,x,head], k[b,row_tap[y][t],col_tap[x][u],head]>
 *   p      = softmax_{t,u}(s)
 *   out[b,y,x,head] = sum_{t,u} p[t,u] * v[b,row_tap[y][t],col_tap[x][u],head]
 * row_tap (Ho,K) / col_tap (Wo,K) [(Wo,Kw) for a rectangular window, see Kw] are int32 DEVICE tables holding the composition of NATTEN's
 * shifted dilated window with F.interpolate(mode="nearest-exact").  They may
 * both be NULL iff Ho % h == 0 and Wo % w == 0: then the window of a pixel is the clamped
 * low-resolution window  clamp(y/(Ho/h) - K/2, 0, h-K) + t  shared by its whole cell.
 * If the rope tables are non-NULL, `q` holds the UN-rotated map and RoPE is applied on the fly
 * (requires rope head dim == D/heads); otherwise q is used as is.
 * `scores` (optional) receives the scaled pre-softmax logits (reference return_weights=True,
 * src/layers/attentions.py:27-28).
 * ---------------------------------------------------------------------------------------- */
typedef struct naf_xattn_params {
  const float* q;        /* (B,Ho,Wo,D) channel stride 1 */
  const float* k;        /* (B,h,w,D) contiguous */
  const float* v;        /* (B,h,w,C) contiguous */
  void* out;             /* (B,Ho,Wo,C) contiguous; float, or bf16 when out_dtype == NAF_DTYPE_BF16 */
  float* scores;         /* NULL or (B,heads,Ho,Wo,K*K) */
  const int32_t* row_tap;/* NULL or (Ho,K) */
  const int32_t* col_tap;/* NULL or (Wo,K) */
  const float* cos_y;    /* NULL or (Ho, D/heads/4) */
  const float* sin_y;
  const float* cos_x;    /* NULL or (Wo, D/heads/4) */
  const float* sin_x;
  int32_t B, D, C, heads, Ho, Wo, h, w, K;
  float scale;           /* (D/heads)^-0.5 in the reference (src/layers/attentions.py:46) */
  int64_t q_stride_b, q_stride_y, q_stride_x; /* elements */
  int32_t algo;          /* NAF_ALGO_* ; AUTO picks the fastest kernel that supports the request */
  int32_t rep_y, rep_x;  /* q is a replicated source map (B,Ho/rep_y,Wo/rep_x,D); see kpool */
  int32_t out_dtype;     /* NAF_DTYPE_F32 (0) or NAF_DTYPE_BF16 (1): element type of `out` only.  bf16 is what the
                            reference returns under torch.autocast(bfloat16) (train.py:120, denoising.py:209);
                            arithmetic stays fp32, the result is rounded to nearest-even on the final store.
                            Served by the pipelined tensor-core kernels and the generic kernel. */
  void* workspace;       /* scratch the TMA kernel keeps its fp16 hi/lo K and V planes in (>= naf_xattn_workspace_bytes,
                            any alignment; contents undefined before and after the call); NULL: AUTO skips that kernel */
  int64_t workspace_bytes;
  int32_t q_dtype, k_dtype, v_dtype; /* NAF_DTYPE_F32 (0) or NAF_DTYPE_BF16 (1): element type of q, k and v (same layouts,
                            strides in elements).  bf16 is what the reference feeds under torch.autocast(bfloat16)
                            (train.py:120); it is read as it is by the TMA kernel (NAF_ALGO_CELL_TMA) -- the other
                            kernels take fp32 only and refuse (callers widen for them). */
  int32_t Kw;            /* window WIDTH (taps along x) of a rectangular window, NATTEN's kernel_size=(K, Kw)
                            (src/layers/attentions.py:20,24 pass the tuple through); 0 = square, Kw = K.  K is then the
                            window HEIGHT, row_tap is (Ho,K), col_tap (Wo,Kw), the scores (…,K*Kw) in tap order
                            t_h*Kw + t_w.  Rectangular windows run on the generic kernel (AUTO picks it). */
} naf_xattn_params;

enum { NAF_DTYPE_F32 = 0, NAF_DTYPE_BF16 = 1, NAF_DTYPE_F16 = 2 /* encoder activations only */ };

enum { NAF_ALGO_AUTO = 0, NAF_ALGO_GENERIC = 1, NAF_ALGO_CELL_SIMT = 2,
       NAF_ALGO_CELL_TC = 3 /* non-pipelined tensor-core cell kernel: the forward one was removed in ABI v3 (forward
                               requests fail with NAF_ERR_UNSUPPORTED); naf_xattn_bwd_f32 has one (naf_xattn_bwd_tc.cu) */,
       NAF_ALGO_CELL_TCWS = 4 /* warp-specialised persistent tcgen05 pipeline, windows converted in the kernel */,
       NAF_ALGO_CELL_TMA = 5  /* the same pipeline fed by tensor-map TMA: pre-split K / V planes, window boxes,
                                  single-pass wide value heads, tensor stores (needs `workspace`) */,
       NAF_ALGO_UNION_TC = 6  /* tensor-core path for tap tables (non-integer ratios), ratio 1 and tiny cells: dense
                                  attention of an 8 x 16 pixel tile against the union of its windows under a
                                  multiplicity mask, online softmax over 128-cell chunks (no score output) */ };

NAF_API int naf_xattn_fwd_f32(const naf_xattn_params* p, void* stream);

/* Bytes of `workspace` the fastest kernel for these parameters wants (0 if none does).  The C side never
 * allocates: the caller passes device scratch of at least this size in naf_xattn_params.workspace. */
NAF_API size_t naf_xattn_workspace_bytes(const naf_xattn_params* p);

/* ------------------------------------------------------------------------------------------
 * Backward of the cross-scale neighbourhood attention.
 * Replaces what autograd does to reference legacy_attention (src/layers/attentions.py:16-29) through
 * NATTEN's backward functionals, as train.py:136 and test/backward_speed.py:51-64 run it.
 * q, k, v, tables, strides, rep and scale exactly as passed to the forward (nothing else is saved for
 * backward: the probabilities are recomputed); dout is dL/dout, (B,Ho,Wo,C) contiguous fp32.  Outputs:
 *   dq (B,Ho,Wo,D) contiguous : gradient w.r.t. the ROTATED queries (= w.r.t. q when no rope tables);
 *   dk (B,h,w,D), dv (B,h,w,C): zeroed by the call, then accumulated (scatter-add over the windows).
 * algo: NAF_ALGO_AUTO (tensor-core cell kernel for integer ratios with 64-wide heads and cells of >= 64 pixels,
 * else the fp32 cell kernel, else generic), NAF_ALGO_GENERIC, NAF_ALGO_CELL_SIMT, NAF_ALGO_CELL_TC (the tensor-core
 * cell kernel: five tcgen05 GEMMs per 128-pixel tile, window gradients accumulated in TMEM over the whole cell).
 * ---------------------------------------------------------------------------------------- */
typedef struct naf_xattn_bwd_params {
  const float* q;
  const float* k;
  const float* v;
  const float* dout;
  float* dq;
  float* dk;
  float* dv;
  const int32_t* row_tap;
  const int32_t* col_tap;
  const float* cos_y;
  const float* sin_y;
  const float* cos_x;
  const float* sin_x;
  int32_t B, D, C, heads, Ho, Wo, h, w, K;
  float scale;
  int64_t q_stride_b, q_stride_y, q_stride_x; /* elements */
  int32_t algo;
  int32_t rep_y, rep_x;
  int32_t Kw;            /* window width of a rectangular window (0 = square), as in naf_xattn_params (ABI v4) */
  int32_t reserved_;
} naf_xattn_bwd_params;

NAF_API int naf_xattn_bwd_f32(const naf_xattn_bwd_params* p, void* stream);

/* Backward of naf_rope_kpool_f32 (reference: autograd through RoPE.forward, src/layers/rope.py:155-174,
 * and adaptive_avg_pool2d in KeyEncoder.forward, src/model/naf.py:63-69):
 *   dx = R^T ( dq + sum over the pooling bins holding the pixel of dk_bin / |bin| )
 * dq (B,Ho,Wo,D) gradient w.r.t. the rotated map (NULL = zero), dk (B,h,w,D) gradient w.r.t. the pooled
 * keys (NULL = none), dx (B,Ho,Wo,D) contiguous, may alias dq.  Tables as in the forward (all NULL: no
 * rotation).  x is taken at the target resolution (replication factors 1). */
typedef struct naf_kpool_bwd_params {
  const float* dq;
  const float* dk;
  float* dx;
  const float* cos_y;
  const float* sin_y;
  const float* cos_x;
  const float* sin_x;
  int32_t B, D, Ho, Wo, h, w;
  int32_t rope_heads;
} naf_kpool_bwd_params;

NAF_API int naf_rope_kpool_bwd_f32(const naf_kpool_bwd_params* p, void* stream);

/* Which NAF_ALGO_* would AUTO select for these parameters (no launch).  Negative = -naf_status. */
NAF_API int naf_xattn_select_algo(const naf_xattn_params* p);

/* Test hook for the "bit-exact neighbourhood index" requirement: writes, for every target pixel,
 * the K*K linear low-res cell indices (row*w + col) the attention kernels gather from, using the
 * SAME device index code as the kernels.  idx_out: (Ho, Wo, K*K) int32.  Tables as above. */
NAF_API int naf_xattn_dump_taps_i32(int32_t* idx_out, const int32_t* row_tap, const int32_t* col_tap,
                            int Ho, int Wo, int h, int w, int K, void* stream);

/* ------------------------------------------------------------------------------------------
 * Guidance-encoder glue, pixel-major (B,H,W,C) fp32 activations.
 * Replaces, per EncBlock half (reference src/layers/convolutions.py:55-67): nn.GroupNorm
 * (statistics + affine), nn.SiLU, the reflect padding of the following Conv2d and the bias add
 * of the PRECEDING Conv2d (`bias`, may be NULL: the convolutions are then run without bias).
 *   naf_gn_stats_f32      : sums[b][g] = { sum(y+bias), sum((y+bias)^2) } in double (zeroed here)
 *   naf_gn_silu_apply_f32 : out = silu(((y+bias) - mean) * rstd * gamma + beta), biased variance,
 *                           written into a (B, H+2*pad, W+2*pad, C) tensor with reflect padding
 *                           (pad = 0 or 1, PyTorch "reflect": no edge repeat).
 * ---------------------------------------------------------------------------------------- */
/* torch.cat([a, b], dim=channels) of two pixel-major tensors (npix pixels) fused with the bias adds
 * of the convolutions that produced them (reference src/model/naf.py:33). bias_* may be NULL. */
NAF_API int naf_concat_bias_nhwc_f32(const float* a, const float* bias_a, int Ca, const float* b,
                                     const float* bias_b, int Cb, float* out, int64_t npix, void* stream);
NAF_API int naf_gn_stats_f32(const float* y, const float* bias, double* sums, int B, int64_t HW,
                             int C, int G, void* stream);
NAF_API int naf_gn_silu_apply_f32(const float* y, const float* bias, const float* gamma,
                                  const float* beta, const double* sums, float* out, int B, int H,
                                  int W, int C, int G, float eps, int pad, void* stream);

/* ------------------------------------------------------------------------------------------
 * Guidance conv encoder on the tensor core: the whole `encoder()` stack of the
 * reference (src/layers/convolutions.py:70-95; 128 channels, 8 groups, SiLU, reflect padding, no
 * residual -- the NAF defaults, src/model/naf.py:26-27) with every activation crossing HBM twice
 * per layer instead of five times.  Activations are pixel-major (B,H,W,128) fp32.
 *
 *   naf_enc_stem_f32      : nn.Conv2d(3, 128, KS, padding=KS/2, padding_mode="reflect") + bias on
 *                           an image (B,3,H,W) of any element strides (convolutions.py:72-79);
 *                           also emits per-tile partial sums {sum, sumsq} per GroupNorm group of
 *                           its output: part (B, ceil(H/16)*ceil(W/8), 8, 2) fp32 (NULL: skipped).
 *   naf_enc_gn_coef_f32   : nn.GroupNorm(8, 128, eps) statistics from those partial sums, reduced
 *                           in double in a fixed order, folded with the affine parameters:
 *                           coef (B,128,2) = { rstd*gamma, beta - mean*rstd*gamma } (biased var).
 *   naf_enc_conv_pack_f32 : Conv2d weight (128,128,KS,KS) fp32 -> the fp16 hi/lo operand image the
 *                           conv kernel streams: KS*KS*65536 bytes.
 *   naf_enc_conv_f32      : EncBlock half  GroupNorm -> SiLU -> Conv2d(128,128,KS, reflect) + bias
 *                           (convolutions.py:55-67): out = conv(silu(in*coef.scale + coef.shift)).
 *                           `out` is a 128-channel slab [out_channel_offset, +128) of a pixel-major
 *                           tensor with out_pix_stride floats per pixel (the last layer of each
 *                           branch writes straight into the concatenated guidance map,
 *                           src/model/naf.py:33).  `part` as for the stem (for the next GroupNorm).
 *                           passes = 1: operands rounded to fp16 (10-bit mantissa: the precision
 *                           class of the TF32 convolutions PyTorch runs by default, i.e. the
 *                           reference with torch.backends.cudnn.allow_tf32 = True);
 *                           passes = 3: split-fp16 (hi*hi + lo*hi + hi*lo), fp32-class accuracy
 *                           (allow_tf32 = False).  KS = 1 or 3; H, W >= 2 when KS = 3.
 * ---------------------------------------------------------------------------------------- */
NAF_API int naf_enc_stem_f32(const float* image, int64_t stride_b, int64_t stride_c, int64_t stride_h,
                             int64_t stride_w, const float* weight, const float* bias, float* out,
                             float* part, int B, int H, int W, int KS, void* stream);
/* The same stem as one 128x128x{16,32} tensor-core GEMM per 16x8 pixel tile (im2col rows built in shared
 * memory, operands rounded to fp16: the TF32 precision class, see naf_enc_conv_f32 passes = 1).  HBM bound
 * instead of FMA bound; identical arguments and partial-sum layout. */
NAF_API int naf_enc_stem_tc_f32(const float* image, int64_t stride_b, int64_t stride_c, int64_t stride_h,
                                int64_t stride_w, const float* weight, const float* bias, float* out,
                                float* part, int B, int H, int W, int KS, void* stream);
NAF_API int naf_enc_gn_coef_f32(const float* part, const float* gamma, const float* beta, float* coef,
                                int B, int H, int W, float eps, void* stream);
NAF_API int naf_enc_conv_pack_f32(const float* weight, void* packed, int KS, void* stream);
NAF_API int naf_enc_conv_f32(const float* in, const float* coef, const void* wpacked, const float* bias,
                             float* out, int64_t out_pix_stride, int out_channel_offset, float* part,
                             int B, int H, int W, int KS, int passes, void* stream);

/* The same three entry points with the element type of the ACTIVATIONS between the layers made explicit
 * (NAF_DTYPE_F32 or NAF_DTYPE_F16).  fp16 activations are part of the 1-pass (TF32) class only: that class
 * rounds the convolution's input operand to fp16 anyway, so storing the previous layer's output as fp16
 * halves the HBM bytes of every layer; GroupNorm statistics are still taken from the fp32 accumulators.
 * `tensor_core` selects naf_enc_stem_tc_f32's GEMM stem (1) or naf_enc_stem_f32's SIMT stem (0). */
NAF_API int naf_enc_stem_ex(const float* image, int64_t stride_b, int64_t stride_c, int64_t stride_h,
                            int64_t stride_w, const float* weight, const float* bias, void* out, float* part,
                            int B, int H, int W, int KS, int tensor_core, int out_dtype, void* stream);
NAF_API int naf_enc_conv_ex(const void* in, const float* coef, const void* wpacked, const float* bias, void* out,
                            int64_t out_pix_stride, int out_channel_offset, float* part, int B, int H, int W,
                            int KS, int passes, int in_dtype, int out_dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NAF_B200_H_ */
