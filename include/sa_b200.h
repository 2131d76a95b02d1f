/*
 * sa_b200.h -- C ABI of the B200-native (sm_100a) kernels behind StableAnimator's denoising hot
 * path (UNetSpatioTemporalConditionModel.forward inside InferenceAnimationPipeline's Euler loop).
 *
 * Conventions
 *   - every entry point is `extern "C"`, takes plain device pointers + sizes + a cudaStream_t
 *     (passed as void*), launches asynchronously on that stream, never allocates, never syncs;
 *   - returns 0 on success, a negative SA_ERR_* code otherwise; sa_last_error() gives the text;
 *   - activations are bf16, channel-innermost ("NHWC": [frames, h, w, C] == tokens [frames*h*w, C]);
 *     weights are bf16 K-major ([out, in]); biases / embeddings / statistics are fp32;
 *   - `act_f16` (argument or struct field) = 1 switches the 16-bit tensors of a call -- activations AND the weights
 *     that multiply them on the tensor core -- to IEEE fp16, the reference's default inference dtype
 *     (inference_basic.py:486-503); the skinny fp32-activation kernels keep bf16 weights in both modes.
 *
 * Each declaration cites the reference call site(s) it replaces (paths relative to the reference
 * checkout, e.g. animation/modules/unet.py:1633).  The reference is pure PyTorch, so the
 * "FFI a maintainer would bind" is the ctypes stub shown in INTEGRATION.md.
 */
#ifndef SA_B200_H_
#define SA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SA_OK 0
#define SA_ERR_INVALID (-1)  /* bad argument / unsupported shape */
#define SA_ERR_CUDA (-2)     /* CUDA runtime / driver error */
#define SA_ERR_TMA (-3)      /* tensor-map encode failed */

int sa_version(void);
const char* sa_last_error(void);

/* ------------------------------------------------------------------------------------------
 * Tapped implicit GEMM on tcgen05/TMEM fed by TMA (the one tensor-core kernel family).
 *
 *   out[m, n] = epilogue( sum_tap sum_c A[pos(m) + tap_off[tap], c] * W[n, tap*(c0+c1) + c] )
 *
 * A is one or two channel-innermost bf16 tensors laid over a position grid d1 x d2 x d3
 * (d1 fastest; m = (p3*d2 + p2)*d1 + p1).  Out-of-grid taps read zeros (conv zero padding).
 *   - Linear layers  (unet.py:576,617,744-749,842,865,1636,1718): ntaps = 1, grid = (M,1,1)
 *   - 3x3 conv, s=1  (unet.py:1633,1642,1779; skip-concat torch.cat at :1347,1433 read as two
 *                     sources instead of being materialised): ntaps = 9, grid = (w, h, frames)
 *   - 1x1 shortcut   (unet.py:1649): ntaps = 1, two sources
 *   - temporal conv  (nn.Conv3d k=(3,1,1), unet.py:1711,1725): ntaps = 3, grid = (S, F, B)
 * Epilogue, in this order (every piece optional):
 *   v = acc * row_scale[g] + bias[n] + row_bias[g, n]          g = m / rows_per_group
 *   if geglu: v = v_value * gelu_erf(v_gate)   (GEGLU, unet.py:868-870; W rows interleaved in
 *                                               64-row blocks: [value block | gate block])
 *   v = v * out_scale + residual[m, n]
 *   (f16_from_col > 0: columns >= f16_from_col are stored as saturating IEEE fp16 instead of bf16 -- the V
 *    block of the spatial-attention q|k|v projection, which sa_attn_spatial consumes in half precision)
 *   if blend_x: v = alpha * blend_x[m, n] + (1 - alpha) * v    (AlphaBlender, unet.py:473)
 * ------------------------------------------------------------------------------------------ */
typedef struct sa_gemm_args {
  const void* a0;        /* bf16 source 0 */
  const void* a1;        /* bf16 source 1 (channel concat after source 0) or NULL */
  int32_t c0, c1;        /* channels taken from each source; multiples of 64; c1 = 0 if unused */
  int32_t lda0, lda1;    /* elements between consecutive dim-1 positions of each source */
  int32_t d1, d2, d3;    /* position grid, d1 fastest; M = d1*d2*d3 */
  int32_t ntaps;         /* 1, 3 or 9 */
  int32_t tap_off[27];   /* (o1,o2,o3) per tap */
  const void* w;         /* bf16 [N (padded to >= tile), ntaps*(c0+c1)] */
  int32_t n;             /* output columns produced by the MMA (before GEGLU halving); n % 8 == 0 */
  const float* bias;     /* [n] or NULL */
  const float* row_bias; /* [groups, ldrb] (first n_out columns) or NULL */
  int32_t ldrb;          /* row stride of row_bias in floats (0 -> n_out) */
  const float* row_scale;/* [groups] or NULL */
  int32_t rows_per_group;
  const void* residual;  /* bf16 [M, ldr] or NULL */
  int32_t ldr;
  const void* blend_x;   /* bf16 [M, ldx] or NULL */
  int32_t ldx;
  float alpha;
  int32_t geglu;         /* 1: out has n/2 columns */
  void* out;             /* bf16 (or fp32 if out_f32) [M, ldc] */
  int32_t ldc;
  int32_t out_f32;
  int32_t tile_n;        /* 0 = auto; else one of 32, 64, 128, 160, 256 */
  float out_scale;       /* 0 = 1.0; scales (acc*row_scale + bias + row_bias [geglu]) before residual */
  int32_t f16_from_col;  /* > 0: output columns >= this (multiple of 32) are stored as fp16; 0 = all bf16 */
  /* Fused GroupNorm statistics of the OUTPUT (the tensor the next nn.GroupNorm normalises, unet.py:1632-1643, */
  /* :1709-1731, :575): when gn_stats != NULL the epilogue adds Sum / Sum^2 of the bf16-rounded results into    */
  /* gn_stats fp64 [M / gn_rows_per_stat, n_out / gn_cpg, 2] (zeroed by the call unless gn_stats_zeroed) -- exactly what the statistics */
  /* pass of sa_groupnorm would read back from HBM; consume them with sa_groupnorm_apply.  bf16 output only      */
  /* (no geglu / out_f32 / f16_from_col), n % tile == 0, M % gn_rows_per_stat == 0, n_out % gn_cpg == 0.         */
  void* gn_stats;
  int32_t gn_cpg;            /* channels per group (n_out / 32 for nn.GroupNorm(32, n_out)) */
  int32_t gn_rows_per_stat;  /* rows sharing one statistic: h*w (per frame) or F*h*w (whole tile) */
  /* Strided output grid: element strides of `out` along grid dims 2 and 3 (0 = dense: ldc*d1, ldc*d1*d2).  Lets one */
  /* launch write every other pixel of a larger image -- the four output phases of Upsample2D (nearest x2 followed   */
  /* by the 3x3 conv, unet.py:1789-1807) computed as 2x2-tap convolutions on the low-resolution input.  bf16 output, */
  /* n % tile == 0, no residual / blend_x / gn_stats.                                                                */
  int64_t ldc2, ldc3;
  /* Strided input grid: 2 = the sources are sampled at every 2nd position along grid dims 1 and 2 (their extents are */
  /* 2*d1 x 2*d2 x d3; tap offsets are in INPUT positions): the stride-2 3x3 conv of Downsample2D (unet.py:1459-1483) */
  /* straight from the full-resolution tensor through TMA element strides, no im2col.  0 / 1 = dense.                */
  int32_t in_stride;
  /* 16-bit format of a0 / a1 / w / residual / blend_x / out: 0 = bf16, 1 = IEEE fp16 (the reference's default inference  */
  /* dtype, inference_basic.py:486); fp16 stores saturate at +-65504.  f16_from_col is implied for every column.        */
  int32_t act_f16;
  /* 1: the caller has already zeroed gn_stats (e.g. one fill for all the statistics buffers of a UNet call), so the call */
  /* issues no memset of its own in front of the kernel; 0: gn_stats is zeroed by the call.                              */
  int32_t gn_stats_zeroed;
} sa_gemm_args;

int sa_gemm_tap(const sa_gemm_args* args, void* stream);
/* tile_n the heuristic would pick (weights for GEGLU must be interleaved with this tile) */
int sa_gemm_pick_tile(int32_t m, int32_t n, int32_t geglu);

/* ------------------------------------------------------------------------------------------
 * Fused transformer feed-forward (one tcgen05 kernel; the [M, 4C] hidden tensor never leaves the SM):
 *
 *   xr  = bf16(x + addv[m / rows_per_group])                       (addv optional)
 *   out = W2 . GEGLU(W1 . LayerNorm(xr) + b1) + b2 + xr            (GEGLU: value * gelu_erf(gate))
 *   if blend_x: out = alpha * blend_x + (1 - alpha) * out          (AlphaBlender, unet.py:473)
 *
 * Replaces nn.LayerNorm -> FeedForward(GEGLU, Linear) -> residual add of BasicTransformerBlock.forward
 * (unet.py:521-523, :553-555) and TemporalBasicTransformerBlock.forward (:892-897 + :936-940 ff_in; :916-917 +
 * :949-953 ff; the additive vectors are the frame-position embedding :686 and the constant attn2 output :945-947),
 * with GEGLU unet.py:859-870, FeedForward :827-857 and the time_mixer blend :693-697, for C in {64, 128, 256, 320}.
 * w1: bf16 [8C, C], rows interleaved in 64-row blocks [value | gate] (as for sa_gemm_tap geglu), b1: fp32 [8C] in the
 * same order; w2: bf16 [C, 4C]; b2, gamma, beta: fp32 [C]; x, blend_x, out: bf16 [M, ld*].
 * ------------------------------------------------------------------------------------------ */
typedef struct sa_ff_args {
  const void* x;          /* bf16 [m, ldx]: LayerNorm input and residual */
  int32_t ldx;
  int64_t m;              /* rows */
  int32_t c;              /* channels: 64, 128, 256 or 320 */
  const float* addv;      /* fp32 [m / rows_per_group, c] added to x before everything else, or NULL */
  int32_t rows_per_group;
  const float* gamma;     /* LayerNorm weight / bias / eps */
  const float* beta;
  float eps;
  const void* w1;
  const float* b1;
  const void* w2;
  const float* b2;
  const void* blend_x;    /* bf16 [m, ldbx] or NULL */
  int32_t ldbx;
  float alpha;
  void* out;              /* bf16 [m, ldc] */
  int32_t ldc;
  int32_t act_f16;        /* 16-bit format of x / w1 / w2 / blend_x / out: 0 = bf16, 1 = IEEE fp16 */
} sa_ff_args;

int sa_ff_fused(const sa_ff_args* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * Spatial self-attention, head_dim 64 (tcgen05 flash attention).
 * Replaces xformers.ops.memory_efficient_attention in AnimationAttnProcessor.__call__
 * (unet.py:1893-1905).  qkv: 16-bit [frames*S, ld], column blocks q | k | v of width C = heads*64
 * (the fused to_q/to_k/to_v projection output): q and k are bf16, v is IEEE fp16 (sa_gemm_tap with
 * f16_from_col = 2C writes it that way) because the probabilities P in (0, 256] are handed to the
 * tensor core as fp16 (11-bit significand) and kind::f16 needs A and B in the same format;
 * out: bf16 [frames*S, ldo].
 * ------------------------------------------------------------------------------------------ */
int sa_attn_spatial(const void* qkv, int32_t ld, int32_t frames, int32_t S, int32_t C,
                    int32_t heads, void* out, int32_t ldo, int32_t act_f16, void* stream);

/* Temporal self-attention over the F <= 16 frames of a tile, per (pixel, head).
 * Replaces XFormersAttnProcessor.__call__ self-attention case (unet.py:1830-1848) together with the
 * [B*F,S,C] <-> [B*S,F,C] permutes of TemporalBasicTransformerBlock.forward (:931-933,:954-956).
 * qkv: bf16 [B*F*S, ld] (row = (b*F + f)*S + s), out: bf16 [B*F*S, ldo]. */
int sa_attn_temporal(const void* qkv, int32_t ld, int32_t B, int32_t F, int32_t S, int32_t C,
                     int32_t heads, void* out, int32_t ldo, int32_t act_f16, void* stream);

/* ID cross-attention against the num_tokens (=4) face tokens + per-frame Sum/Sum^2 of the result.
 * Replaces AnimationIDAttnNormalizedProcessor.__call__ lines unet.py:1999-2019 (id_to_k / id_to_v
 * come from sa_small_linear).  q: bf16 [B*F*S, ldq]; ik, iv: fp32, token t of batch b at
 * ik + b*kv_batch_stride + t*kv_tok_stride (C contiguous floats); out: bf16
 * [B*F*S, ldo]; stats: fp64 [B*F, 2] workspace (zeroed by the call). */
int sa_attn_id(const void* q, int32_t ldq, const float* ik, const float* iv, int64_t kv_tok_stride,
               int64_t kv_batch_stride, int32_t B, int32_t F, int32_t S, int32_t C, int32_t heads,
               int32_t num_tokens, void* out, int32_t ldo, void* stats, int32_t act_f16, void* stream);

/* Distribution alignment folded through to_out (unet.py:2018-2024): per frame
 *   row_scale[f]   = std_lat / (std_ip + 1e-5)          (torch.std: unbiased, N = S*C)
 *   row_bias[f, n] = (W.v_clip)[b, n] + bias[n] + (mean_lat - row_scale[f]*mean_ip) * rowsum(W)[n]
 * so that to_out(h + aligned(ip)) == row_scale * (ip @ W^T) + row_bias, with h == v_clip broadcast
 * (softmax over the single CLIP key).  vclip, wv: fp32 [B, C]; w_rowsum, bias: fp32 [C]. */
int sa_attn_id_finalize(const void* stats, const float* vclip, int64_t vclip_batch_stride,
                        const float* wv, const float* w_rowsum, const float* bias, int32_t B,
                        int32_t F, int32_t S, int32_t C, float* row_scale, float* row_bias,
                        void* stream);

/* GroupNorm(32 groups) (+SiLU): statistics over rows_per_stat rows x (C/32) channels, fp64 sums.
 * nn.GroupNorm + nn.SiLU at unet.py:171-173,575,1632,1640,1643 (rows_per_stat = h*w: per frame)
 * and :1709,1722,1731 (rows_per_stat = F*h*w: TemporalResnetBlock, whole tile).  x0 (+ optional x1,
 * channel concat = torch.cat skip input :1347,1433) bf16 [ns*rows_per_stat, ld*]; out bf16
 * [.., c0+c1]; stats_ws: fp64 [ns, 32, 2]. */
/* sa_groupnorm_apply: the normalise (+SiLU) pass alone, on statistics already accumulated in stats_ws by the
 * epilogue of the sa_gemm_tap launch that produced x0 (sa_gemm_args.gn_stats); same arguments. */
int sa_groupnorm(const void* x0, const void* x1, int32_t c0, int32_t c1, int32_t ld0, int32_t ld1,
                 int64_t rows_per_stat, int32_t ns, const float* gamma, const float* beta, float eps,
                 int32_t do_silu, void* stats_ws, void* out, int32_t ldo, int32_t act_f16, void* stream);
int sa_groupnorm_apply(const void* x0, const void* x1, int32_t c0, int32_t c1, int32_t ld0, int32_t ld1,
                       int64_t rows_per_stat, int32_t ns, const float* gamma, const float* beta, float eps,
                       int32_t do_silu, const void* stats_ws, void* out, int32_t ldo, int32_t act_f16, void* stream);

/* LayerNorm over C (unet.py:499,510,521,892,898,907,916), eps 1e-5:
 *   y = LN(x + addv[row / rows_per_group]); if x_out: x_out = x + addv (the new residual stream).
 * addv fp32 [groups, C] or NULL (frame position embedding :686, temporal attn2 constant :946-947). */
int sa_layernorm(const void* x, int32_t ldx, int64_t rows, int32_t c, const float* addv,
                 int32_t rows_per_group, void* x_out, int32_t ldxo, const float* gamma,
                 const float* beta, float eps, void* y, int32_t ldy, int32_t act_f16, void* stream);

/* Skinny linear (m <= a few dozen rows), fp32 activations, bf16 weights [n, k], fp32 out:
 *   out (+)= act_out(act_in(x) @ W^T + b); silu_out: 0 none, 1 SiLU, 2 exact GELU.  TimestepEmbedding (unet.py:429-436), time_emb_proj(silu(.))
 *   (:1669-1670,:1754-1755), context K/V projections (:1974-1975,:1999-2000,:1834-1835). */
int sa_small_linear(const float* x, int32_t ldx, int32_t m, const void* w, int32_t k, int32_t n,
                    const float* bias, int32_t silu_in, int32_t silu_out, int32_t accumulate,
                    float* out, int32_t ldo, void* stream);

/* FusionFaceId (animation/modules/id_encoder.py:102-148; face tokens of the conditioning, once per video).
 * sa_small_linear above does its Linear layers (silu_out = 2 selects the exact GELU of nn.GELU, :60,:118).
 * LayerNorm over c on m fp32 rows (:24-25,:59,:78,:125). */
int sa_layernorm_f32(const float* x, int32_t ldx, int32_t m, int32_t c, const float* gamma, const float* beta,
                     float eps, float* y, int32_t ldy, void* stream);
/* PerceiverAttention core (:43-55): out[b, l, h*64+d] = softmax_n(q k^T / 8) v in fp32, head_dim 64, n <= 32.
 * q fp32 [b, l, ldq]; kv fp32 [b, n, ldkv] with k at column h*64 and v at column voff + h*64. */
int sa_perceiver_attn(const float* q, int32_t ldq, const float* kv, int32_t ldkv, int32_t voff, int32_t b,
                      int32_t l, int32_t n, int32_t heads, float* out, int32_t ldo, void* stream);

/* Temporal VAE decoder (animation/modules/vae.py:387-479).  Its convolutions, GroupNorms, the
 * temporal residual blocks and the q/k/v/out projections run on sa_gemm_tap / sa_groupnorm / sa_conv_direct /
 * sa_upsample2x above; two helpers complete it:
 * row softmax of the single-head mid-block attention (vae.py:92-94): scores fp32 [m, n] (Q K^T GEMM with
 * out_f32) -> probabilities bf16 [m, n], p = softmax(scale * s). */
int sa_softmax_rows(const float* s, int64_t lds, int32_t m, int32_t n, float scale, void* p, int64_t ldp,
                    void* stream);
/* time_conv_out (vae.py:423-429,470-475): Conv3d(c -> c, (3,1,1), zero padding over the f frames of each of the b
 * clips).  x fp32 [b*f*hw, ldx] channel-innermost (first c columns) -> out fp32 NCHW [b*f, c, hw]; w [c, c, 3]. */
int sa_time_conv_out(const float* x, int32_t ldx, int32_t b, int32_t f, int32_t hw, int32_t c, const float* w,
                     const float* bias, float* out, void* stream);

/* get_timestep_embedding (unet.py:365-390): out[i] = [cos(t_i f) | sin(t_i f)], fp32 [n, dim]. */
int sa_sinusoid(const float* t, int32_t n, int32_t dim, float* out, void* stream);

/* Direct NHWC convolution for tiny channel counts: conv_in (unet.py:69,283) with the PoseNet
 * residual add (:284-286; `add` applies to images n >= add_from) and the PoseNet conv stack
 * (pose_net.py:14-37).  wt: bf16 [k*k*cin, cout_pad]; out = (act(conv + bias)) * scale (+ add). */
int sa_conv_direct(const void* x, int32_t n, int32_t h, int32_t w, int32_t cin, const void* wt,
                   const float* bias, int32_t cout, int32_t cout_pad, int32_t ksz, int32_t stride,
                   int32_t pad, int32_t silu_out, float scale, const void* add, int32_t add_from,
                   void* out, int32_t act_f16, void* stream);

/* F.interpolate(scale_factor=2, mode="nearest") on NHWC bf16 (unet.py:1799). */
int sa_upsample2x(const void* in, int32_t n, int32_t h, int32_t w, int32_t c, void* out, void* stream);
/* im2col of the 3x3 stride-2 pad-1 Downsample2D conv (unet.py:1474): out bf16 [n*h/2*w/2, 9*c]. */
int sa_im2col_s2(const void* in, int32_t n, int32_t h, int32_t w, int32_t c, void* out, void* stream);
/* same gather with selectable padding: pad = 1 as above; pad = 0 is the VAE encoder's Downsample2D
 * (F.pad(x, (0,1,0,1)) then a padding-0 stride-2 conv, vae.py:731-741). */
int sa_im2col_s2_pad(const void* in, int32_t n, int32_t h, int32_t w, int32_t c, int32_t pad, void* out,
                     void* stream);
/* im2col of conv_in (3x3, pad 1, 8 input channels; unet.py:69,283): in bf16 [n, h, w, 8] ->
 * out bf16 [n*h*w, 128] = 72 gathered values (ky, kx, ci) + 56 zeros, the K = 128 operand of sa_gemm_tap. */
int sa_im2col3x3_c8(const void* in, int32_t n, int32_t h, int32_t w, void* out, void* stream);

/* Boundary layout conversion: reference tensors are NCHW, the kernels are NHWC. */
int sa_nchw_to_nhwc(const void* src, int32_t src_is_bf16, int32_t n, int32_t c, int32_t hw, void* dst,
                    int32_t act_f16, void* stream);
int sa_nhwc_to_nchw_f32(const float* src, int32_t lds, int32_t n, int32_t c, int32_t hw, float* dst,
                        void* stream);

/* Sampler (InferenceAnimationPipeline.__call__ loop body, inference_pipeline_animation.py:644-699):
 *  prepare   : sample[b,t] = [latents[idx[t]] / sqrt(sigma^2+1) | image_latents[b0+b, idx[t]]]  (:644-650)
 *  accumulate: noise[b0+b, idx[t]] += unet_out[b,t] * weight[t]                                 (:673,685)
 *  step      : eps = (n_u + g_f (n_c - n_u)) / cnt_f ; Euler v-prediction update in fp32         (:689-699)
 * latents fp32 [F,4,h,w]; image_latents fp32 [2,F,4,h,w]; noise fp32 [2,F,4,h,w]. */
int sa_sampler_prepare(const float* latents, const float* image_latents, const int32_t* idx,
                       int32_t nb, int32_t b0, int32_t t, int32_t f, int32_t hw, float sigma, void* out,
                       int32_t act_f16, void* stream);
int sa_sampler_accumulate(const float* unet_out, int32_t ldo, const int32_t* idx, const float* weight,
                          int32_t nb, int32_t b0, int32_t t, int32_t f, int32_t hw, float* noise,
                          void* stream);
int sa_sampler_step(const float* noise, const float* cnt, const float* guidance, int32_t f, int32_t chw,
                    float sigma, float sigma_next, float* latents, void* stream);

/* Image conditioning, once per video: InferenceAnimationPipeline._encode_image
 * (inference_pipeline_animation.py:192-244).  The vision tower's Linear / LayerNorm layers run on sa_gemm_tap /
 * sa_layernorm / sa_small_linear / sa_layernorm_f32 above; these four complete it.
 *  sa_gauss_blur_axis     : one pass of _gaussian_blur2d (:83-96, _filter2d :44-64): depthwise k-tap correlation along
 *                           rows (axis 0) or columns (axis 1) of `planes` fp32 [h, w] planes, reflect padding
 *                           ((k - 1) / 2 each side, k odd); taps: k device floats; in != out.  Each input value is
 *                           read as in_scale * x + in_shift (the `image * 2.0 - 1.0` of :208 on the first pass).
 *  sa_resize_bicubic_norm : F.interpolate(mode="bicubic", align_corners=True) (:124) of fp32 [b, 3, h, w] to
 *                           [b, 3, oh, ow], then (x + 1) / 2 (:210) and the CLIPImageProcessor normalisation
 *                           (x - mean[c]) / std[c] (:213-220); mean3 / std3: three HOST floats each.
 *  sa_patchify            : im2col of the patch embedding Conv2d(3 -> C, kernel = stride = patch, no bias) of
 *                           transformers' CLIPVisionEmbeddings: pix fp32 [b, c, hw, hw] -> bf16
 *                           [b * (hw/patch)^2, ldo], column (ci * patch + ky) * patch + kx, zero beyond c*patch^2.
 *  sa_attn_tokens         : CLIPAttention core: out[i, h*d + :] = softmax_j(q_i . k_j * scale) v_j per image and head
 *                           over n tokens; qkv bf16 [b*n, ld] with q | k | v column blocks of width heads*d
 *                           (d even, <= 128); out bf16 [b*n, ldo]; fp32 math. */
int sa_gauss_blur_axis(const float* in, float* out, int32_t planes, int32_t h, int32_t w, const float* taps,
                       int32_t k, int32_t axis, float in_scale, float in_shift, void* stream);
int sa_resize_bicubic_norm(const float* in, int32_t b, int32_t c, int32_t h, int32_t w, int32_t oh, int32_t ow,
                           const float* mean3, const float* std3, float* out, void* stream);
int sa_patchify(const float* pix, int32_t b, int32_t c, int32_t hw, int32_t patch, void* out, int32_t ldo,
                void* stream);
int sa_attn_tokens(const void* qkv, int32_t ld, int32_t b, int32_t n, int32_t heads, int32_t d, float scale,
                   void* out, int32_t ldo, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SA_B200_H_ */
