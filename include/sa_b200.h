/*
 * sa_b200.h -- C ABI of the B200-native (sm_100a) kernels behind StableAnimator's denoising hot
 * path (UNetSpatioTemporalConditionModel.forward inside InferenceAnimationPipeline's Euler loop).
 *
 * Conventions
 *   - every entry point is `extern "C"`, takes plain device pointers + sizes + a cudaStream_t
 *     (passed as void*), launches asynchronously on that stream, never allocates, never syncs;
 *   - returns 0 on success, a negative SA_ERR_* code otherwise; sa_last_error() gives the text;
 *   - activations are bf16, channel-innermost ("NHWC": [frames, h, w, C] == tokens [frames*h*w, C]);
 *     weights are bf16 K-major ([out, in]); biases / embeddings / statistics are fp32.
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
 *   if geglu: v = v_value * gelu_erf(v_gate)   (GEGLU, unet.py:868-870; W rows interleaved by
 *                                               tile_n/2 blocks: [value block | gate block])
 *   v += residual[m, n]
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
  const float* row_bias; /* [groups, n_out] or NULL */
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
} sa_gemm_args;

int sa_gemm_tap(const sa_gemm_args* args, void* stream);
/* tile_n the heuristic would pick (weights for GEGLU must be interleaved with this tile) */
int sa_gemm_pick_tile(int32_t m, int32_t n, int32_t geglu);

#ifdef __cplusplus
}
#endif
#endif /* SA_B200_H_ */
