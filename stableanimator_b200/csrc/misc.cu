// Small / memory-bound kernels around the tensor-core path: skinny linears (time / context
// embeddings), sinusoidal embeddings, direct convolution for tiny channel counts (conv_in,
// PoseNet), nearest x2 upsample, stride-2 im2col, layout conversion at the drop-in boundary and
// the fused sampler kernels (input assembly, tile blend, CFG + Euler step).
#include "common.cuh"
#include "ptx.cuh"

namespace sa {

// ------------------------------------------------------------------ skinny linear (M <= 16)
// out[m, n] = act_out( sum_k act_in(x[m, k]) * W[n, k] + b[n] ).  One warp per output column.
// act_out: 0 none, 1 SiLU, 2 exact GELU (FusionFaceId proj / FeedForward, id_encoder.py:57-64,116-120).
// TimestepEmbedding (unet.py:429-436), time_emb_proj(silu(temb)) (:1669-1670, :1754-1755), the
// K/V projections of the 5 context tokens (:1974-1975, :1999-2000, :1834-1835).
constexpr int SL_NPW = 4;   // output columns per warp: four weight rows in flight per lane, the input rows read once for all
template <int MT>
__global__ void __launch_bounds__(256)
small_linear_kernel(const float* __restrict__ x, int ldx, int M, const __nv_bfloat16* __restrict__ w,
                    int K, int N, const float* __restrict__ bias, int silu_in, int silu_out,
                    int accumulate, float* __restrict__ out, int ldo, int x_smem) {
  // the MT input rows (SiLU applied once, not once per output column) staged in shared memory when they fit (x_smem)
  extern __shared__ float sl_x[];   // [MT][K]
  const int lane = threadIdx.x & 31;
  const int n0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * SL_NPW;
  const int m0 = blockIdx.y * MT;
  if (x_smem) {
    for (int i = threadIdx.x; i < MT * K; i += blockDim.x) {
      const int m = i / K, k = i - m * K;
      float v = 0.f;
      if (m0 + m < M) {
        v = x[static_cast<long long>(m0 + m) * ldx + k];
        if (silu_in) v = v / (1.0f + expf(-v));
      }
      sl_x[i] = v;
    }
    __syncthreads();
  }
  if (n0 >= N) return;
  float acc[SL_NPW][MT];
#pragma unroll
  for (int j = 0; j < SL_NPW; ++j)
#pragma unroll
    for (int m = 0; m < MT; ++m) acc[j][m] = 0.f;
  for (int k = lane * 8; k < K; k += 256) {
    uint4 u[SL_NPW];
#pragma unroll
    for (int j = 0; j < SL_NPW; ++j) {
      const int n = n0 + j < N ? n0 + j : N - 1;   // a clamped duplicate row, discarded below
      u[j] = __ldg(reinterpret_cast<const uint4*>(w + static_cast<long long>(n) * K + k));
    }
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      if (m0 + m < M) {
        float4 a, b;
        if (x_smem) {
          a = *reinterpret_cast<const float4*>(sl_x + m * K + k);
          b = *reinterpret_cast<const float4*>(sl_x + m * K + k + 4);
        } else {
          a = __ldg(reinterpret_cast<const float4*>(x + static_cast<long long>(m0 + m) * ldx + k));
          b = __ldg(reinterpret_cast<const float4*>(x + static_cast<long long>(m0 + m) * ldx + k + 4));
        }
        float xv[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        if (silu_in && !x_smem) {
#pragma unroll
          for (int t = 0; t < 8; ++t) xv[t] = xv[t] / (1.0f + expf(-xv[t]));
        }
#pragma unroll
        for (int j = 0; j < SL_NPW; ++j) {
          const uint32_t ww[4] = {u[j].x, u[j].y, u[j].z, u[j].w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 f = unpack_bf16x2(ww[t]);
            acc[j][m] = fmaf(xv[2 * t], f.x, acc[j][m]);
            acc[j][m] = fmaf(xv[2 * t + 1], f.y, acc[j][m]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < SL_NPW; ++j)
#pragma unroll
    for (int m = 0; m < MT; ++m) {
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) acc[j][m] += __shfl_xor_sync(0xffffffffu, acc[j][m], d);
    }
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < SL_NPW; ++j) {
      const int n = n0 + j;
      if (n >= N) break;
      const float b = bias ? bias[n] : 0.f;
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        if (m0 + m < M) {
          float v = acc[j][m] + b;
          if (silu_out == 1) v = v / (1.0f + expf(-v));
          else if (silu_out == 2) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));  // exact GELU
          float* dst = out + static_cast<long long>(m0 + m) * ldo + n;
          *dst = accumulate ? *dst + v : v;
        }
      }
    }
  }
}

// ------------------------------------------------------------------ FusionFaceId helpers (id_encoder.py)
// The face-token encoder runs once per video on <= 10 rows of 1024 channels: fp32 activations, one warp per
// row (LayerNorm) or per (batch, head, query) (PerceiverAttention).
// LayerNorm over C on fp32 rows (id_encoder.py:24-25,34-35,59,78,125), two-pass variance.
__global__ void layernorm_f32_kernel(const float* __restrict__ x, int ldx, int M, int C, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float eps, float* __restrict__ y, int ldy) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* xr = x + static_cast<long long>(row) * ldx;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += xr[c];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  const float mean = s / static_cast<float>(C);
  float v = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float dlt = xr[c] - mean;
    v += dlt * dlt;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  const float rstd = rsqrtf(v / static_cast<float>(C) + eps);
  float* yr = y + static_cast<long long>(row) * ldy;
  for (int c = lane; c < C; c += 32) yr[c] = (xr[c] - mean) * rstd * gamma[c] + beta[c];
}

// PerceiverAttention core (id_encoder.py:48-55): out[b, l, h*64 + d] = softmax_n((q k^T) / 8) v, head_dim 64,
// n <= 32 keys.  q fp32 [B, L, ldq]; k and v fp32 rows [B, N, ldkv] (k at column h*64, v at voff + h*64).
// The reference scales q and k by 64^-1/4 each and takes the softmax in fp32 -- the same product.
__global__ void perceiver_attn_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ kv, int ldkv,
                                      int voff, int B, int L, int N, int heads, float* __restrict__ out, int ldo) {
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= B * heads * L) return;
  const int l = w % L, h = (w / L) % heads, b = w / (L * heads);
  const float* qr = q + (static_cast<long long>(b) * L + l) * ldq + h * 64;
  const float q0 = qr[lane], q1 = qr[lane + 32];
  float sc[32];
  float mx = -INFINITY;
  for (int n = 0; n < N; ++n) {
    const float* kr = kv + (static_cast<long long>(b) * N + n) * ldkv + h * 64;
    float d = q0 * kr[lane] + q1 * kr[lane + 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    sc[n] = d * 0.125f;
    mx = fmaxf(mx, sc[n]);
  }
  float den = 0.f, o0 = 0.f, o1 = 0.f;
  for (int n = 0; n < N; ++n) {
    const float pn = expf(sc[n] - mx);
    const float* vr = kv + (static_cast<long long>(b) * N + n) * ldkv + voff + h * 64;
    den += pn;
    o0 += pn * vr[lane];
    o1 += pn * vr[lane + 32];
  }
  float* orow = out + (static_cast<long long>(b) * L + l) * ldo + h * 64;
  orow[lane] = o0 / den;
  orow[lane + 32] = o1 / den;
}

// ------------------------------------------------------------------ temporal VAE decoder helpers (vae.py)
// Row softmax of the single-head, head_dim = C mid-block attention (vae.py:92-94): scores fp32 [M, N] from
// the Q K^T GEMM -> probabilities bf16 [M, N] for the P V GEMM.  One warp per row, three L2-resident passes.
__global__ void softmax_rows_kernel(const float* __restrict__ s, long long lds, int M, int N, float scale,
                                    __nv_bfloat16* __restrict__ p, long long ldp) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* sr = s + row * lds;
  float mx = -INFINITY;
  for (int c = lane; c < N; c += 32) mx = fmaxf(mx, sr[c]);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
  const float sl2 = scale * 1.4426950408889634f;
  float sum = 0.f;
  for (int c = lane; c < N; c += 32) sum += exp2f((sr[c] - mx) * sl2);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
  const float inv = 1.0f / sum;
  __nv_bfloat16* pr = p + row * ldp;
  for (int c = lane; c < N; c += 32) pr[c] = __float2bfloat16_rn(exp2f((sr[c] - mx) * sl2) * inv);
}

// time_conv_out (vae.py:423-429,470-475): Conv3d(C -> C, kernel (3,1,1), zero padding over frames) on the
// conv_out result.  x fp32 [B*F*HW, ldx] (first C columns, channel-innermost) -> out fp32 NCHW [B*F, C, HW].
__global__ void time_conv_out_kernel(const float* __restrict__ x, int ldx, int B, int F, int HW, int C,
                                     const float* __restrict__ w, const float* __restrict__ bias,
                                     float* __restrict__ out) {
  const long long total = static_cast<long long>(B) * F * HW;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int pix = static_cast<int>(idx % HW);
    const long long bf = idx / HW;
    const int f = static_cast<int>(bf % F);
    for (int co = 0; co < C; ++co) {
      float acc = bias[co];
      for (int kt = 0; kt < 3; ++kt) {
        const int fi = f + kt - 1;
        if (fi < 0 || fi >= F) continue;
        const float* xr = x + ((bf - f + fi) * HW + pix) * ldx;
        for (int ci = 0; ci < C; ++ci) acc += w[(co * C + ci) * 3 + kt] * xr[ci];
      }
      out[(bf * C + co) * HW + pix] = acc;
    }
  }
}

// ------------------------------------------------------------------ sinusoidal embedding
// get_timestep_embedding (unet.py:365-390): out[i] = [cos(t_i f_k) | sin(t_i f_k)], f_k = 1e4^(-k/half)
__global__ void sinusoid_kernel(const float* __restrict__ t, int n, int dim, float* __restrict__ out) {
  const int half = dim >> 1;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * half) return;
  const int i = idx / half, k = idx % half;
  const float freq = expf(-9.210340371976184f * static_cast<float>(k) / static_cast<float>(half));
  const float arg = t[i] * freq;
  out[static_cast<long long>(i) * dim + k] = cosf(arg);
  out[static_cast<long long>(i) * dim + half + k] = sinf(arg);
}

// ------------------------------------------------------------------ direct convolution
// NHWC bf16 in, [k*k*Cin, CoutPad] bf16 weights, one thread = one output pixel x 8 out channels.
// conv_in (unet.py:69,283 + pose add :284-286) and the PoseNet stack (pose_net.py:14-37).
__global__ void __launch_bounds__(256)
conv_direct_kernel(const __nv_bfloat16* __restrict__ x, int N, int H, int W, int Cin,
                   const __nv_bfloat16* __restrict__ w, const float* __restrict__ bias, int Cout,
                   int CoutPad, int ksz, int stride, int pad, int Ho, int Wo, int silu_out,
                   float scale, const __nv_bfloat16* __restrict__ add, int add_from,
                   __nv_bfloat16* __restrict__ out, int f16_) {
  const bool f16 = f16_ != 0;   // x / w / add / out hold IEEE fp16 instead of bf16
  const int octs = CoutPad >> 3;
  const long long total = static_cast<long long>(N) * Ho * Wo * octs;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int oc = static_cast<int>(idx % octs) * 8;
    long long pix = idx / octs;
    const int xo = static_cast<int>(pix % Wo);
    pix /= Wo;
    const int yo = static_cast<int>(pix % Ho);
    const int n = static_cast<int>(pix / Ho);
    float acc[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t] = 0.f;
    for (int ky = 0; ky < ksz; ++ky) {
      const int yi = yo * stride + ky - pad;
      if (yi < 0 || yi >= H) continue;
      for (int kx = 0; kx < ksz; ++kx) {
        const int xi = xo * stride + kx - pad;
        if (xi < 0 || xi >= W) continue;
        const __nv_bfloat16* xp = x + ((static_cast<long long>(n) * H + yi) * W + xi) * Cin;
        const __nv_bfloat16* wp = w + static_cast<long long>((ky * ksz + kx) * Cin) * CoutPad + oc;
        for (int ci = 0; ci < Cin; ++ci) {
          const float xv = load16(xp + ci, f16);
          const uint4 u = __ldg(reinterpret_cast<const uint4*>(wp + static_cast<long long>(ci) * CoutPad));
          const uint32_t ww[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 f = unpack16x2(ww[t], f16);
            acc[2 * t] += xv * f.x;
            acc[2 * t + 1] += xv * f.y;
          }
        }
      }
    }
    const long long orow = (static_cast<long long>(n) * Ho + yo) * Wo + xo;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      if (oc + t < Cout) {
        float v = acc[t] + (bias ? bias[oc + t] : 0.f);
        if (silu_out) v = silu(v);
        v *= scale;
        if (add && n >= add_from)
          v += load16(add + ((static_cast<long long>(n - add_from) * Ho + yo) * Wo + xo) * Cout + oc + t, f16);
        if (f16) reinterpret_cast<__half*>(out)[orow * Cout + oc + t] = __float2half_rn(fminf(fmaxf(v, -65504.0f), 65504.0f));
        else out[orow * Cout + oc + t] = __float2bfloat16_rn(v);
      }
    }
  }
}

// ------------------------------------------------------------------ nearest x2 upsample (NHWC)
// F.interpolate(scale_factor=2, mode="nearest"), unet.py:1799: out[y, x] = in[y / 2, x / 2]
__global__ void upsample2x_kernel(const uint4* __restrict__ in, int N, int H, int W, int C8,
                                  uint4* __restrict__ out) {
  const long long total = static_cast<long long>(N) * 4 * H * W * C8;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % C8);
    long long p = idx / C8;
    const int xo = static_cast<int>(p % (2 * W));
    p /= 2 * W;
    const int yo = static_cast<int>(p % (2 * H));
    const long long n = p / (2 * H);
    out[idx] = __ldg(in + ((n * H + (yo >> 1)) * W + (xo >> 1)) * C8 + c);
  }
}

// ------------------------------------------------------------------ im2col for 3x3 stride 2 pad 1
// Downsample2D conv (unet.py:1474): out[(n, yo, xo), tap*C + c] = in[n, 2yo+ky-1, 2xo+kx-1, c] | 0
// pad = 1: symmetric zero padding (UNet Downsample2D); pad = 0: F.pad(x, (0, 1, 0, 1)) + padding-0 conv of the
// VAE encoder's Downsample2D (vae.py:731-741) -- taps reach one pixel past the right / bottom edge only.
__global__ void im2col_s2_kernel(const uint4* __restrict__ in, int N, int H, int W, int C8, int pad,
                                 uint4* __restrict__ out) {
  const int Ho = H >> 1, Wo = W >> 1;
  const long long total = static_cast<long long>(N) * Ho * Wo * 9 * C8;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % C8);
    long long p = idx / C8;
    const int tap = static_cast<int>(p % 9);
    p /= 9;
    const int xo = static_cast<int>(p % Wo);
    p /= Wo;
    const int yo = static_cast<int>(p % Ho);
    const long long n = p / Ho;
    const int yi = 2 * yo + tap / 3 - pad, xi = 2 * xo + tap % 3 - pad;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (yi >= 0 && yi < H && xi >= 0 && xi < W) v = __ldg(in + ((n * H + yi) * W + xi) * C8 + c);
    out[idx] = v;
  }
}

// ------------------------------------------------------------------ im2col for conv_in (3x3, pad 1, Cin = 8)
// conv_in (unet.py:69,283) has K = 9*8 = 72: too thin for an implicit-GEMM tap (64-channel TMA boxes), so
// its rows are gathered once into [M, 128] (72 values + zero padding to two 64-wide K blocks) and the
// tensor-core GEMM does the rest (with the pose residual in its epilogue).  One uint4 = one pixel's 8 channels.
__global__ void im2col3x3_c8_kernel(const uint4* __restrict__ in, int N, int H, int W, uint4* __restrict__ out) {
  const long long total = static_cast<long long>(N) * H * W * 16;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int tap = static_cast<int>(idx & 15);
    long long p = idx >> 4;
    const int x = static_cast<int>(p % W);
    p /= W;
    const int y = static_cast<int>(p % H);
    const long long n = p / H;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (tap < 9) {
      const int yi = y + tap / 3 - 1, xi = x + tap % 3 - 1;
      if (yi >= 0 && yi < H && xi >= 0 && xi < W) v = __ldg(in + (n * H + yi) * W + xi);
    }
    out[idx] = v;
  }
}

// ------------------------------------------------------------------ layout conversion (boundary)
// src [N, C, HW] (fp32, or 16-bit in the activation format) -> dst 16-bit [N, HW, C]
template <typename T>
__global__ void nchw_to_nhwc_kernel(const T* __restrict__ src, int C, int HW,
                                    __nv_bfloat16* __restrict__ dst, int f16) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, p = p0 + threadIdx.x;
    if (c < C && p < HW) {
      const T* sp = src + (static_cast<long long>(n) * C + c) * HW + p;
      if constexpr (sizeof(T) == 2) tile[j][threadIdx.x] = load16(sp, f16 != 0);
      else tile[j][threadIdx.x] = static_cast<float>(*sp);
    }
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int p = p0 + j, c = c0 + threadIdx.x;
    if (c < C && p < HW) {
      const float v = tile[threadIdx.x][j];
      const long long o = (static_cast<long long>(n) * HW + p) * C + c;
      if (f16) reinterpret_cast<__half*>(dst)[o] = __float2half_rn(fminf(fmaxf(v, -65504.0f), 65504.0f));
      else dst[o] = __float2bfloat16_rn(v);
    }
  }
}
// src fp32 [N, HW, lds] (first C columns) -> dst [N, C, HW] fp32, dst = dst*keep + scale*src
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ src, int lds, int C, int HW,
                                    float* __restrict__ dst) {
  const long long total = static_cast<long long>(gridDim.y) * C * HW;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (idx >= static_cast<long long>(C) * HW) return;
  (void)total;
  const int c = static_cast<int>(idx / HW), p = static_cast<int>(idx % HW);
  dst[(static_cast<long long>(n) * C + c) * HW + p] = src[(static_cast<long long>(n) * HW + p) * lds + c];
}

// ------------------------------------------------------------------ sampler kernels
// inference_pipeline_animation.py:644-650 for one frame tile: sample[b, t, y, x, 0:4] =
// latents[idx[t]] / sqrt(sigma^2+1), [4:8] = image_latents[b, idx[t]]  (NHWC bf16, C = 8)
__global__ void sampler_prepare_kernel(const float* __restrict__ latents,
                                       const float* __restrict__ image_latents,
                                       const int* __restrict__ idx, int nb, int b0, int T, int F, int HW,
                                       float inv_scale, __nv_bfloat16* __restrict__ out, int f16_) {
  const bool f16 = f16_ != 0;
  const long long total = static_cast<long long>(nb) * T * HW;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int p = static_cast<int>(i % HW);
  const int t = static_cast<int>((i / HW) % T);
  const int b = static_cast<int>(i / (static_cast<long long>(HW) * T));
  const int f = idx[t];
  float v[8];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    v[c] = latents[(static_cast<long long>(f) * 4 + c) * HW + p] * inv_scale;
    v[4 + c] = image_latents[((static_cast<long long>(b0 + b) * F + f) * 4 + c) * HW + p];
  }
  *reinterpret_cast<uint4*>(out + i * 8) = make_uint4(pack16x2(v[0], v[1], f16), pack16x2(v[2], v[3], f16),
                                                      pack16x2(v[4], v[5], f16), pack16x2(v[6], v[7], f16));
}

// :673,685: noise[b0 + b, idx[t]] += unet_out[b, t] * weight[t]   (unet_out: fp32 NHWC, ld = ldo)
__global__ void sampler_accumulate_kernel(const float* __restrict__ unet_out, int ldo,
                                          const int* __restrict__ idx, const float* __restrict__ weight,
                                          int nb, int b0, int T, int F, int HW, float* __restrict__ noise) {
  const long long total = static_cast<long long>(nb) * T * HW;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int p = static_cast<int>(i % HW);
  const int t = static_cast<int>((i / HW) % T);
  const int b = static_cast<int>(i / (static_cast<long long>(HW) * T));
  const int f = idx[t];
  const float w = weight[t];
  const float4 o = *reinterpret_cast<const float4*>(unet_out + i * ldo);
  float* dst = noise + ((static_cast<long long>(b0 + b) * F + f) * 4) * HW + p;
  dst[0] += o.x * w;
  dst[static_cast<long long>(HW)] += o.y * w;
  dst[2LL * HW] += o.z * w;
  dst[3LL * HW] += o.w * w;
}

// :689-699: eps = (n_u + g_f (n_c - n_u)) / cnt_f ; Euler v-prediction step in fp32:
//   x0 = eps * (-sigma / sqrt(sigma^2+1)) + x / (sigma^2+1);  x += (x - x0) / sigma * (sigma_next - sigma)
__global__ void sampler_step_kernel(const float* __restrict__ noise, const float* __restrict__ cnt,
                                    const float* __restrict__ guidance, int F, int CHW, float sigma,
                                    float sigma_next, float* __restrict__ latents) {
  const long long total = static_cast<long long>(F) * CHW;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int f = static_cast<int>(i / CHW);
  const float inv = 1.0f / cnt[f];
  const float nu = noise[i] * inv, nc = noise[total + i] * inv;
  const float eps = nu + guidance[f] * (nc - nu);
  const float s2 = sigma * sigma + 1.0f;
  const float x = latents[i];
  const float x0 = eps * (-sigma / sqrtf(s2)) + x / s2;
  const float d = (x - x0) / sigma;
  latents[i] = x + d * (sigma_next - sigma);
}

}  // namespace sa

using namespace sa;

static inline unsigned grid_for(long long total, int block, int max_blocks = 148 * 16) {
  long long g = (total + block - 1) / block;
  if (g > max_blocks) g = max_blocks;
  if (g < 1) g = 1;
  return static_cast<unsigned>(g);
}

extern "C" int sa_small_linear(const float* x, int32_t ldx, int32_t m, const void* w, int32_t k,
                               int32_t n, const float* bias, int32_t silu_in, int32_t silu_out,
                               int32_t accumulate, float* out, int32_t ldo, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !w || !out) return fail(SA_ERR_INVALID, "sa_small_linear: null pointer");
  if (k % 8 || ldx % 4 || m < 1) return fail(SA_ERR_INVALID, "sa_small_linear: k %% 8, ldx %% 4 required");
  const __nv_bfloat16* wp = static_cast<const __nv_bfloat16*>(w);
  const unsigned gx = (n + 8 * SL_NPW - 1) / (8 * SL_NPW);
  // input rows staged in shared memory when MT * K floats fit the default 48 KB
  const int mt = m <= 2 ? 2 : m <= 4 ? 4 : 8;
  const size_t xs = static_cast<size_t>(mt) * k * sizeof(float);
  const int x_smem = xs <= 48 * 1024 ? 1 : 0;
  const size_t smem = x_smem ? xs : 0;
  if (m <= 2) {
    small_linear_kernel<2><<<dim3(gx, (m + 1) / 2), 256, smem, stream>>>(x, ldx, m, wp, k, n, bias, silu_in, silu_out, accumulate, out, ldo, x_smem);
  } else if (m <= 4) {
    small_linear_kernel<4><<<dim3(gx, (m + 3) / 4), 256, smem, stream>>>(x, ldx, m, wp, k, n, bias, silu_in, silu_out, accumulate, out, ldo, x_smem);
  } else {
    small_linear_kernel<8><<<dim3(gx, (m + 7) / 8), 256, smem, stream>>>(x, ldx, m, wp, k, n, bias, silu_in, silu_out, accumulate, out, ldo, x_smem);
  }
  SA_CUDA_CHECK(cudaGetLastError());
  return SA_OK;
}

extern "C" int sa_layernorm_f32(const float* x, int32_t ldx, int32_t m, int32_t c, const float* gamma,
                                const float* beta, float eps, float* y, int32_t ldy, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !gamma || !beta || !y) return fail(SA_ERR_INVALID, "sa_layernorm_f32: null pointer");
  if (m < 1 || c < 1 || ldx < c || ldy < c) return fail(SA_ERR_INVALID, "sa_layernorm_f32: bad sizes");
  layernorm_f32_kernel<<<(m + 3) / 4, 128, 0, stream>>>(x, ldx, m, c, gamma, beta, eps, y, ldy);
  SA_CUDA_CHECK(cudaGetLastError());
  return SA_OK;
}

extern "C" int sa_perceiver_attn(const float* q, int32_t ldq, const float* kv, int32_t ldkv, int32_t voff, int32_t b,
                                 int32_t l, int32_t n, int32_t heads, float* out, int32_t ldo, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!q || !kv || !out) return fail(SA_ERR_INVALID, "sa_perceiver_attn: null pointer");
  if (b < 1 || l < 1 || n < 1 || n > 32 || heads < 1) return fail(SA_ERR_INVALID, "sa_perceiver_attn: n must be in [1, 32]");
  const int warps = b * heads * l;
  perceiver_attn_kernel<<<(warps + 3) / 4, 128, 0, stream>>>(q, ldq, kv, ldkv, voff, b, l, n, heads, out, ldo);
  SA_CUDA_CHECK(cudaGetLastError());
  return SA_OK;
}

extern "C" int sa_softmax_rows(const float* s, int64_t lds, int32_t m, int32_t n, float scale, void* p, int64_t ldp,
                               void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!s || !p) return fail(SA_ERR_INVALID, "sa_softmax_rows: null pointer");
  if (m < 1 || n < 1 || lds < n || ldp < n) return fail(SA_ERR_INVALID, "sa_softmax_rows: bad sizes");
  softmax_rows_kernel<<<(m + 7) / 8, 256, 0, stream>>>(s, lds, m, n, scale, static_cast<__nv_bfloat16*>(p), ldp);
  SA_CUDA_CHECK(cudaGetLastError());
  return SA_OK;
}

extern "C" int sa_time_conv_out(const float* x, int32_t ldx, int32_t b, int32_t f, int32_t hw, int32_t c, const float* w,
                                const float* bias, float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !w || !bias || !out) return fail(SA_ERR_INVALID, "sa_time_conv_out: null pointer");
  if (b < 1 || f < 1 || hw < 1 || c < 1 || c > 8 || ldx < c) return fail(SA_ERR_INVALID, "sa_time_conv_out: bad sizes");
  const long long total = static_cast<long long>(b) * f * hw;
  time_conv_out_kernel<<<grid_for(total, 256), 256, 0, stream>>>(x, ldx, b, f, hw, c, w, bias, out);
  SA_CUDA_CHECK(cudaGetLastError());
  return SA_OK;
}

extern "C" int sa_sinusoid(const float* t, int32_t n, int32_t dim, float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!t || !out || dim % 2) return fail(SA_ERR_INVALID, "sa_sinusoid: bad arguments");
  const int total = n * (dim / 2);
  sinusoid_kernel<<<(total + 255) / 256, 256, 0, stream>>>(t, n, dim, out);
  SA_CUDA_CHECK(cudaGetLastError());
  return SA_OK;
}

extern "C" int sa_conv_direct(const void* x, int32_t n, int32_t h, int32_t w, int32_t cin,
                              const void* wt, const float* bias, int32_t cout, int32_t cout_pad,
                              int32_t ksz, int32_t stride, int32_t pad, int32_t silu_out, float scale,
                              const void* add, int32_t add_from, void* out, int32_t act_f16, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !wt || !out) return fail(SA_ERR_INVALID, "sa_conv_direct: null pointer");
  if (cout_pad % 8 || cout_pad < cout) return fail(SA_ERR_INVALID, "sa_conv_direct: cout_pad must be a multiple of 8");
  const int ho = (h + 2 * pad - ksz) / stride + 1, wo = (w + 2 * pad - ksz) / stride + 1;
  const long long total = static_cast<long long>(n) * ho * wo * (cout_pad / 8);
  conv_direct_kernel<<<grid_for(total, 256, 148 * 32), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(x), n, h, w, cin, static_cast<const __nv_bfloat16*>(wt), bias, cout,
      cout_pad, ksz, stride, pad, ho, wo, silu_out, scale, static_cast<const __nv_bfloat16*>(add), add_from,
      static_cast<__nv_bfloat16*>(out), act_f16);
  SA_CUDA_CHECK(cudaGetLastError());
  return SA_OK;
}

extern "C" int sa_upsample2x(const void* in, int32_t n, int32_t h, int32_t w, int32_t c, void* out,
                             void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!in || !out || c % 8) return fail(SA_ERR_INVALID, "sa_upsample2x: bad arguments");
  const long long total = static_cast<long long>(n) * 4 * h * w * (c / 8);
  upsample2x_kernel<<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const uint4*>(in), n, h, w, c / 8,
                                                             static_cast<uint4*>(out));
  SA_CUDA_CHECK(cudaGetLastError());
  return SA_OK;
}

extern "C" int sa_im2col_s2(const void* in, int32_t n, int32_t h, int32_t w, int32_t c, void* out,
                            void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!in || !out || c % 8 || h % 2 || w % 2) return fail(SA_ERR_INVALID, "sa_im2col_s2: bad arguments");
  const long long total = static_cast<long long>(n) * (h / 2) * (w / 2) * 9 * (c / 8);
  im2col_s2_kernel<<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const uint4*>(in), n, h, w, c / 8, 1,
                                                            static_cast<uint4*>(out));
  SA_CUDA_CHECK(cudaGetLastError());
  return SA_OK;
}

extern "C" int sa_im2col_s2_pad(const void* in, int32_t n, int32_t h, int32_t w, int32_t c, int32_t pad, void* out,
                                void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!in || !out || c % 8 || h % 2 || w % 2 || pad < 0 || pad > 1)
    return fail(SA_ERR_INVALID, "sa_im2col_s2_pad: bad arguments");
  const long long total = static_cast<long long>(n) * (h / 2) * (w / 2) * 9 * (c / 8);
  im2col_s2_kernel<<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const uint4*>(in), n, h, w, c / 8, pad,
                                                            static_cast<uint4*>(out));
  SA_CUDA_CHECK(cudaGetLastError());
  return SA_OK;
}

extern "C" int sa_im2col3x3_c8(const void* in, int32_t n, int32_t h, int32_t w, void* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!in || !out || n <= 0 || h <= 0 || w <= 0) return fail(SA_ERR_INVALID, "sa_im2col3x3_c8: bad arguments");
  const long long total = static_cast<long long>(n) * h * w * 16;
  im2col3x3_c8_kernel<<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const uint4*>(in), n, h, w,
                                                               static_cast<uint4*>(out));
  SA_CUDA_CHECK(cudaGetLastError());
  return SA_OK;
}

extern "C" int sa_nchw_to_nhwc(const void* src, int32_t src_is_bf16, int32_t n, int32_t c, int32_t hw,
                               void* dst, int32_t act_f16, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!src || !dst) return fail(SA_ERR_INVALID, "sa_nchw_to_nhwc: null pointer");
  dim3 grid((hw + 31) / 32, (c + 31) / 32, n), block(32, 8);
  if (src_is_bf16)
    nchw_to_nhwc_kernel<__nv_bfloat16><<<grid, block, 0, stream>>>(static_cast<const __nv_bfloat16*>(src), c, hw,
                                                                   static_cast<__nv_bfloat16*>(dst), act_f16);
  else
    nchw_to_nhwc_kernel<float><<<grid, block, 0, stream>>>(static_cast<const float*>(src), c, hw,
                                                           static_cast<__nv_bfloat16*>(dst), act_f16);
  SA_CUDA_CHECK(cudaGetLastError());
  return SA_OK;
}

extern "C" int sa_nhwc_to_nchw_f32(const float* src, int32_t lds, int32_t n, int32_t c, int32_t hw,
                                   float* dst, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!src || !dst) return fail(SA_ERR_INVALID, "sa_nhwc_to_nchw_f32: null pointer");
  dim3 grid((static_cast<long long>(c) * hw + 255) / 256, n);
  nhwc_to_nchw_kernel<<<grid, 256, 0, stream>>>(src, lds, c, hw, dst);
  SA_CUDA_CHECK(cudaGetLastError());
  return SA_OK;
}

extern "C" int sa_sampler_prepare(const float* latents, const float* image_latents, const int32_t* idx,
                                  int32_t nb, int32_t b0, int32_t t, int32_t f, int32_t hw, float sigma,
                                  void* out, int32_t act_f16, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!latents || !image_latents || !idx || !out) return fail(SA_ERR_INVALID, "sa_sampler_prepare: null pointer");
  const long long total = static_cast<long long>(nb) * t * hw;
  const float inv = 1.0f / sqrtf(sigma * sigma + 1.0f);
  sampler_prepare_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      latents, image_latents, idx, nb, b0, t, f, hw, inv, static_cast<__nv_bfloat16*>(out), act_f16);
  SA_CUDA_CHECK(cudaGetLastError());
  return SA_OK;
}

extern "C" int sa_sampler_accumulate(const float* unet_out, int32_t ldo, const int32_t* idx,
                                     const float* weight, int32_t nb, int32_t b0, int32_t t, int32_t f,
                                     int32_t hw, float* noise, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!unet_out || !idx || !weight || !noise || ldo % 4) return fail(SA_ERR_INVALID, "sa_sampler_accumulate: bad arguments");
  const long long total = static_cast<long long>(nb) * t * hw;
  sampler_accumulate_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      unet_out, ldo, idx, weight, nb, b0, t, f, hw, noise);
  SA_CUDA_CHECK(cudaGetLastError());
  return SA_OK;
}

extern "C" int sa_sampler_step(const float* noise, const float* cnt, const float* guidance, int32_t f,
                               int32_t chw, float sigma, float sigma_next, float* latents, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!noise || !cnt || !guidance || !latents) return fail(SA_ERR_INVALID, "sa_sampler_step: null pointer");
  const long long total = static_cast<long long>(f) * chw;
  sampler_step_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(noise, cnt, guidance, f, chw,
                                                                                    sigma, sigma_next, latents);
  SA_CUDA_CHECK(cudaGetLastError());
  return SA_OK;
}
