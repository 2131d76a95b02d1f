// CUDA-core attention kernels for the two tiny-key-count attentions of the path.  Both are
// memory/latency bound (a handful of keys per query), so they use 128-bit loads, registers and
// warp shuffles instead of tensor cores.
//
//  * temporal self-attention over the <=16 frames of a tile, per (pixel, head):
//      XFormersAttnProcessor self-attn case, animation/modules/unet.py:1830-1848, called from
//      TemporalBasicTransformerBlock.forward :940-942.  Works directly on the [B*F, S, 3C] fused
//      q|k|v tensor (frame stride S*3C), so the reference's [B*F,S,C] <-> [B*S,F,C] permutes
//      (:931-933, :954-956) and head_to_batch_dim copies (:1837-1839) never materialise.
//  * ID cross-attention against the 4 face tokens + the Sum / Sum^2 needed by the distribution
//      alignment: AnimationIDAttnNormalizedProcessor.__call__, unet.py:1999-2019.
#include "common.cuh"
#include "ptx.cuh"

namespace sa {

__device__ __forceinline__ void ld32(const __nv_bfloat16* p, float (&v)[32]) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint4 u = __ldg(q + i);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 f = unpack_bf16x2(w[t]);
      v[8 * i + 2 * t] = f.x;
      v[8 * i + 2 * t + 1] = f.y;
    }
  }
}

// One warp per (batch b, pixel s, head h).  lane = (half << 4) | frame: each lane owns query frame
// i and 32 of the 64 head dims.
__global__ void __launch_bounds__(256)
attn_temporal_kernel(const __nv_bfloat16* __restrict__ qkv, int ld, int B, int F, int S, int C,
                     int heads, __nv_bfloat16* __restrict__ out, int ldo) {
  const int lane = threadIdx.x & 31;
  const long long task = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long ntasks = static_cast<long long>(B) * S * heads;
  if (task >= ntasks) return;
  const int h = static_cast<int>(task % heads);
  const long long bs = task / heads;
  const int s = static_cast<int>(bs % S);
  const int b = static_cast<int>(bs / S);
  const int i = lane & 15, half = lane >> 4;
  const int iq = i < F ? i : F - 1;
  const long long row0 = static_cast<long long>(b) * F * S + s;  // row of frame 0
  const int col = h * 64 + half * 32;

  float q[32];
  ld32(qkv + (row0 + static_cast<long long>(iq) * S) * ld + col, q);
#pragma unroll
  for (int d = 0; d < 32; ++d) q[d] *= 0.125f;  // attn.scale = 64^-0.5 (unet.py:737)

  float sc[16];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    sc[j] = -INFINITY;
    if (j < F) {
      float k[32];
      ld32(qkv + (row0 + static_cast<long long>(j) * S) * ld + C + col, k);
      float p = 0.f;
#pragma unroll
      for (int d = 0; d < 32; ++d) p += q[d] * k[d];
      p += __shfl_xor_sync(0xffffffffu, p, 16);
      sc[j] = p;
      mx = fmaxf(mx, p);
    }
  }
  float l = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    if (j < F) {
      sc[j] = __expf(sc[j] - mx);
      l += sc[j];
    }
  }
  const float inv = 1.0f / l;
  float o[32];
#pragma unroll
  for (int d = 0; d < 32; ++d) o[d] = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    if (j < F) {
      float v[32];
      ld32(qkv + (row0 + static_cast<long long>(j) * S) * ld + 2 * C + col, v);
      const float p = sc[j] * inv;
#pragma unroll
      for (int d = 0; d < 32; ++d) o[d] += p * v[d];
    }
  }
  if (i < F) {
    uint4* dst = reinterpret_cast<uint4*>(out + (row0 + static_cast<long long>(i) * S) * ldo + col);
#pragma unroll
    for (int t = 0; t < 4; ++t)
      dst[t] = make_uint4(pack_bf16x2(o[8 * t], o[8 * t + 1]), pack_bf16x2(o[8 * t + 2], o[8 * t + 3]),
                          pack_bf16x2(o[8 * t + 4], o[8 * t + 5]),
                          pack_bf16x2(o[8 * t + 6], o[8 * t + 7]));
  }
}

// ID cross-attention: grid (ceil(S/32), B*F); 256 threads; task = (row in chunk, head), lanes of a
// warp share the head so K/V smem reads broadcast.  ik/iv: fp32 [B, NTOK, C].
template <int NTOK>
__global__ void __launch_bounds__(256)
attn_id_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const float* __restrict__ ik,
               const float* __restrict__ iv, long long tok_stride, long long batch_stride, int F, int S,
               int C, int heads,
               __nv_bfloat16* __restrict__ out, int ldo, double* __restrict__ stats) {
  extern __shared__ float sm_kv[];  // [2][NTOK][C]
  __shared__ double red[2][8];
  const int frame = blockIdx.y;
  const int b = frame / F;
  float* sk = sm_kv;
  float* sv = sm_kv + NTOK * C;
  for (int t = threadIdx.x; t < NTOK * C; t += blockDim.x) {
    const long long off = static_cast<long long>(b) * batch_stride + (t / C) * tok_stride + (t % C);
    sk[t] = ik[off];
    sv[t] = iv[off];
  }
  __syncthreads();
  const int s0 = blockIdx.x * 32;
  double lsum = 0.0, lsq = 0.0;
  for (int t = threadIdx.x; t < 32 * heads; t += blockDim.x) {
    const int r = t & 31, h = t >> 5;
    const int s = s0 + r;
    if (s >= S) continue;
    const long long row = static_cast<long long>(frame) * S + s;
    float qv[64];
    {
      const uint4* qp = reinterpret_cast<const uint4*>(q + row * ldq + h * 64);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint4 u = __ldg(qp + i);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = unpack_bf16x2(w[k]);
          qv[8 * i + 2 * k] = f.x;
          qv[8 * i + 2 * k + 1] = f.y;
        }
      }
    }
    float sc[NTOK];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < NTOK; ++j) {
      const float4* kp = reinterpret_cast<const float4*>(sk + j * C + h * 64);
      float p = 0.f;
#pragma unroll
      for (int d = 0; d < 16; ++d) {
        const float4 kk = kp[d];
        p += qv[4 * d] * kk.x + qv[4 * d + 1] * kk.y + qv[4 * d + 2] * kk.z + qv[4 * d + 3] * kk.w;
      }
      sc[j] = p * 0.125f;
      mx = fmaxf(mx, sc[j]);
    }
    float l = 0.f;
#pragma unroll
    for (int j = 0; j < NTOK; ++j) {
      sc[j] = __expf(sc[j] - mx);
      l += sc[j];
    }
    const float inv = 1.0f / l;
    float o[64];
#pragma unroll
    for (int d = 0; d < 64; ++d) o[d] = 0.f;
#pragma unroll
    for (int j = 0; j < NTOK; ++j) {
      const float4* vp = reinterpret_cast<const float4*>(sv + j * C + h * 64);
      const float p = sc[j] * inv;
#pragma unroll
      for (int d = 0; d < 16; ++d) {
        const float4 vv = vp[d];
        o[4 * d] += p * vv.x;
        o[4 * d + 1] += p * vv.y;
        o[4 * d + 2] += p * vv.z;
        o[4 * d + 3] += p * vv.w;
      }
    }
    float ts = 0.f, tq = 0.f;
#pragma unroll
    for (int d = 0; d < 64; ++d) {
      ts += o[d];
      tq += o[d] * o[d];
    }
    lsum += ts;
    lsq += tq;
    uint4* dst = reinterpret_cast<uint4*>(out + row * ldo + h * 64);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      dst[i] = make_uint4(pack_bf16x2(o[8 * i], o[8 * i + 1]), pack_bf16x2(o[8 * i + 2], o[8 * i + 3]),
                          pack_bf16x2(o[8 * i + 4], o[8 * i + 5]),
                          pack_bf16x2(o[8 * i + 6], o[8 * i + 7]));
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    lsum += __shfl_xor_sync(0xffffffffu, lsum, d);
    lsq += __shfl_xor_sync(0xffffffffu, lsq, d);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[0][warp] = lsum;
    red[1][warp] = lsq;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, c2 = 0.0;
    for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) {
      a += red[0][w];
      c2 += red[1][w];
    }
    atomicAdd(&stats[2 * frame], a);
    atomicAdd(&stats[2 * frame + 1], c2);
  }
}

// Per frame: a = std_lat / (std_ip + 1e-5);  row_bias[f, n] = W.v_clip[b, n] + bias[n]
//   + (mean_lat - a * mean_ip) * rowsum(W)[n]          (unet.py:2018-2024 folded through to_out)
// where the CLIP branch of the attention is the constant vector v_clip[b] (softmax over 1 key).
__global__ void __launch_bounds__(256)
attn_id_finalize_kernel(const double* __restrict__ stats, const float* __restrict__ vclip,
                        long long vclip_bstride, const float* __restrict__ wv, const float* __restrict__ w_rowsum,
                        const float* __restrict__ bias, int F, int S, int C,
                        float* __restrict__ row_scale, float* __restrict__ row_bias) {
  __shared__ double red[2][8];
  __shared__ float s_a, s_shift;
  const int frame = blockIdx.x;
  const int b = frame / F;
  const float* vc = vclip + static_cast<long long>(b) * vclip_bstride;
  double s1 = 0.0, s2 = 0.0;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double v = vc[c];
    s1 += v;
    s2 += v * v;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, d);
    s2 += __shfl_xor_sync(0xffffffffu, s2, d);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = s1;
    red[1][threadIdx.x >> 5] = s2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a1 = 0.0, a2 = 0.0;
    for (int w = 0; w < 8; ++w) {
      a1 += red[0][w];
      a2 += red[1][w];
    }
    const double n = static_cast<double>(S) * C;
    // latents branch: every pixel row equals v_clip -> sums are S * (per-channel sums)
    const double mean_l = a1 / C;
    double var_l = (static_cast<double>(S) * a2 - n * mean_l * mean_l) / (n - 1.0);
    if (var_l < 0.0) var_l = 0.0;
    const double mean_i = stats[2 * frame] / n;
    double var_i = (stats[2 * frame + 1] - n * mean_i * mean_i) / (n - 1.0);
    if (var_i < 0.0) var_i = 0.0;
    const double a = sqrt(var_l) / (sqrt(var_i) + 1e-5);
    s_a = static_cast<float>(a);
    s_shift = static_cast<float>(mean_l - a * mean_i);
    row_scale[frame] = static_cast<float>(a);
  }
  __syncthreads();
  const float shift = s_shift;
  for (int n = threadIdx.x; n < C; n += blockDim.x)
    row_bias[static_cast<long long>(frame) * C + n] =
        wv[static_cast<long long>(b) * C + n] + bias[n] + shift * w_rowsum[n];
}

}  // namespace sa

using namespace sa;

extern "C" int sa_attn_temporal(const void* qkv, int32_t ld, int32_t B, int32_t F, int32_t S,
                                int32_t C, int32_t heads, void* out, int32_t ldo, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!qkv || !out) return fail(SA_ERR_INVALID, "sa_attn_temporal: null pointer");
  if (F < 1 || F > 16) return fail(SA_ERR_INVALID, "sa_attn_temporal: F=%d must be in [1,16]", F);
  if (C != heads * 64 || ld % 8 || ldo % 8) return fail(SA_ERR_INVALID, "sa_attn_temporal: head_dim must be 64");
  const long long ntasks = static_cast<long long>(B) * S * heads;
  const unsigned grid = static_cast<unsigned>((ntasks + 7) / 8);
  attn_temporal_kernel<<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(qkv), ld, B, F, S, C,
                                                 heads, static_cast<__nv_bfloat16*>(out), ldo);
  SA_CUDA_CHECK(cudaGetLastError());
  return SA_OK;
}

// q [B*F*S, ldq] bf16; ik, iv fp32 [B, 4, C]; out [B*F*S, ldo] bf16; stats fp64 [B*F, 2] (zeroed here)
extern "C" int sa_attn_id(const void* q, int32_t ldq, const float* ik, const float* iv,
                          int64_t kv_tok_stride, int64_t kv_batch_stride, int32_t B, int32_t F,
                          int32_t S, int32_t C, int32_t heads, int32_t num_tokens, void* out,
                          int32_t ldo, void* stats, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!q || !ik || !iv || !out || !stats) return fail(SA_ERR_INVALID, "sa_attn_id: null pointer");
  if (num_tokens != 4) return fail(SA_ERR_INVALID, "sa_attn_id: num_tokens must be 4");
  if (C != heads * 64 || ldq % 8 || ldo % 8) return fail(SA_ERR_INVALID, "sa_attn_id: head_dim must be 64");
  const size_t smem = sizeof(float) * 2 * 4 * C;
  static bool configured = false;
  if (!configured) {
    SA_CUDA_CHECK(cudaFuncSetAttribute(attn_id_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       96 * 1024));
    configured = true;
  }
  SA_CUDA_CHECK(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * B * F, stream));
  dim3 grid((S + 31) / 32, B * F, 1);
  attn_id_kernel<4><<<grid, 256, smem, stream>>>(static_cast<const __nv_bfloat16*>(q), ldq, ik, iv,
                                                 kv_tok_stride, kv_batch_stride, F, S, C,
                                                 heads, static_cast<__nv_bfloat16*>(out), ldo,
                                                 static_cast<double*>(stats));
  SA_CUDA_CHECK(cudaGetLastError());
  return SA_OK;
}

extern "C" int sa_attn_id_finalize(const void* stats, const float* vclip, int64_t vclip_batch_stride,
                                   const float* wv, const float* w_rowsum, const float* bias, int32_t B,
                                   int32_t F, int32_t S, int32_t C, float* row_scale, float* row_bias,
                                   void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!stats || !vclip || !wv || !w_rowsum || !bias || !row_scale || !row_bias)
    return fail(SA_ERR_INVALID, "sa_attn_id_finalize: null pointer");
  attn_id_finalize_kernel<<<B * F, 256, 0, stream>>>(static_cast<const double*>(stats), vclip,
                                                     vclip_batch_stride, wv, w_rowsum,
                                                     bias, F, S, C, row_scale, row_bias);
  SA_CUDA_CHECK(cudaGetLastError());
  return SA_OK;
}
