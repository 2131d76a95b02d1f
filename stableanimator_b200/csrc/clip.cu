// Image-conditioning front end: the arithmetic of InferenceAnimationPipeline._encode_image
// (animation/pipelines/inference_pipeline_animation.py:192-244) that is not a GEMM / LayerNorm:
//   * _resize_with_antialiasing (:99-125): separable Gaussian blur with reflect padding (:44-96), then bicubic
//     interpolation with align_corners=True, fused with the (x + 1) / 2 un-normalisation (:210) and the
//     CLIPImageProcessor mean / std normalisation (:213-220);
//   * the patch gather of CLIPVisionEmbeddings (Conv2d(3 -> C, k = stride = patch), third-party transformers) as an
//     im2col so the projection runs on sa_gemm_tap;
//   * multi-head self-attention over the 257 tokens of the vision tower (head_dim 80): a few hundred MFLOP per
//     layer, once per video -> fp32 CUDA-core kernel with the head's K / V resident in shared memory.
// All of it runs once per video (not per denoising step); these kernels are sized for latency, not throughput.
#include <math.h>

#include "common.cuh"
#include "ptx.cuh"

namespace sa {

__device__ __forceinline__ int reflect_idx(int i, int n) {
  // F.pad(mode="reflect"): -1 -> 1, n -> n - 2 (pad < n guaranteed by the launcher)
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}

// out[p, y, x] = sum_t taps[t] * (a * in[p, reflect(y + t - pad)] + c) along rows (axis 0) or columns (axis 1)
__global__ void __launch_bounds__(256)
gauss_blur_axis_kernel(const float* __restrict__ in, float* __restrict__ out, int planes, int h, int w,
                       const float* __restrict__ taps, int k, int axis, float a, float c) {
  const long long total = static_cast<long long>(planes) * h * w;
  const int pad = (k - 1) / 2;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % w);
    const int y = static_cast<int>((i / w) % h);
    const float* plane = in + (i / (static_cast<long long>(h) * w)) * h * w;
    float acc = 0.f;
    if (axis == 1) {
      for (int t = 0; t < k; ++t) acc += __ldg(taps + t) * fmaf(a, __ldg(plane + y * w + reflect_idx(x + t - pad, w)), c);
    } else {
      for (int t = 0; t < k; ++t) acc += __ldg(taps + t) * fmaf(a, __ldg(plane + reflect_idx(y + t - pad, h) * w + x), c);
    }
    out[i] = acc;
  }
}

// cubic convolution weights, A = -0.75 (torch upsample_bicubic2d)
__device__ __forceinline__ void cubic_coeffs(float t, float c[4]) {
  const float A = -0.75f;
  const float x0 = t + 1.0f, x3 = 2.0f - t, x2 = 1.0f - t;
  c[0] = ((A * x0 - 5.0f * A) * x0 + 8.0f * A) * x0 - 4.0f * A;
  c[1] = ((A + 2.0f) * t - (A + 3.0f)) * t * t + 1.0f;
  c[2] = ((A + 2.0f) * x2 - (A + 3.0f)) * x2 * x2 + 1.0f;
  c[3] = ((A * x3 - 5.0f * A) * x3 + 8.0f * A) * x3 - 4.0f * A;
}

// out[b, c, oy, ox] = ((bicubic(in)[b, c, oy, ox] + 1) / 2 - mean[c]) / std[c], align_corners = True
__global__ void __launch_bounds__(256)
resize_bicubic_norm_kernel(const float* __restrict__ in, int b, int c, int h, int w, int oh, int ow, float m0,
                           float m1, float m2, float s0, float s1, float s2, float* __restrict__ out) {
  const long long total = static_cast<long long>(b) * c * oh * ow;
  const float sy = oh > 1 ? static_cast<float>(h - 1) / static_cast<float>(oh - 1) : 0.f;
  const float sx = ow > 1 ? static_cast<float>(w - 1) / static_cast<float>(ow - 1) : 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ox = static_cast<int>(i % ow);
    const int oy = static_cast<int>((i / ow) % oh);
    const int ch = static_cast<int>((i / (static_cast<long long>(ow) * oh)) % c);
    const float* plane = in + (i / (static_cast<long long>(ow) * oh)) * h * w;
    const float fy = sy * oy, fx = sx * ox;
    const int iy = static_cast<int>(floorf(fy)), ix = static_cast<int>(floorf(fx));
    float cy[4], cx[4];
    cubic_coeffs(fy - iy, cy);
    cubic_coeffs(fx - ix, cx);
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int yy = min(max(iy - 1 + r, 0), h - 1);
      float row = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) row += cx[q] * __ldg(plane + yy * w + min(max(ix - 1 + q, 0), w - 1));
      acc += cy[r] * row;
    }
    const float mean = ch == 0 ? m0 : (ch == 1 ? m1 : m2);
    const float sd = ch == 0 ? s0 : (ch == 1 ? s1 : s2);
    out[i] = ((acc + 1.0f) / 2.0f - mean) / sd;
  }
}

// out[(b*np + py*npx + px), (ci*p + ky)*p + kx] = bf16(pix[b, ci, py*p + ky, px*p + kx]); columns >= c*p*p are zero
__global__ void __launch_bounds__(256)
patchify_kernel(const float* __restrict__ pix, int b, int c, int hw, int p, __nv_bfloat16* __restrict__ out,
                int ldo) {
  const int npx = hw / p;
  const long long total = static_cast<long long>(b) * npx * npx * ldo;
  const int kk = c * p * p;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int col = static_cast<int>(i % ldo);
    const long long row = i / ldo;
    float v = 0.f;
    if (col < kk) {
      const int kx = col % p, ky = (col / p) % p, ci = col / (p * p);
      const int px = static_cast<int>(row % npx), py = static_cast<int>((row / npx) % npx);
      const long long img = row / (static_cast<long long>(npx) * npx);
      v = __ldg(pix + ((img * c + ci) * hw + py * p + ky) * hw + px * p + kx);
    }
    out[i] = __float2bfloat16(v);
  }
}

// ----------------------------------------------------------------------------------------------------------
// softmax(q k^T * scale) v per (image, head) over n tokens, head_dim d (even, <= 128), fp32 math on bf16 q|k|v.
// Block = (head, image, 32-query slab), 8 warps; K (rows padded to an odd word count: conflict-free column walks)
// and V of the head live in shared memory; a warp owns one query at a time: lanes stride the keys for the scores,
// then stride the head dimension for P V.
// ----------------------------------------------------------------------------------------------------------
constexpr int AT_WARPS = 8, AT_QB = 32;

__global__ void __launch_bounds__(AT_WARPS * 32)
attn_tokens_kernel(const __nv_bfloat16* __restrict__ qkv, int ld, int n, int heads, int d, float scale,
                   __nv_bfloat16* __restrict__ out, int ldo) {
  extern __shared__ __align__(16) uint8_t at_smem[];
  const int C = heads * d;
  const int head = blockIdx.x, img = blockIdx.y;
  const int kw = d / 2 + 1;                                   // K row stride in 32-bit words (odd)
  uint32_t* sK = reinterpret_cast<uint32_t*>(at_smem);        // [n][kw] bf16 pairs
  __nv_bfloat16* sV = reinterpret_cast<__nv_bfloat16*>(sK + static_cast<size_t>(n) * kw);  // [n][d]
  float* sP = reinterpret_cast<float*>(sV + static_cast<size_t>(n) * d + (static_cast<size_t>(n) * d & 1));
  float* sQ = sP + AT_WARPS * n;                              // [warps][d]
  const __nv_bfloat16* base = qkv + static_cast<long long>(img) * n * ld;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < n * (d / 2); i += blockDim.x) {
    const int j = i / (d / 2), c = i % (d / 2);
    sK[j * kw + c] = *reinterpret_cast<const uint32_t*>(base + static_cast<long long>(j) * ld + C + head * d + 2 * c);
    *reinterpret_cast<uint32_t*>(sV + static_cast<size_t>(j) * d + 2 * c) =
        *reinterpret_cast<const uint32_t*>(base + static_cast<long long>(j) * ld + 2 * C + head * d + 2 * c);
  }
  __syncthreads();

  float* p = sP + warp * n;
  float* q = sQ + warp * d;
  const int q_end = min(n, (static_cast<int>(blockIdx.z) + 1) * AT_QB);
  for (int qi = blockIdx.z * AT_QB + warp; qi < q_end; qi += AT_WARPS) {
    for (int c = lane; c < d; c += 32)
      q[c] = __bfloat162float(base[static_cast<long long>(qi) * ld + head * d + c]) * scale;
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j < n; j += 32) {
      const uint32_t* kr = sK + j * kw;
      float s = 0.f;
      for (int c = 0; c < d / 2; ++c) {
        const float2 kf = unpack_bf16x2(kr[c]);
        s = fmaf(q[2 * c], kf.x, s);
        s = fmaf(q[2 * c + 1], kf.y, s);
      }
      p[j] = s;
      mx = fmaxf(mx, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < n; j += 32) {
      const float e = expf(p[j] - mx);
      p[j] = e;
      sum += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncwarp();
    const float inv = 1.0f / sum;
    for (int c = lane; c < d; c += 32) {
      float acc = 0.f;
      for (int j = 0; j < n; ++j) acc = fmaf(p[j], __bfloat162float(sV[static_cast<size_t>(j) * d + c]), acc);
      out[(static_cast<long long>(img) * n + qi) * ldo + head * d + c] = __float2bfloat16(acc * inv);
    }
    __syncwarp();
  }
}

static int grid_for(long long total) {
  const long long blocks = (total + 255) / 256;
  return static_cast<int>(blocks < 1 ? 1 : (blocks > 148 * 16 ? 148 * 16 : blocks));
}

}  // namespace sa

using namespace sa;

extern "C" int sa_gauss_blur_axis(const float* in, float* out, int32_t planes, int32_t h, int32_t w,
                                  const float* taps, int32_t k, int32_t axis, float in_scale, float in_shift,
                                  void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!in || !out || !taps || in == out) return fail(SA_ERR_INVALID, "sa_gauss_blur_axis: bad pointers");
  if (planes < 1 || h < 1 || w < 1 || k < 1 || (k & 1) == 0 || (axis != 0 && axis != 1))
    return fail(SA_ERR_INVALID, "sa_gauss_blur_axis: bad sizes (k must be odd)");
  if ((k - 1) / 2 >= (axis == 1 ? w : h)) return fail(SA_ERR_INVALID, "sa_gauss_blur_axis: reflect padding >= extent");
  gauss_blur_axis_kernel<<<grid_for(static_cast<long long>(planes) * h * w), 256, 0, stream>>>(in, out, planes, h, w,
                                                                                                 taps, k, axis, in_scale,
                                                                                                 in_shift);
  SA_CUDA_CHECK(cudaGetLastError());
  return SA_OK;
}

extern "C" int sa_resize_bicubic_norm(const float* in, int32_t b, int32_t c, int32_t h, int32_t w, int32_t oh,
                                      int32_t ow, const float* mean3, const float* std3, float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!in || !out || !mean3 || !std3) return fail(SA_ERR_INVALID, "sa_resize_bicubic_norm: null pointer");
  if (b < 1 || c != 3 || h < 1 || w < 1 || oh < 1 || ow < 1)
    return fail(SA_ERR_INVALID, "sa_resize_bicubic_norm: bad sizes (c must be 3)");
  resize_bicubic_norm_kernel<<<grid_for(static_cast<long long>(b) * c * oh * ow), 256, 0, stream>>>(
      in, b, c, h, w, oh, ow, mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2], out);
  SA_CUDA_CHECK(cudaGetLastError());
  return SA_OK;
}

extern "C" int sa_patchify(const float* pix, int32_t b, int32_t c, int32_t hw, int32_t patch, void* out, int32_t ldo,
                           void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!pix || !out) return fail(SA_ERR_INVALID, "sa_patchify: null pointer");
  if (b < 1 || c < 1 || patch < 1 || hw < patch || hw % patch || ldo < c * patch * patch)
    return fail(SA_ERR_INVALID, "sa_patchify: bad sizes");
  const int np = (hw / patch) * (hw / patch);
  patchify_kernel<<<grid_for(static_cast<long long>(b) * np * ldo), 256, 0, stream>>>(
      pix, b, c, hw, patch, static_cast<__nv_bfloat16*>(out), ldo);
  SA_CUDA_CHECK(cudaGetLastError());
  return SA_OK;
}

extern "C" int sa_attn_tokens(const void* qkv, int32_t ld, int32_t b, int32_t n, int32_t heads, int32_t d,
                              float scale, void* out, int32_t ldo, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!qkv || !out) return fail(SA_ERR_INVALID, "sa_attn_tokens: null pointer");
  if (b < 1 || b > 65535 || n < 1 || heads < 1 || d < 2 || d > 128 || (d & 1) || ld < 3 * heads * d || (ld & 1) ||
      ldo < heads * d)
    return fail(SA_ERR_INVALID, "sa_attn_tokens: bad sizes (head_dim even and <= 128, ld >= 3*heads*d)");
  size_t smem = static_cast<size_t>(n) * (d / 2 + 1) * 4 + (static_cast<size_t>(n) * d + (static_cast<size_t>(n) * d & 1)) * 2 +
                static_cast<size_t>(AT_WARPS) * (n + d) * 4;
  if (smem > 227 * 1024) return fail(SA_ERR_INVALID, "sa_attn_tokens: %d tokens x head_dim %d do not fit shared memory", n, d);
  // the whole opt-in range once per device (the requested size varies with the token count)
  static DeviceOnce once;
  if (int rc = once.run([&]() -> int {
        SA_CUDA_CHECK(cudaFuncSetAttribute(attn_tokens_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        return SA_OK;
      }))
    return rc;
  dim3 grid(heads, b, (n + AT_QB - 1) / AT_QB);
  attn_tokens_kernel<<<grid, AT_WARPS * 32, smem, stream>>>(static_cast<const __nv_bfloat16*>(qkv), ld, n, heads, d,
                                                            scale, static_cast<__nv_bfloat16*>(out), ldo);
  SA_CUDA_CHECK(cudaGetLastError());
  return SA_OK;
}
