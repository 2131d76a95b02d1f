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

// One warp per (batch b, pixel s, head h): the 16x16 attention over frames maps exactly onto
// mma.sync m16n8k16 tiles (bf16 in, fp32 accumulate) -- the problem is far too small for tcgen05 and is
// HBM-bound; the legacy warp-level MMA just removes the ~2000 FMA instructions per task that made the
// first version instruction-bound.  Per task: cp.async the three [F x 64] tiles (128-byte rows, XOR
// swizzled) -> S = Q K^T (8 MMAs) -> softmax in the accumulator fragments -> P as the A fragment of
// O = P V (8 MMAs, V via ldmatrix.trans) -> O staged through smem -> 16-byte coalesced stores.
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_f16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                              uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <bool F16>
__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
  if constexpr (F16) mma_f16_16816(c, a0, a1, a2, a3, b0, b1);
  else mma_bf16_16816(c, a0, a1, a2, a3, b0, b1);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

constexpr int AT_WARPS = 8;
constexpr int AT_TILE = 16 * 128;  // one [16 frames x 64 dims] bf16 tile

template <bool F16>   // F16: q | k | v and out hold IEEE fp16 instead of bf16
__global__ void __launch_bounds__(AT_WARPS * 32)
attn_temporal_kernel(const __nv_bfloat16* __restrict__ qkv, int ld, int B, int F, int S, int C,
                     int heads, __nv_bfloat16* __restrict__ out, int ldo) {
  grid_dep_sync_light();
  __shared__ __align__(128) uint8_t sm[AT_WARPS][3][AT_TILE];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long task = static_cast<long long>(blockIdx.x) * AT_WARPS + warp;
  const long long ntasks = static_cast<long long>(B) * S * heads;
  if (task >= ntasks) return;
  const int h = static_cast<int>(task % heads);
  const long long bs = task / heads;
  const int s = static_cast<int>(bs % S);
  const int b = static_cast<int>(bs / S);
  const long long row0 = static_cast<long long>(b) * F * S + s;  // global row of frame 0
  uint8_t* sq = sm[warp][0];
  uint8_t* sk = sm[warp][1];
  uint8_t* sv = sm[warp][2];

  // ---- load: chunk id = lane + 32*i -> (frame = id / 8, 16-byte chunk = id % 8); rows >= F are zeroed
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int id = lane + 32 * i;
    const int f = id >> 3, ch = id & 7;
    const uint32_t off = f * 128 + ((ch ^ (f & 7)) << 4);
    if (f < F) {
      const __nv_bfloat16* g = qkv + (row0 + static_cast<long long>(f) * S) * ld + h * 64 + ch * 8;
      cp_async16(smem_u32(sq + off), g);
      cp_async16(smem_u32(sk + off), g + C);
      cp_async16(smem_u32(sv + off), g + 2 * C);
    } else {
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(sq + off) = z;
      *reinterpret_cast<uint4*>(sk + off) = z;
      *reinterpret_cast<uint4*>(sv + off) = z;
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();

  // ---- S = Q K^T : two n-tiles (keys 0-7, 8-15), 4 k-steps over the 64 dims
  float sc[2][4];
#pragma unroll
  for (int n = 0; n < 2; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) sc[n][e] = 0.f;
  {
    const int ar = (lane & 7) + 8 * ((lane >> 3) & 1);   // A: row, k-half = lane >> 4
    const int br = (lane & 7) + 8 * (lane >> 4);          // B (K rows = keys): key row, k-half = (lane >> 3) & 1
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t a0, a1, a2, a3, b0, b1, b2, b3;
      ldsm_x4(smem_u32(sq + ar * 128 + (((2 * ks + (lane >> 4)) ^ (ar & 7)) << 4)), a0, a1, a2, a3);
      ldsm_x4(smem_u32(sk + br * 128 + (((2 * ks + ((lane >> 3) & 1)) ^ (br & 7)) << 4)), b0, b1, b2, b3);
      mma_16816<F16>(sc[0], a0, a1, a2, a3, b0, b1);
      mma_16816<F16>(sc[1], a0, a1, a2, a3, b2, b3);
    }
  }
  // ---- softmax over keys (columns); fragment: e0,e1 -> row lane/4, cols 8n + 2(lane%4) + {0,1};
  //      e2,e3 -> row lane/4 + 8.  attn.scale = 64^-0.5 folded into the exp2 argument.
  const float sl2 = 0.125f * 1.4426950408889634f;
  float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
  for (int n = 0; n < 2; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int col = 8 * n + 2 * (lane & 3) + (e & 1);
      if (col >= F) sc[n][e] = -INFINITY;
      mx[e >> 1] = fmaxf(mx[e >> 1], sc[n][e]);
    }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
    mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
  }
  float sum[2] = {0.f, 0.f};
#pragma unroll
  for (int n = 0; n < 2; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      sc[n][e] = ex2_approx((sc[n][e] - mx[e >> 1]) * sl2);
      sum[e >> 1] += sc[n][e];
    }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 1);
    sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 2);
  }
  const float inv0 = 1.0f / sum[0], inv1 = 1.0f / sum[1];
  // P as the A fragment of the PV MMA (k = keys): the two S n-tiles are exactly the k-halves.
  // P is split into bf16 hi + lo parts (two MMAs per tile, free here: the kernel is HBM-bound) so the
  // probabilities carry ~16 significant bits instead of 8.
  float pn[8] = {sc[0][0] * inv0, sc[0][1] * inv0, sc[0][2] * inv1, sc[0][3] * inv1,
                 sc[1][0] * inv0, sc[1][1] * inv0, sc[1][2] * inv1, sc[1][3] * inv1};
  uint32_t ph[4], pl[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    ph[i] = pack16x2(pn[2 * i], pn[2 * i + 1], F16);
    const float2 hf = unpack16x2(ph[i], F16);
    pl[i] = pack16x2(pn[2 * i] - hf.x, pn[2 * i + 1] - hf.y, F16);
  }

  // ---- O = P V : 8 n-tiles of 8 dims, one k-step (16 keys); V^T fragments via ldmatrix.trans
  float o[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[n][e] = 0.f;
  {
    const int vr = (lane & 7) + 8 * ((lane >> 3) & 1);  // key row; n-tile pair member = lane >> 4
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(smem_u32(sv + vr * 128 + (((2 * np + (lane >> 4)) ^ (vr & 7)) << 4)), b0, b1, b2, b3);
      mma_16816<F16>(o[2 * np], ph[0], ph[1], ph[2], ph[3], b0, b1);
      mma_16816<F16>(o[2 * np], pl[0], pl[1], pl[2], pl[3], b0, b1);
      mma_16816<F16>(o[2 * np + 1], ph[0], ph[1], ph[2], ph[3], b2, b3);
      mma_16816<F16>(o[2 * np + 1], pl[0], pl[1], pl[2], pl[3], b2, b3);
    }
  }
  // ---- stage O (bf16) in the Q tile, then coalesced 16-byte stores
  __syncwarp();
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const int r = lane >> 2;
    *reinterpret_cast<uint32_t*>(sq + r * 128 + ((n ^ (r & 7)) << 4) + (lane & 3) * 4) = pack16x2(o[n][0], o[n][1], F16);
    *reinterpret_cast<uint32_t*>(sq + (r + 8) * 128 + ((n ^ ((r + 8) & 7)) << 4) + (lane & 3) * 4) =
        pack16x2(o[n][2], o[n][3], F16);
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int id = lane + 32 * i;
    const int f = id >> 3, ch = id & 7;
    if (f < F) {
      const uint4 v = *reinterpret_cast<const uint4*>(sq + f * 128 + ((ch ^ (f & 7)) << 4));
      *reinterpret_cast<uint4*>(out + (row0 + static_cast<long long>(f) * S) * ldo + h * 64 + ch * 8) = v;
    }
  }
}

// ID cross-attention: grid (ceil(S / rc), B*F); thread = (row lane, channel octet): the octet's K / V values of the four
// tokens stay in registers for all rows of the CTA, q is read and the result written with one coalesced 16-byte access per
// thread and row, and the 64-channel dot products of a head are reduced over its 8 neighbouring lanes (three shuffles); RPT rows are loaded before the first is used.
// ik / iv: fp32 [B, NTOK, C].  (Round 1: thread = (row, head) with 128 registers of q / o per thread, one CTA of 8 warps
// per SM and 640-byte lane strides: 1.7 TB/s.)
template <int NTOK, int RPT, bool F16>
__global__ void __launch_bounds__(320, 2)
attn_id_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const float* __restrict__ ik,
               const float* __restrict__ iv, long long tok_stride, long long batch_stride, int F, int S,
               int C, int rl_n, int nb,
               __nv_bfloat16* __restrict__ out, int ldo, double* __restrict__ stats) {
  grid_dep_sync_light();
  constexpr bool f16 = F16;   // q / out hold IEEE fp16 instead of bf16
  __shared__ double red[2][32];
  const int frame = blockIdx.y;
  const int b = frame / F;
  const int noct = C >> 3;
  const int oc = threadIdx.x % noct, rl = threadIdx.x / noct;
  float kk[NTOK][8], vv[NTOK][8];
  const bool kv_vec = ((reinterpret_cast<uintptr_t>(ik) | reinterpret_cast<uintptr_t>(iv)) & 15) == 0 && (tok_stride & 3) == 0 &&
                      (batch_stride & 3) == 0;
#pragma unroll
  for (int j = 0; j < NTOK; ++j) {
    const long long off = static_cast<long long>(b) * batch_stride + j * tok_stride + oc * 8;
    if (kv_vec) {
      const float4 k0 = __ldg(reinterpret_cast<const float4*>(ik + off)), k1 = __ldg(reinterpret_cast<const float4*>(ik + off + 4));
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(iv + off)), v1 = __ldg(reinterpret_cast<const float4*>(iv + off + 4));
      kk[j][0] = k0.x; kk[j][1] = k0.y; kk[j][2] = k0.z; kk[j][3] = k0.w;
      kk[j][4] = k1.x; kk[j][5] = k1.y; kk[j][6] = k1.z; kk[j][7] = k1.w;
      vv[j][0] = v0.x; vv[j][1] = v0.y; vv[j][2] = v0.z; vv[j][3] = v0.w;
      vv[j][4] = v1.x; vv[j][5] = v1.y; vv[j][6] = v1.z; vv[j][7] = v1.w;
    } else {
#pragma unroll
      for (int d = 0; d < 8; ++d) {   // views with unaligned strides: scalar loads
        kk[j][d] = __ldg(ik + off + d);
        vv[j][d] = __ldg(iv + off + d);
      }
    }
#pragma unroll
    for (int d = 0; d < 8; ++d) kk[j][d] *= 0.125f;   // the 1/sqrt(64) of the scores is folded into K
  }
  // RPT rows per thread, all q loads issued before the first is used: with one load in flight per thread the kernel
  // ran at the memory latency (1.7 TB/s)
  double lsum = 0.0, lsq = 0.0;
  for (int bt = 0; bt < nb; ++bt) {   // nb batches of rl_n * RPT rows per CTA: K / V are loaded once for all of them
  const int s0 = (blockIdx.x * nb + bt) * (rl_n * RPT);
  if (s0 >= S) break;
  float fsum = 0.f, fsq = 0.f;   // this thread's RPT rows x 8 channels in fp32, then fp64
  uint4 qraw[RPT];
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    const int sidx = s0 + rl + i * rl_n;
    qraw[i] = make_uint4(0u, 0u, 0u, 0u);
    if (sidx < S) qraw[i] = __ldg(reinterpret_cast<const uint4*>(q + (static_cast<long long>(frame) * S + sidx) * ldq + oc * 8));
  }
#pragma unroll
  for (int i = 0; i < RPT; ++i) {   // same trip count for every lane of a warp (the shuffles below)
    const int sidx = s0 + rl + i * rl_n;
    const bool valid = sidx < S;
    const long long row = static_cast<long long>(frame) * S + (valid ? sidx : 0);
    float qv[8];
    {
      const uint32_t w[4] = {qraw[i].x, qraw[i].y, qraw[i].z, qraw[i].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = unpack16x2(w[k], f16);
        qv[2 * k] = f.x;
        qv[2 * k + 1] = f.y;
      }
    }
    float sc[NTOK];
#pragma unroll
    for (int j = 0; j < NTOK; ++j) {
      float p = 0.f;
#pragma unroll
      for (int d = 0; d < 8; ++d) p = fmaf(qv[d], kk[j][d], p);
      sc[j] = p;
    }
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) {
#pragma unroll
      for (int j = 0; j < NTOK; ++j) sc[j] += __shfl_xor_sync(0xffffffffu, sc[j], d);
    }
    float mx = sc[0];
#pragma unroll
    for (int j = 1; j < NTOK; ++j) mx = fmaxf(mx, sc[j]);
    float l = 0.f;
#pragma unroll
    for (int j = 0; j < NTOK; ++j) {
      sc[j] = __expf(sc[j] - mx);
      l += sc[j];
    }
    const float inv = 1.0f / l;
    float o[8];
#pragma unroll
    for (int d = 0; d < 8; ++d) o[d] = 0.f;
#pragma unroll
    for (int j = 0; j < NTOK; ++j) {
      const float p = sc[j] * inv;
#pragma unroll
      for (int d = 0; d < 8; ++d) o[d] = fmaf(p, vv[j][d], o[d]);
    }
    if (valid) {
#pragma unroll
      for (int d = 0; d < 8; ++d) {
        fsum += o[d];
        fsq = fmaf(o[d], o[d], fsq);
      }
      *reinterpret_cast<uint4*>(out + row * ldo + oc * 8) =
          make_uint4(pack16x2(o[0], o[1], f16), pack16x2(o[2], o[3], f16), pack16x2(o[4], o[5], f16), pack16x2(o[6], o[7], f16));
    }
  }
  lsum += static_cast<double>(fsum);
  lsq += static_cast<double>(fsq);
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
  grid_dep_sync_light();
  __shared__ double red[2][8];
  __shared__ float s_shift;
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
                                int32_t C, int32_t heads, void* out, int32_t ldo, int32_t act_f16, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!qkv || !out) return fail(SA_ERR_INVALID, "sa_attn_temporal: null pointer");
  if (F < 1 || F > 16) return fail(SA_ERR_INVALID, "sa_attn_temporal: F=%d must be in [1,16]", F);
  if (C != heads * 64 || ld % 8 || ldo % 8) return fail(SA_ERR_INVALID, "sa_attn_temporal: head_dim must be 64");
  const long long ntasks = static_cast<long long>(B) * S * heads;
  const unsigned grid = static_cast<unsigned>((ntasks + AT_WARPS - 1) / AT_WARPS);
  if (act_f16)
    SA_CUDA_CHECK(launch_k(attn_temporal_kernel<true>, dim3(grid), dim3(AT_WARPS * 32), 0, stream, 1, false,
                           static_cast<const __nv_bfloat16*>(qkv), ld, B, F, S, C, heads, static_cast<__nv_bfloat16*>(out), ldo));
  else
    SA_CUDA_CHECK(launch_k(attn_temporal_kernel<false>, dim3(grid), dim3(AT_WARPS * 32), 0, stream, 1, false,
                           static_cast<const __nv_bfloat16*>(qkv), ld, B, F, S, C, heads, static_cast<__nv_bfloat16*>(out), ldo));
  return SA_OK;
}

// q [B*F*S, ldq] bf16; ik, iv fp32 [B, 4, C]; out [B*F*S, ldo] bf16; stats fp64 [B*F, 2] (zeroed here)
extern "C" int sa_attn_id(const void* q, int32_t ldq, const float* ik, const float* iv,
                          int64_t kv_tok_stride, int64_t kv_batch_stride, int32_t B, int32_t F,
                          int32_t S, int32_t C, int32_t heads, int32_t num_tokens, void* out,
                          int32_t ldo, void* stats, int32_t act_f16, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!q || !ik || !iv || !out || !stats) return fail(SA_ERR_INVALID, "sa_attn_id: null pointer");
  if (num_tokens != 4) return fail(SA_ERR_INVALID, "sa_attn_id: num_tokens must be 4");
  if (C != heads * 64 || ldq % 8 || ldo % 8) return fail(SA_ERR_INVALID, "sa_attn_id: head_dim must be 64");
  // thread = (row lane, channel octet): ~320 threads, a whole number of warps, every 8-lane head group inside one warp
  const int noct = C / 8;
  if (noct > 320) return fail(SA_ERR_INVALID, "sa_attn_id: C=%d unsupported (<= 2560)", C);
  int rl_n = 320 / noct;
  while (rl_n > 1 && (rl_n * noct) % 32) --rl_n;
  if ((rl_n * noct) % 32) return fail(SA_ERR_INVALID, "sa_attn_id: C=%d unsupported", C);
  SA_CUDA_CHECK(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * B * F, stream));
  // 2 rows per thread (4 spill under the 2-CTA register cap: 82 vs 67 us at [131072, 320]) and up to 8 batches per CTA while
  // the launch still fills the GPU a few times over
  const long long want = 4ll * num_sms();
  const bool big = static_cast<long long>((S + rl_n * 2 - 1) / (rl_n * 2)) * B * F >= want;
  const int rows_b = rl_n * (big ? 2 : 1);
  int nb = 8;
  while (nb > 1 && static_cast<long long>((S + rows_b * nb - 1) / (rows_b * nb)) * B * F < want) nb >>= 1;
  dim3 grid((S + rows_b * nb - 1) / (rows_b * nb), B * F, 1);
  auto* kern = big ? (act_f16 ? attn_id_kernel<4, 2, true> : attn_id_kernel<4, 2, false>)
                   : (act_f16 ? attn_id_kernel<4, 1, true> : attn_id_kernel<4, 1, false>);
  SA_CUDA_CHECK(launch_k(kern, grid, dim3(rl_n * noct), 0, stream, 1, false, static_cast<const __nv_bfloat16*>(q), ldq, ik, iv,
                         static_cast<long long>(kv_tok_stride), static_cast<long long>(kv_batch_stride), F, S, C, rl_n, nb,
                         static_cast<__nv_bfloat16*>(out), ldo, static_cast<double*>(stats)));
  return SA_OK;
}

extern "C" int sa_attn_id_finalize(const void* stats, const float* vclip, int64_t vclip_batch_stride,
                                   const float* wv, const float* w_rowsum, const float* bias, int32_t B,
                                   int32_t F, int32_t S, int32_t C, float* row_scale, float* row_bias,
                                   void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!stats || !vclip || !wv || !w_rowsum || !bias || !row_scale || !row_bias)
    return fail(SA_ERR_INVALID, "sa_attn_id_finalize: null pointer");
  SA_CUDA_CHECK(launch_k(attn_id_finalize_kernel, dim3(B * F), dim3(256), 0, stream, 1, false, static_cast<const double*>(stats), vclip,
                         static_cast<long long>(vclip_batch_stride), wv, w_rowsum, bias, F, S, C, row_scale, row_bias));
  return SA_OK;
}
