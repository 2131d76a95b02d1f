// Memory-bound normalisation kernels: GroupNorm(32) statistics + apply(+SiLU), LayerNorm.
// 128-bit vectorised bf16 access (8 channels per load), fp32 math, fp64 cross-thread accumulation
// of the GroupNorm sums.  Layout: [rows, C] channel-innermost (NHWC).
//
// Replaces nn.GroupNorm(+nn.SiLU) at animation/modules/unet.py:171-173,575,1632,1640,1643
// (per-frame statistics) and :1709,1722,1731 (TemporalResnetBlock: statistics over all frames of
// the tile), and nn.LayerNorm at :499,510,521,892,898,907,916.
#include "common.cuh"
#include "ptx.cuh"

namespace sa {

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 f = unpack_bf16x2(w[t]);
    v[2 * t] = f.x;
    v[2 * t + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                            pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}

// ------------------------------------------------------------------ GroupNorm statistics
// grid (chunks, NS); block (C/8, TY).  Stat group ns covers rows [ns*P, (ns+1)*P).
__global__ void gn_stats_kernel(const __nv_bfloat16* __restrict__ x0,
                                const __nv_bfloat16* __restrict__ x1, int C0, int C1, int ld0,
                                int ld1, long long P, int cpg, int rows_per_cta,
                                double* __restrict__ stats) {
  __shared__ double sm[32][2];
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  if (tid < 64) sm[tid >> 1][tid & 1] = 0.0;
  __syncthreads();
  const int c = threadIdx.x * 8;
  const long long ns = blockIdx.y;
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  long long r1 = r0 + rows_per_cta;
  if (r1 > P) r1 = P;
  const __nv_bfloat16* src;
  long long ld;
  if (c < C0) {
    src = x0 + c;
    ld = ld0;
  } else {
    src = x1 + (c - C0);
    ld = ld1;
  }
  float s[8], q[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s[k] = q[k] = 0.f;
  {
    // 4 independent 128-bit loads in flight per thread
    const long long step = blockDim.y;
    long long r = r0 + threadIdx.y;
    for (; r + 3 * step < r1; r += 4 * step) {
      float v[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) load8(src + (ns * P + r + u * step) * ld, v[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          s[k] += v[u][k];
          q[k] += v[u][k] * v[u][k];
        }
    }
    for (; r < r1; r += step) {
      float v[8];
      load8(src + (ns * P + r) * ld, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        s[k] += v[k];
        q[k] += v[k] * v[k];
      }
    }
  }
  // fold the 8 channels into their groups (an octet may straddle two groups)
  int g_prev = c / cpg;
  double as = 0.0, aq = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int g = (c + k) / cpg;
    if (g != g_prev) {
      atomicAdd(&sm[g_prev][0], as);
      atomicAdd(&sm[g_prev][1], aq);
      as = aq = 0.0;
      g_prev = g;
    }
    as += static_cast<double>(s[k]);
    aq += static_cast<double>(q[k]);
  }
  atomicAdd(&sm[g_prev][0], as);
  atomicAdd(&sm[g_prev][1], aq);
  __syncthreads();
  if (tid < 64) atomicAdd(&stats[(ns * 32 + (tid >> 1)) * 2 + (tid & 1)], sm[tid >> 1][tid & 1]);
}

// ------------------------------------------------------------------ GroupNorm apply (+SiLU)
__global__ void gn_apply_kernel(const __nv_bfloat16* __restrict__ x0,
                                const __nv_bfloat16* __restrict__ x1, int C0, int C1, int ld0,
                                int ld1, long long P, int cpg, int rows_per_cta,
                                const double* __restrict__ stats, const float* __restrict__ gamma,
                                const float* __restrict__ beta, float eps, int do_silu,
                                __nv_bfloat16* __restrict__ out, int ldo) {
  const int c = threadIdx.x * 8;
  const long long ns = blockIdx.y;
  const double n = static_cast<double>(P) * cpg;
  float a[8], b[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int g = (c + k) / cpg;
    const double mean = stats[(ns * 32 + g) * 2] / n;
    double var = stats[(ns * 32 + g) * 2 + 1] / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    const float ga = gamma[c + k];
    a[k] = rstd * ga;
    b[k] = beta[c + k] - static_cast<float>(mean) * rstd * ga;
  }
  const __nv_bfloat16* src;
  long long ld;
  if (c < C0) {
    src = x0 + c;
    ld = ld0;
  } else {
    src = x1 + (c - C0);
    ld = ld1;
  }
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  long long r1 = r0 + rows_per_cta;
  if (r1 > P) r1 = P;
  const long long step = blockDim.y;
  long long r = r0 + threadIdx.y;
  for (; r + 3 * step < r1; r += 4 * step) {
    float v[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) load8(src + (ns * P + r + u * step) * ld, v[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float y = v[u][k] * a[k] + b[k];
        v[u][k] = do_silu ? silu(y) : y;
      }
      store8(out + (ns * P + r + u * step) * ldo + c, v[u]);
    }
  }
  for (; r < r1; r += step) {
    float v[8];
    load8(src + (ns * P + r) * ld, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float y = v[k] * a[k] + b[k];
      v[k] = do_silu ? silu(y) : y;
    }
    store8(out + (ns * P + r) * ldo + c, v);
  }
}

// ------------------------------------------------------------------ LayerNorm
// LPR lanes per row (8, 16 or 32), OPL octets (8 channels, 16 B) per lane, LPR * OPL == C / 8, so a
// warp normalises 32 / LPR rows at once with every lane busy (C = 320 -> 8 lanes x 5 octets, 4 rows
// per warp).  y = LN(x + addv[row / rows_per_group]) * gamma + beta; optionally stores the pre-norm sum
// (the new residual stream) to x_out.  Two-pass variance in registers.
template <int LPR, int OPL>
__global__ void __launch_bounds__(256)
layernorm_kernel(const __nv_bfloat16* __restrict__ x, int ldx, long long M, int C,
                 const float* __restrict__ addv, int rows_per_group,
                 __nv_bfloat16* __restrict__ x_out, int ldxo, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ y, int ldy) {
  constexpr int RPW = 32 / LPR;  // rows per warp
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, l = lane % LPR;
  const long long warp_id = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long row = warp_id * RPW + sub;
  const bool active = row < M;
  const long long rr = active ? row : 0;
  float v[OPL][8];
  float sum = 0.f;
  const float* av = addv ? addv + (rr / rows_per_group) * C : nullptr;
#pragma unroll
  for (int o = 0; o < OPL; ++o) {
    const int oc = l + LPR * o;
    load8(x + rr * ldx + oc * 8, v[o]);
    if (av) {
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(av + oc * 8));
      const float4 a1 = __ldg(reinterpret_cast<const float4*>(av + oc * 8 + 4));
      v[o][0] += a0.x; v[o][1] += a0.y; v[o][2] += a0.z; v[o][3] += a0.w;
      v[o][4] += a1.x; v[o][5] += a1.y; v[o][6] += a1.z; v[o][7] += a1.w;
      if (x_out) {
        if (active) store8(x_out + rr * ldxo + oc * 8, v[o]);
        // keep the bf16-rounded value: it is what every later consumer of the stream reads
#pragma unroll
        for (int k = 0; k < 8; ++k) v[o][k] = __bfloat162float(__float2bfloat16_rn(v[o][k]));
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) sum += v[o][k];
  }
#pragma unroll
  for (int d = LPR / 2; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
  const float mean = sum / static_cast<float>(C);
  float var = 0.f;
#pragma unroll
  for (int o = 0; o < OPL; ++o) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float d = v[o][k] - mean;
      var += d * d;
    }
  }
#pragma unroll
  for (int d = LPR / 2; d > 0; d >>= 1) var += __shfl_xor_sync(0xffffffffu, var, d);
  const float rstd = rsqrtf(var / static_cast<float>(C) + eps);
  if (!active) return;
#pragma unroll
  for (int o = 0; o < OPL; ++o) {
    const int oc = l + LPR * o;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + oc * 8));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + oc * 8 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + oc * 8));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + oc * 8 + 4));
    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float o8[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o8[k] = fmaf((v[o][k] - mean) * rstd, g[k], bb[k]);
    store8(y + rr * ldy + oc * 8, o8);
  }
}

}  // namespace sa

using namespace sa;

static int gn_geometry(int C, long long P, int ns, dim3* block, dim3* grid, int* rows_per_cta) {
  const int c8 = C / 8;
  if (C % 32 || C % 8 || c8 > 1024) return -1;
  int ty = 256 / c8;
  if (ty < 1) ty = 1;
  *block = dim3(c8, ty, 1);
  // aim for ~8 CTAs per SM over the whole launch
  long long target_ctas = 8LL * num_sms();
  long long chunks = (target_ctas + ns - 1) / ns;
  if (chunks < 1) chunks = 1;
  long long rpc = (P + chunks - 1) / chunks;
  const long long min_rows = static_cast<long long>(ty) * 8;
  if (rpc < min_rows) rpc = min_rows;
  *rows_per_cta = static_cast<int>(rpc);
  *grid = dim3(static_cast<unsigned>((P + rpc - 1) / rpc), ns, 1);
  return 0;
}

// x0 [rows, C0] (+ optional x1 [rows, C1], channel concat) -> stats [ns, 32, 2] (fp64 sum, sumsq)
// then out [rows, C0+C1] = silu?(GN(x)).  rows = ns * rows_per_stat.
extern "C" int sa_groupnorm(const void* x0, const void* x1, int32_t c0, int32_t c1, int32_t ld0,
                            int32_t ld1, int64_t rows_per_stat, int32_t ns, const float* gamma,
                            const float* beta, float eps, int32_t do_silu, void* stats_ws,
                            void* out, int32_t ldo, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int C = c0 + c1;
  if (!x0 || !out || !stats_ws || !gamma || !beta) return fail(SA_ERR_INVALID, "sa_groupnorm: null pointer");
  if (c0 % 8 || c1 % 8 || (c1 > 0 && !x1)) return fail(SA_ERR_INVALID, "sa_groupnorm: bad channel split");
  dim3 block, grid;
  int rpc;
  if (gn_geometry(C, rows_per_stat, ns, &block, &grid, &rpc))
    return fail(SA_ERR_INVALID, "sa_groupnorm: C=%d unsupported", C);
  const int cpg = C / 32;
  SA_CUDA_CHECK(cudaMemsetAsync(stats_ws, 0, sizeof(double) * 64 * ns, stream));
  gn_stats_kernel<<<grid, block, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(x0), static_cast<const __nv_bfloat16*>(x1), c0, c1, ld0, ld1,
      rows_per_stat, cpg, rpc, static_cast<double*>(stats_ws));
  SA_CUDA_CHECK(cudaGetLastError());
  gn_apply_kernel<<<grid, block, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(x0), static_cast<const __nv_bfloat16*>(x1), c0, c1, ld0, ld1,
      rows_per_stat, cpg, rpc, static_cast<const double*>(stats_ws), gamma, beta, eps, do_silu,
      static_cast<__nv_bfloat16*>(out), ldo);
  SA_CUDA_CHECK(cudaGetLastError());
  return SA_OK;
}

extern "C" int sa_layernorm(const void* x, int32_t ldx, int64_t rows, int32_t c, const float* addv,
                            int32_t rows_per_group, void* x_out, int32_t ldxo, const float* gamma,
                            const float* beta, float eps, void* y, int32_t ldy, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !y || !gamma || !beta) return fail(SA_ERR_INVALID, "sa_layernorm: null pointer");
  if (c % 64 || c > 2048 || c <= 0) return fail(SA_ERR_INVALID, "sa_layernorm: C=%d unsupported (multiple of 64, <= 2048)", c);
  if (rows <= 0) return SA_OK;
  const int rpg = rows_per_group > 0 ? rows_per_group : 1;
  const int noct = c / 8;
  // lanes per row: the largest of 32/16/8 that divides C/8 with at most 8 octets per lane
  int lpr = 0;
  for (int cand = 32; cand >= 8; cand >>= 1)
    if (noct % cand == 0 && noct / cand <= 8) {
      lpr = cand;
      break;
    }
  if (!lpr) return fail(SA_ERR_INVALID, "sa_layernorm: C=%d unsupported", c);
  const int opl = noct / lpr;
  const int warps = 8, rpw = 32 / lpr;
  const unsigned grid = static_cast<unsigned>((rows + static_cast<long long>(warps) * rpw - 1) / (warps * rpw));
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* xo = static_cast<__nv_bfloat16*>(x_out);
  __nv_bfloat16* yp = static_cast<__nv_bfloat16*>(y);
#define SA_LN(LPR, OPL)                                                                             \
  layernorm_kernel<LPR, OPL><<<grid, warps * 32, 0, stream>>>(xp, ldx, rows, c, addv, rpg, xo, ldxo, \
                                                              gamma, beta, eps, yp, ldy)
#define SA_LN_ROW(LPR)                                                         \
  switch (opl) {                                                               \
    case 1: SA_LN(LPR, 1); break;                                              \
    case 2: SA_LN(LPR, 2); break;                                              \
    case 3: SA_LN(LPR, 3); break;                                              \
    case 4: SA_LN(LPR, 4); break;                                              \
    case 5: SA_LN(LPR, 5); break;                                              \
    case 6: SA_LN(LPR, 6); break;                                              \
    case 7: SA_LN(LPR, 7); break;                                              \
    default: SA_LN(LPR, 8); break;                                             \
  }
  if (lpr == 32) { SA_LN_ROW(32) } else if (lpr == 16) { SA_LN_ROW(16) } else { SA_LN_ROW(8) }
#undef SA_LN_ROW
#undef SA_LN
  SA_CUDA_CHECK(cudaGetLastError());
  return SA_OK;
}
