// Memory-bound normalisation kernels: GroupNorm(32) statistics + apply(+SiLU), LayerNorm.
// 128-bit vectorised bf16 access (8 channels per load), fp32 math, fp64 cross-thread accumulation
// of the GroupNorm sums.  Layout: [rows, C] channel-innermost (NHWC).
//
// Replaces nn.GroupNorm(+nn.SiLU) at animation/modules/unet.py:171-173,575,1632,1640,1643
// (per-frame statistics) and :1709,1722,1731 (TemporalResnetBlock: statistics over all frames of
// the tile), and nn.LayerNorm at :499,510,521,892,898,907,916.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"

namespace sa {

// `f16`: the 16-bit tensors hold IEEE fp16 instead of bf16 (warp-uniform run-time flag; pointers stay typed bf16)
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8], bool f16) {
  const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 f = unpack16x2(w[t], f16);
    v[2 * t] = f.x;
    v[2 * t + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8], bool f16) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack16x2(v[0], v[1], f16), pack16x2(v[2], v[3], f16),
                                            pack16x2(v[4], v[5], f16), pack16x2(v[6], v[7], f16));
}

// ------------------------------------------------------------------ GroupNorm statistics
// grid (chunks, NS); block (C/8, TY).  Stat group ns covers rows [ns*P, (ns+1)*P).
__global__ void gn_stats_kernel(const __nv_bfloat16* __restrict__ x0,
                                const __nv_bfloat16* __restrict__ x1, int C0, int C1, int ld0,
                                int ld1, long long P, int cpg, int rows_per_cta,
                                double* __restrict__ stats, int f16) {
  __shared__ double sm[32][2];
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  if (tid < 64) sm[tid >> 1][tid & 1] = 0.0;
  __syncthreads();
  grid_dep_sync_light();
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
      for (int u = 0; u < 4; ++u) load8(src + (ns * P + r + u * step) * ld, v[u], f16);
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
      load8(src + (ns * P + r) * ld, v, f16);
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

// (rstd, -mean * rstd) of one group from its fp64 (Sum, Sum^2): three fp64 operations, the rest in fp32
__device__ __forceinline__ float2 gn_group_scale(const double* __restrict__ st, double inv_n, float eps) {
  const double mean = st[0] * inv_n;
  double var = st[1] * inv_n - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = rsqrtf(static_cast<float>(var) + eps);
  return make_float2(rstd, -static_cast<float>(mean) * rstd);
}

// ------------------------------------------------------------------ GroupNorm apply (+SiLU)
__global__ void gn_apply_kernel(const __nv_bfloat16* __restrict__ x0,
                                const __nv_bfloat16* __restrict__ x1, int C0, int C1, int ld0,
                                int ld1, long long P, int cpg, int rows_per_cta,
                                const double* __restrict__ stats, const float* __restrict__ gamma,
                                const float* __restrict__ beta, float eps, int do_silu,
                                __nv_bfloat16* __restrict__ out, int ldo, int f16) {
  grid_dep_sync_light();
  const int c = threadIdx.x * 8;
  const long long ns = blockIdx.y;
  // mean / rstd of the 32 groups once per CTA (fp64 only for the cancellation in E[x^2] - mean^2; fp64 division and
  // square root in every thread used to be most of this kernel's fixed cost)
  __shared__ float2 s_mr[32];   // (rstd, -mean * rstd)
  {
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    if (tid < 32) s_mr[tid] = gn_group_scale(stats + (ns * 32 + tid) * 2, 1.0 / (static_cast<double>(P) * cpg), eps);
    __syncthreads();
  }
  float a[8], b[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float2 mr = s_mr[(c + k) / cpg];
    const float ga = gamma[c + k];
    a[k] = mr.x * ga;
    b[k] = fmaf(mr.y, ga, beta[c + k]);
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
    for (int u = 0; u < 4; ++u) load8(src + (ns * P + r + u * step) * ld, v[u], f16);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int k = 0; k < 8; k += 2) {   // same arithmetic as the streamed kernel (fma2 + one-MUFU SiLU): identical bits
        const uint64_t y2 = fma2(pack2(v[u][k], v[u][k + 1]), pack2(a[k], a[k + 1]), pack2(b[k], b[k + 1]));
        const float2 o = unpack2(do_silu ? silu2(y2) : y2);
        v[u][k] = o.x;
        v[u][k + 1] = o.y;
      }
      store8(out + (ns * P + r + u * step) * ldo + c, v[u], f16);
    }
  }
  for (; r < r1; r += step) {
    float v[8];
    load8(src + (ns * P + r) * ld, v, f16);
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
      const uint64_t y2 = fma2(pack2(v[k], v[k + 1]), pack2(a[k], a[k + 1]), pack2(b[k], b[k + 1]));
      const float2 o = unpack2(do_silu ? silu2(y2) : y2);
      v[k] = o.x;
      v[k + 1] = o.y;
    }
    store8(out + (ns * P + r) * ldo + c, v, f16);
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
                 const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ y, int ldy, int f16) {
  grid_dep_sync_light();
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
    load8(x + rr * ldx + oc * 8, v[o], f16);
    if (av) {
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(av + oc * 8));
      const float4 a1 = __ldg(reinterpret_cast<const float4*>(av + oc * 8 + 4));
      v[o][0] += a0.x; v[o][1] += a0.y; v[o][2] += a0.z; v[o][3] += a0.w;
      v[o][4] += a1.x; v[o][5] += a1.y; v[o][6] += a1.z; v[o][7] += a1.w;
      if (x_out) {
        if (active) store8(x_out + rr * ldxo + oc * 8, v[o], f16);
        // keep the rounded value: it is what every later consumer of the stream reads
#pragma unroll
        for (int k = 0; k < 8; ++k) v[o][k] = round16(v[o][k], f16);
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
    store8(y + rr * ldy + oc * 8, o8, f16);
  }
}


// ------------------------------------------------------------------ LayerNorm, bulk-copy streamed
// Same math and lane mapping as layernorm_kernel, but rows travel through shared memory: one producer
// thread keeps LN_ST tiles of TILE_ROWS dense rows in flight with 1-D bulk copies (cp.async.bulk, mbarrier
// completion), eight consumer warps normalise a tile IN PLACE in shared memory, and the same producer
// thread writes it back with a bulk store.  ~60-130 KB of loads/stores are in flight per SM regardless of
// register pressure -- the register-resident kernel above peaks near 2 TB/s on B200 (3 CTAs/SM, every warp
// alternating load -> reduce -> store), this one streams.  Needs dense rows (ld == C).
constexpr int LN_ST = 4;        // tiles in flight per CTA
constexpr int LN_CW = 8;        // consumer warps

template <int LPR, int OPL, bool XOUT>
struct LnStream {
  static constexpr int RPW = 32 / LPR;
  static constexpr int TILE_ROWS = LN_CW * RPW;
  static constexpr int C = LPR * OPL * 8;
  static constexpr int TILE_BYTES = TILE_ROWS * C * 2;
  static constexpr int STAGE_BYTES = TILE_BYTES * (XOUT ? 2 : 1);
  static constexpr int SMEM = LN_ST * STAGE_BYTES + 2 * LN_ST * 8 + 128;
};

__device__ __forceinline__ void load8s(const uint8_t* p, float (&v)[8], bool f16) {
  const uint4 q = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 f = unpack16x2(w[t], f16);
    v[2 * t] = f.x;
    v[2 * t + 1] = f.y;
  }
}

template <int LPR, int OPL, bool XOUT>
__global__ void __launch_bounds__(32 * (LN_CW + 1))
layernorm_stream_kernel(const __nv_bfloat16* __restrict__ x, long long M, const float* __restrict__ addv,
                        int rows_per_group, __nv_bfloat16* __restrict__ x_out, const float* __restrict__ gamma,
                        const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ y, int f16) {
  using L = LnStream<LPR, OPL, XOUT>;
  constexpr int C = L::C;
  extern __shared__ uint8_t ln_smem_raw[];
  uint8_t* smem = ln_smem_raw + ((128u - (smem_u32(ln_smem_raw) & 127u)) & 127u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + LN_ST * L::STAGE_BYTES);
  uint64_t* done = full + LN_ST;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long n_tiles = (M + L::TILE_ROWS - 1) / L::TILE_ROWS;
  // contiguous tile range per CTA
  const long long t_begin = n_tiles * blockIdx.x / gridDim.x, t_end = n_tiles * (blockIdx.x + 1) / gridDim.x;
  const int my_tiles = static_cast<int>(t_end - t_begin);
  if (threadIdx.x == 0) {
    for (int i = 0; i < LN_ST; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&done[i], LN_CW);
    }
    fence_barrier_init();
  }
  __syncthreads();
  grid_dep_sync_light();
  auto tile_bytes = [&](long long tile) -> uint32_t {
    long long rows = M - tile * L::TILE_ROWS;
    if (rows > L::TILE_ROWS) rows = L::TILE_ROWS;
    return static_cast<uint32_t>(rows * C * 2);
  };
  if (warp == LN_CW) {
    // ------------------------------------------------------------ producer / store issuer (one thread)
    if (lane == 0) {
      auto issue_load = [&](int i) {
        const long long tile = t_begin + i;
        const int s = i % LN_ST;
        const uint32_t bytes = tile_bytes(tile);
        mbar_expect_tx(&full[s], bytes);
        bulk_load(smem + s * L::STAGE_BYTES, x + tile * L::TILE_ROWS * C, bytes, &full[s]);
      };
      for (int i = 0; i < LN_ST && i < my_tiles; ++i) issue_load(i);
      for (int i = 0; i < my_tiles; ++i) {
        const int s = i % LN_ST;
        const long long tile = t_begin + i;
        mbar_wait(&done[s], (i / LN_ST) & 1);  // consumers wrote the normalised tile (and fenced it)
        const uint32_t bytes = tile_bytes(tile);
        bulk_store(y + tile * L::TILE_ROWS * C, smem + s * L::STAGE_BYTES, bytes);
        if (XOUT) bulk_store(x_out + tile * L::TILE_ROWS * C, smem + s * L::STAGE_BYTES + L::TILE_BYTES, bytes);
        tma_store_commit();
        // the slot stored one iteration ago is reusable once that store has read it
        if (i >= 1 && i - 1 + LN_ST < my_tiles) {
          tma_store_wait_read<1>();
          issue_load(i - 1 + LN_ST);
        }
      }
      tma_store_wait_read<0>();   // shared memory may go once the stores have read it; the writes land by grid end
    }
  } else {
    // ------------------------------------------------------------ consumers
    // packed fp32x2 arithmetic throughout (add2 / fma2), gamma / beta / addv loaded as 64-bit pairs, and the 16-bit format
    // hoisted out of the loop (one instantiation of the body per format): the first packed version still executed 620
    // instructions per warp and tile, a third of them moves between scalar and paired registers and the conversions of BOTH
    // formats
    auto consume = [&](auto fmt) {
      constexpr bool H16 = decltype(fmt)::value;
      constexpr int RPW = L::RPW;
      const int sub = lane / LPR, l = lane % LPR;
      // group (addv row) of the current tile's row, tracked without a 64-bit division per tile
      long long grp = 0, grp_row = 0;   // grp_row = first row of group `grp`
      if (addv) {
        const long long row0 = t_begin * L::TILE_ROWS + warp * RPW + sub;
        grp = row0 / rows_per_group;
        grp_row = grp * rows_per_group;
      }
      for (int i = 0; i < my_tiles; ++i) {
        const int s = i % LN_ST;
        const long long row = (t_begin + i) * L::TILE_ROWS + warp * RPW + sub;
        const bool active = row < M;
        uint8_t* base = smem + s * L::STAGE_BYTES + (warp * RPW + sub) * (C * 2);
        const float* av = nullptr;
        if (addv && active) {
          while (row >= grp_row + rows_per_group) {
            ++grp;
            grp_row += rows_per_group;
          }
          av = addv + grp * C;
        }
        mbar_wait(&full[s], (i / LN_ST) & 1);
        uint64_t v[OPL][4];
        uint64_t sum2 = 0ull;   // (0.f, 0.f)
#pragma unroll
        for (int o = 0; o < OPL; ++o) {
          const int oc = l + LPR * o;
          if (active) {
            const uint4 q = *reinterpret_cast<const uint4*>(base + oc * 16);
            v[o][0] = h16x2_to_f32x2(q.x, H16);
            v[o][1] = h16x2_to_f32x2(q.y, H16);
            v[o][2] = h16x2_to_f32x2(q.z, H16);
            v[o][3] = h16x2_to_f32x2(q.w, H16);
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) v[o][k] = 0ull;
          }
          if (av) {
            const ulonglong2 a0 = __ldg(reinterpret_cast<const ulonglong2*>(av + oc * 8));
            const ulonglong2 a1 = __ldg(reinterpret_cast<const ulonglong2*>(av + oc * 8 + 4));
            v[o][0] = add2(v[o][0], a0.x);
            v[o][1] = add2(v[o][1], a0.y);
            v[o][2] = add2(v[o][2], a1.x);
            v[o][3] = add2(v[o][3], a1.y);
            if (XOUT) {
              // store x + addv and keep the ROUNDED value: it is what every later consumer of the stream reads
              uint32_t w[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                w[k] = pack16x2_f2(v[o][k], H16);
                v[o][k] = h16x2_to_f32x2(w[k], H16);
              }
              *reinterpret_cast<uint4*>(base + L::TILE_BYTES + oc * 16) = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) sum2 = add2(sum2, v[o][k]);
        }
        const float2 sf = unpack2(sum2);
        float sum = sf.x + sf.y;
#pragma unroll
        for (int d = LPR / 2; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
        const float mean = sum / static_cast<float>(C);
        const uint64_t nmean2 = pack2(-mean, -mean);
        uint64_t var2 = 0ull;
#pragma unroll
        for (int o = 0; o < OPL; ++o) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t d2 = add2(v[o][k], nmean2);
            var2 = fma2(d2, d2, var2);
            v[o][k] = d2;   // keep the centred value: the normalise step is one multiply
          }
        }
        const float2 vf = unpack2(var2);
        float var = vf.x + vf.y;
#pragma unroll
        for (int d = LPR / 2; d > 0; d >>= 1) var += __shfl_xor_sync(0xffffffffu, var, d);
        const float rstd = rsqrtf(var / static_cast<float>(C) + eps);
        const uint64_t rstd2 = pack2(rstd, rstd);
        if (active) {
#pragma unroll
          for (int o = 0; o < OPL; ++o) {
            const int oc = l + LPR * o;
            const ulonglong2 g0 = __ldg(reinterpret_cast<const ulonglong2*>(gamma + oc * 8));
            const ulonglong2 g1 = __ldg(reinterpret_cast<const ulonglong2*>(gamma + oc * 8 + 4));
            const ulonglong2 b0 = __ldg(reinterpret_cast<const ulonglong2*>(beta + oc * 8));
            const ulonglong2 b1 = __ldg(reinterpret_cast<const ulonglong2*>(beta + oc * 8 + 4));
            const uint32_t w0 = pack16x2_f2(fma2(mul2(v[o][0], rstd2), g0.x, b0.x), H16);
            const uint32_t w1 = pack16x2_f2(fma2(mul2(v[o][1], rstd2), g0.y, b0.y), H16);
            const uint32_t w2 = pack16x2_f2(fma2(mul2(v[o][2], rstd2), g1.x, b1.x), H16);
            const uint32_t w3 = pack16x2_f2(fma2(mul2(v[o][3], rstd2), g1.y, b1.y), H16);
            *reinterpret_cast<uint4*>(base + oc * 16) = make_uint4(w0, w1, w2, w3);
          }
        }
        fence_proxy_async();  // our generic-proxy writes must be visible to the bulk store
        __syncwarp();
        if (lane == 0) mbar_arrive(&done[s]);
      }
    };
    if (f16) consume(std::true_type{});
    else consume(std::false_type{});
  }
}


// ------------------------------------------------------------------ GroupNorm, bulk-copy streamed
// The register-resident kernels above run at ~2.2-2.8 TB/s on B200 (gn_apply: 96 registers -> 2 CTAs per SM;
// ncu: profiles/r1_v4_ncu_memory_bound.txt).  Here one producer thread keeps GN_ST tiles of R dense rows in
// flight with 1-D bulk copies (one per concat source), 256 consumer threads reduce (stats) or normalise
// (apply, in place for a single source) them in shared memory, and the producer writes tiles back with bulk
// stores.  Thread -> channel octet is fixed, so the per-channel scale/shift (apply) and the partial sums
// (stats) stay in registers for all tiles of a statistics group.  512 consumer threads x 2 CTAs per SM: the
// apply pass is bound by its own instruction stream (SiLU = 2 MUFU + packed FMA per element), not by HBM.
constexpr int GN_ST = 4, GN_CW = 16, GN_CT = GN_CW * 32, GN_MAXCOL = 1;

struct GnStreamP {
  const __nv_bfloat16* x0;
  const __nv_bfloat16* x1;
  int C0, C1, R;            // channels per source, rows per tile
  long long P;              // rows per statistics group (multiple of R)
  long long n_tiles;        // ns * P / R
  int cpg;
  double* stats;            // [ns, 32, 2]
  const float* gamma;
  const float* beta;
  float eps;
  int do_silu;
  __nv_bfloat16* out;       // apply only, dense [rows, C0 + C1]
  int f16;                  // 16-bit tensors are IEEE fp16 instead of bf16
  double inv_n;             // 1 / (P * cpg)
};

template <bool APPLY>
__global__ void __launch_bounds__(GN_CT + 32, 2)
gn_stream_kernel(const GnStreamP p) {
  extern __shared__ uint8_t gn_smem_raw[];
  uint8_t* smem = gn_smem_raw + ((128u - (smem_u32(gn_smem_raw) & 127u)) & 127u);
  const int C = p.C0 + p.C1;
  const int in_bytes0 = p.R * p.C0 * 2, in_bytes1 = p.R * p.C1 * 2, in_bytes = in_bytes0 + in_bytes1;
  const bool two_src = p.C1 > 0;
  const int stage_bytes = in_bytes * ((APPLY && two_src) ? 2 : 1);  // concat input cannot be normalised in place
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + GN_ST * stage_bytes);
  uint64_t* done = full + GN_ST;
  double* sm_stats = reinterpret_cast<double*>(done + GN_ST);  // [32][2], stats kernel only
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t_begin = p.n_tiles * blockIdx.x / gridDim.x, t_end = p.n_tiles * (blockIdx.x + 1) / gridDim.x;
  const int my_tiles = static_cast<int>(t_end - t_begin);
  const long long tiles_per_stat = p.P / p.R;
  if (threadIdx.x == 0) {
    for (int i = 0; i < GN_ST; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&done[i], GN_CW);
    }
    fence_barrier_init();
  }
  if (!APPLY && threadIdx.x < 64) sm_stats[threadIdx.x] = 0.0;
  __syncthreads();
  grid_dep_sync_light();

  if (warp == GN_CW) {
    // ------------------------------------------------------------ producer / store issuer
    if (lane == 0) {
      auto issue_load = [&](int i) {
        const long long row0 = (t_begin + i) * p.R;
        uint8_t* dst = smem + (i % GN_ST) * stage_bytes;
        mbar_expect_tx(&full[i % GN_ST], in_bytes);
        bulk_load(dst, p.x0 + row0 * p.C0, in_bytes0, &full[i % GN_ST]);
        if (two_src) bulk_load(dst + in_bytes0, p.x1 + row0 * p.C1, in_bytes1, &full[i % GN_ST]);
      };
      for (int i = 0; i < GN_ST && i < my_tiles; ++i) issue_load(i);
      for (int i = 0; i < my_tiles; ++i) {
        const int s = i % GN_ST;
        mbar_wait(&done[s], (i / GN_ST) & 1);
        if (APPLY) {
          const long long row0 = (t_begin + i) * p.R;
          bulk_store(p.out + row0 * C, smem + s * stage_bytes + (two_src ? in_bytes : 0), in_bytes);
          tma_store_commit();
          if (i >= 1 && i - 1 + GN_ST < my_tiles) {
            tma_store_wait_read<1>();  // the slot stored one iteration ago has been read
            issue_load(i - 1 + GN_ST);
          }
        } else if (i + GN_ST < my_tiles) {
          issue_load(i + GN_ST);
        }
      }
      if (APPLY) tma_store_wait_read<0>();   // shared memory may go once the stores have read it; the writes land by grid end
    }
    return;
  }

  // ---------------------------------------------------------------- consumers
  const int noct = C >> 3, noct0 = p.C0 >> 3;
  const int rl_n = noct >= GN_CT ? 1 : GN_CT / noct;     // row lanes
  const int ncol = (noct + GN_CT - 1) / GN_CT;           // column sets per thread (1 or 2)
  const int t = threadIdx.x;
  const int rl = noct >= GN_CT ? 0 : t / noct;
  const int oc_first = noct >= GN_CT ? t : t % noct;
  const bool lane_active = rl < rl_n;
  float a[GN_MAXCOL][8], b[GN_MAXCOL][8];   // apply: scale / shift; stats: sum / sum of squares
  long long cur_stat = -1;

  auto flush_stats = [&](long long ns) {
    // fold this thread's channel sums into their groups, CTA-reduce in fp64 shared atomics, then one
    // global fp64 atomic per (group, moment)
#pragma unroll
    for (int cs = 0; cs < GN_MAXCOL; ++cs) {
      const int oc = oc_first + cs * GN_CT;
      if (cs < ncol && oc < noct && lane_active) {
        const int c = oc * 8;
        int g_prev = c / p.cpg;
        double as = 0.0, aq = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int g = (c + k) / p.cpg;
          if (g != g_prev) {
            atomicAdd(&sm_stats[g_prev * 2], as);
            atomicAdd(&sm_stats[g_prev * 2 + 1], aq);
            as = aq = 0.0;
            g_prev = g;
          }
          as += static_cast<double>(a[cs][k]);
          aq += static_cast<double>(b[cs][k]);
          a[cs][k] = b[cs][k] = 0.f;
        }
        atomicAdd(&sm_stats[g_prev * 2], as);
        atomicAdd(&sm_stats[g_prev * 2 + 1], aq);
      }
    }
    named_bar_sync(1, GN_CT);
    if (t < 64) {
      atomicAdd(&p.stats[ns * 64 + t], sm_stats[t]);
      sm_stats[t] = 0.0;
    }
    named_bar_sync(1, GN_CT);
  };

  if (!APPLY) {
#pragma unroll
    for (int cs = 0; cs < GN_MAXCOL; ++cs)
#pragma unroll
      for (int k = 0; k < 8; ++k) a[cs][k] = b[cs][k] = 0.f;
  }

  // statistics group of the current tile, tracked incrementally (a 64-bit division per tile was a third of the apply
  // pass's instruction stream)
  long long ns = t_begin / tiles_per_stat, in_stat = t_begin - ns * tiles_per_stat;
  // the 16-bit format and the SiLU switch are hoisted out of the tile loop (one instantiation of the body per combination):
  // tested per element, the conversions of BOTH formats were executed
  auto run_tiles = [&](auto fmt, auto act) {
  constexpr bool H16 = decltype(fmt)::value, SILU = decltype(act)::value;
  for (int i = 0; i < my_tiles; ++i) {
    const int s = i % GN_ST;
    if (i > 0 && ++in_stat == tiles_per_stat) {
      in_stat = 0;
      ++ns;
    }
    if (ns != cur_stat) {
      if (APPLY) {
        // mean / rstd of the 32 groups once per CTA and statistics group (every consumer thread reaches this at the same
        // tile); doing the fp64 division / square root in each thread cost ~10 us per launch
        float2* s_mr = reinterpret_cast<float2*>(sm_stats);   // [32] (rstd, -mean * rstd)
        named_bar_sync(1, GN_CT);   // nobody still reads the previous group's values
        if (t < 32) s_mr[t] = gn_group_scale(p.stats + (ns * 32 + t) * 2, p.inv_n, p.eps);
        named_bar_sync(1, GN_CT);
#pragma unroll
        for (int cs = 0; cs < GN_MAXCOL; ++cs) {
          const int oc = oc_first + cs * GN_CT;
          if (cs < ncol && oc < noct && lane_active) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int c = oc * 8 + k;
              const float2 mr = s_mr[c / p.cpg];
              const float ga = p.gamma[c];
              a[cs][k] = mr.x * ga;
              b[cs][k] = fmaf(mr.y, ga, p.beta[c]);
            }
          }
        }
      } else if (cur_stat >= 0) {
        flush_stats(cur_stat);
      }
      cur_stat = ns;
    }
    uint8_t* slot = smem + s * stage_bytes;
    mbar_wait(&full[s], (i / GN_ST) & 1);
    if (lane_active) {
#pragma unroll
      for (int cs = 0; cs < GN_MAXCOL; ++cs) {
        const int oc = oc_first + cs * GN_CT;
        if (cs < ncol && oc < noct) {
          const bool src0 = oc < noct0;
          const uint8_t* src = src0 ? slot + oc * 16 : slot + in_bytes0 + (oc - noct0) * 16;
          const int src_ld = (src0 ? p.C0 : p.C1) * 2;
          uint8_t* dst = two_src ? slot + in_bytes + oc * 16 : slot + oc * 16;
          for (int r = rl; r < p.R; r += rl_n) {
            float v[8];
            load8s(src + r * src_ld, v, H16);
            if (APPLY) {
              // packed fp32x2 lanes; silu2: one MUFU (tanh) per element
              uint32_t w[4];
#pragma unroll
              for (int k = 0; k < 8; k += 2) {
                const uint64_t y2 = fma2(pack2(v[k], v[k + 1]), pack2(a[cs][k], a[cs][k + 1]),
                                         pack2(b[cs][k], b[cs][k + 1]));
                w[k >> 1] = pack16x2_f2(SILU ? silu2(y2) : y2, H16);
              }
              *reinterpret_cast<uint4*>(dst + r * (C * 2)) = make_uint4(w[0], w[1], w[2], w[3]);
            } else {
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                a[cs][k] += v[k];
                b[cs][k] = fmaf(v[k], v[k], b[cs][k]);
              }
            }
          }
        }
      }
    }
    if (APPLY) fence_proxy_async();  // normalised rows visible to the bulk store
    __syncwarp();
    if (lane == 0) mbar_arrive(&done[s]);
  }
  };
  if (p.f16) {
    if (APPLY && p.do_silu) run_tiles(std::true_type{}, std::true_type{});
    else run_tiles(std::true_type{}, std::false_type{});
  } else {
    if (APPLY && p.do_silu) run_tiles(std::false_type{}, std::true_type{});
    else run_tiles(std::false_type{}, std::false_type{});
  }
  if (!APPLY && cur_stat >= 0) flush_stats(cur_stat);
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
// have_stats: stats_ws already holds the sums (accumulated by the producing GEMM's epilogue) -> apply pass only
static int groupnorm_impl(const void* x0, const void* x1, int32_t c0, int32_t c1, int32_t ld0,
                          int32_t ld1, int64_t rows_per_stat, int32_t ns, const float* gamma,
                          const float* beta, float eps, int32_t do_silu, void* stats_ws,
                          void* out, int32_t ldo, void* stream_, bool have_stats, int32_t f16) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int C = c0 + c1;
  if (!x0 || !out || !stats_ws || !gamma || !beta) return fail(SA_ERR_INVALID, "sa_groupnorm: null pointer");
  if (c0 % 8 || c1 % 8 || (c1 > 0 && !x1)) return fail(SA_ERR_INVALID, "sa_groupnorm: bad channel split");
  dim3 block, grid;
  int rpc;
  if (gn_geometry(C, rows_per_stat, ns, &block, &grid, &rpc))
    return fail(SA_ERR_INVALID, "sa_groupnorm: C=%d unsupported", C);
  const int cpg = C / 32;
  if (!have_stats) SA_CUDA_CHECK(cudaMemsetAsync(stats_ws, 0, sizeof(double) * 64 * ns, stream));
  // streamed path: dense sources / output, enough bytes to fill the pipeline, tile rows dividing the group
  {
    int R = 32;
    while (R > 1 && R * C * 2 > 20 * 1024) R >>= 1;
    const long long total_bytes = static_cast<long long>(rows_per_stat) * ns * C * 2;
    const bool dense = ld0 == c0 && (c1 == 0 || ld1 == c1) && ldo == C;
    static const bool force_reg = getenv("SA_GN_REG") != nullptr;
    if (dense && !force_reg && rows_per_stat % R == 0 && total_bytes >= (8ll << 20) && C / 8 <= GN_MAXCOL * GN_CT &&
        R * C * 2 <= 24 * 1024) {
      GnStreamP sp;
      sp.x0 = static_cast<const __nv_bfloat16*>(x0);
      sp.x1 = static_cast<const __nv_bfloat16*>(x1);
      sp.C0 = c0;
      sp.C1 = c1;
      sp.R = R;
      sp.P = rows_per_stat;
      sp.n_tiles = rows_per_stat / R * ns;
      sp.cpg = cpg;
      sp.stats = static_cast<double*>(stats_ws);
      sp.gamma = gamma;
      sp.beta = beta;
      sp.eps = eps;
      sp.do_silu = do_silu;
      sp.out = static_cast<__nv_bfloat16*>(out);
      sp.f16 = f16;
      sp.inv_n = 1.0 / (static_cast<double>(rows_per_stat) * cpg);
      const int tile = R * C * 2;
      const int smem_stats = GN_ST * tile + 2 * GN_ST * 8 + 64 * 8 + 128;
      const int smem_apply = GN_ST * tile * (c1 > 0 ? 2 : 1) + 2 * GN_ST * 8 + 64 * 8 + 128;
      static DeviceOnce once;
      if (int rc = once.run([&]() -> int {
            SA_CUDA_CHECK(cudaFuncSetAttribute(gn_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            SA_CUDA_CHECK(cudaFuncSetAttribute(gn_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            return SA_OK;
          }))
        return rc;
      const int sms = num_sms();
      long long g1 = static_cast<long long>(sms) * (smem_stats <= 110 * 1024 ? 2 : 1);
      long long g2 = static_cast<long long>(sms) * (smem_apply <= 110 * 1024 ? 2 : 1);
      if (g1 > sp.n_tiles) g1 = sp.n_tiles;
      if (g2 > sp.n_tiles) g2 = sp.n_tiles;
      if (!have_stats) {
        SA_CUDA_CHECK(launch_k(gn_stream_kernel<false>, dim3(static_cast<unsigned>(g1)), dim3(GN_CT + 32), smem_stats, stream, 1, false, sp));
      }
      SA_CUDA_CHECK(launch_k(gn_stream_kernel<true>, dim3(static_cast<unsigned>(g2)), dim3(GN_CT + 32), smem_apply, stream, 1, false, sp));
      return SA_OK;
    }
  }
  if (!have_stats) {
    SA_CUDA_CHECK(launch_k(gn_stats_kernel, grid, block, 0, stream, 1, false, static_cast<const __nv_bfloat16*>(x0),
                           static_cast<const __nv_bfloat16*>(x1), c0, c1, ld0, ld1, static_cast<long long>(rows_per_stat), cpg, rpc,
                           static_cast<double*>(stats_ws), f16));
  }
  SA_CUDA_CHECK(launch_k(gn_apply_kernel, grid, block, 0, stream, 1, false, static_cast<const __nv_bfloat16*>(x0),
                         static_cast<const __nv_bfloat16*>(x1), c0, c1, ld0, ld1, static_cast<long long>(rows_per_stat), cpg, rpc,
                         static_cast<const double*>(stats_ws), gamma, beta, eps, do_silu, static_cast<__nv_bfloat16*>(out), ldo,
                         f16));
  return SA_OK;
}

extern "C" int sa_groupnorm(const void* x0, const void* x1, int32_t c0, int32_t c1, int32_t ld0,
                            int32_t ld1, int64_t rows_per_stat, int32_t ns, const float* gamma,
                            const float* beta, float eps, int32_t do_silu, void* stats_ws,
                            void* out, int32_t ldo, int32_t act_f16, void* stream_) {
  return groupnorm_impl(x0, x1, c0, c1, ld0, ld1, rows_per_stat, ns, gamma, beta, eps, do_silu, stats_ws, out, ldo,
                        stream_, false, act_f16);
}

extern "C" int sa_groupnorm_apply(const void* x0, const void* x1, int32_t c0, int32_t c1, int32_t ld0,
                                  int32_t ld1, int64_t rows_per_stat, int32_t ns, const float* gamma,
                                  const float* beta, float eps, int32_t do_silu, const void* stats_ws,
                                  void* out, int32_t ldo, int32_t act_f16, void* stream_) {
  return groupnorm_impl(x0, x1, c0, c1, ld0, ld1, rows_per_stat, ns, gamma, beta, eps, do_silu,
                        const_cast<void*>(stats_ws), out, ldo, stream_, true, act_f16);
}

extern "C" int sa_layernorm(const void* x, int32_t ldx, int64_t rows, int32_t c, const float* addv,
                            int32_t rows_per_group, void* x_out, int32_t ldxo, const float* gamma,
                            const float* beta, float eps, void* y, int32_t ldy, int32_t act_f16, void* stream_) {
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
  const bool dense = ldx == c && ldy == c && (!x_out || ldxo == c) && (x_out == nullptr || addv != nullptr) &&
                     rows * c * 2 >= (8ll << 20) && getenv("SA_LN_REG") == nullptr;
  const int n_sm = num_sms();
#define SA_LN_STREAM(LPR, OPL, XO)                                                                        \
  do {                                                                                                    \
    using L = LnStream<LPR, OPL, XO>;                                                                     \
    static DeviceOnce once;                                                                               \
    if (int rc = once.run([&]() -> int {                                                                  \
          SA_CUDA_CHECK(cudaFuncSetAttribute(layernorm_stream_kernel<LPR, OPL, XO>,                       \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, L::SMEM));      \
          return SA_OK;                                                                                   \
        }))                                                                                               \
      return rc;                                                                                          \
    const long long nt = (rows + L::TILE_ROWS - 1) / L::TILE_ROWS;                                        \
    long long g = static_cast<long long>(n_sm) * (L::SMEM <= 110 * 1024 ? 2 : 1);                         \
    if (g > nt) g = nt;                                                                                   \
    SA_CUDA_CHECK(launch_k(layernorm_stream_kernel<LPR, OPL, XO>, dim3(static_cast<unsigned>(g)), dim3(32 * (LN_CW + 1)), \
                           L::SMEM, stream, 1, false, xp, rows, addv, rpg, xo, gamma, beta, eps, yp, act_f16));  \
  } while (0)
#define SA_LN(LPR, OPL)                                                                               \
  if (dense && LnStream<LPR, OPL, true>::SMEM <= 227 * 1024) {                                        \
    if (xo) SA_LN_STREAM(LPR, OPL, true); else SA_LN_STREAM(LPR, OPL, false);                         \
  } else                                                                                              \
    SA_CUDA_CHECK(launch_k(layernorm_kernel<LPR, OPL>, dim3(grid), dim3(warps * 32), 0, stream, 1, false, xp, ldx, rows, c, addv, rpg, \
                           xo, ldxo, gamma, beta, eps, yp, ldy, act_f16))
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
#undef SA_LN_STREAM
  SA_CUDA_CHECK(cudaGetLastError());
  return SA_OK;
}
