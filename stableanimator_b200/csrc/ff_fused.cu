// Fused transformer feed-forward on tcgen05 / TMEM, sm_100a:
//
//   out = [alpha * blend_x + (1 - alpha) *] ( W2 . GEGLU( W1 . LN(x + addv) + b1 ) + b2 + (x + addv) )
//
// i.e. nn.LayerNorm -> FeedForward(GEGLU -> Linear) -> residual add (-> AlphaBlender) of BasicTransformerBlock /
// TemporalBasicTransformerBlock (animation/modules/unet.py:521-523,553-555; :892-897,:936-940 (ff_in);
// :916-917,:949-953 (ff); GEGLU :859-870; FeedForward :827-857; time_mixer :693-697), for C <= 320 channels.
//
// Why one kernel: at the 64x64 level (M = 131072 rows, C = 320) the unfused chain is LayerNorm (read + write [M, C]) ->
// GEGLU GEMM (writes the [M, 4C] hidden tensor, 335 MB, and is bound by its epilogue's issue slots: K = 320 gives the
// tensor pipe only 1280 cycles of work per 128 x 128 tile against ~1300 issue slots of GELU per sub-partition) ->
// out-projection GEMM (reads the 335 MB back): 58 + 231 + 105 us.  Here the hidden tensor never leaves the SM:
//   * the 128-row activation tile stays in shared memory for all 4C / 64 hidden chunks (LayerNorm applied to it in place
//     once, by the epilogue warps, straight after the TMA load);
//   * per hidden chunk of 64:  S = A . W1_chunk^T  (128 x 128 fp32 in TMEM: 64 value | 64 gate columns)
//                              P = GEGLU(S + b1)   (epilogue warps; bf16, written to shared memory in the UMMA
//                                                   K-major SWIZZLE_128B layout)
//                              O += P . W2_chunk^T (128 x C fp32 accumulator resident in TMEM over all chunks);
//   * the GELU work of chunk i runs under the MMAs of S(i+1) and O(i-1): 1300 issue slots against 1920 tensor cycles,
//     so the kernel is tensor-bound where the unfused GEGLU GEMM was issue-bound;
//   * final epilogue: O + b2 + residual (+ blend) -> bf16 -> TMA store.
// Warp roles (320 threads, persistent, one CTA per SM): warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 epilogue
// (warp w: TMEM lane quadrant w & 3; set (w - 2) >> 2 takes every other hidden chunk, and every other 32-column chunk of
// the final output tile).
// The weight stream is what this kernel moves (every 128-row tile re-reads all of W1 and W2 from L2: 2.4 MB per tile, 2.5 GB
// per launch at M = 131072 -- the first version ran at the L2 -> SM bandwidth, 351 us).  CTAs therefore run as CLUSTER
// PAIRS on the same weight stream: each CTA of a pair TMA-loads half of every weight stage and multicasts it to both, a
// stage is released when the MMAs of BOTH CTAs have read it (tcgen05.commit multicast to both empty barriers), and the two
// CTAs work on neighbouring row tiles -- half the L2 traffic per tile.
// TMEM: S [0, 128)  O [128, 128 + C).   smem: A  KC x 16 KB | weight ring 5 x 20 KB | P 2 x 16 KB (P[1] doubles as the
// TMA-store staging of the final epilogue) | b1, b2 -> 224 KB at C = 320.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace sa {

constexpr int FF_ST = 5;                   // weight ring stages
constexpr int FF_STAGE_BYTES = 20 * 1024;  // W1 stage: [128 rows x 64 k] = 16 KB; W2 stage: [C/2 rows x 64 k] <= 20 KB
constexpr int FF_A_CHUNK = 128 * 128;      // [128 rows][64 ch] bf16
constexpr int FF_P_BYTES = 128 * 128;      // [128 rows][64 hidden] bf16
constexpr int FF_EW = 16;                  // epilogue warps: 4 per TMEM lane quadrant = 2 chunk sets x 2 column halves

// -DSA_FF_TRACE (tools/ff_trace.py builds a private copy of this file with it): CTA 0 records clock64() at the hand-off
// points of its first tile -- [0][c] S(c) issued, [1][c] O(c) issued (MMA thread); [2][c] S(c) seen complete, [3][c] P(c)
// written (lane 0 of epilogue warp 0 of each set) -- into the buffer registered with sa_ff_set_trace().
#ifdef SA_FF_TRACE
__device__ long long* g_ff_trace = nullptr;
#define FF_TRACE(ev, idx) do { if (blockIdx.x == 0 && ti == 0 && g_ff_trace) g_ff_trace[(ev) * 64 + (idx)] = clock64(); } while (0)
#else
#define FF_TRACE(ev, idx) do { } while (0)
#endif

struct FfParams {
  int M, C;
  const __nv_bfloat16* x;   // [M, ldx_in]: residual source (re-read by the final epilogue)
  int ldx_in;
  const float* addv;        // [groups, C] or null
  int rows_per_group;
  const float* gamma;
  const float* beta;
  float eps;
  const float* b1;          // [8C] interleaved like W1 (64 value | 64 gate per 128)
  const float* b2;          // [C]
  const __nv_bfloat16* blend_x;
  int ldbx;
  float alpha;
  int m_tiles;
  int act_f16;   // x / w1 / w2 / blend_x / out are IEEE fp16 instead of bf16
};

template <int KC>
struct FfSmem {
  static constexpr int A_OFF = 0;
  static constexpr int RING_OFF = KC * FF_A_CHUNK;
  static constexpr int P_OFF = RING_OFF + FF_ST * FF_STAGE_BYTES;
  static constexpr int B1_OFF = P_OFF + 2 * FF_P_BYTES;       // fp32 [8C] b1 (interleaved) | fp32 [C] b2
  static constexpr int BAR_OFF = B1_OFF + 9 * KC * 64 * 4;
  static constexpr int NBARS = 2 * FF_ST + 14;
  static constexpr int TOTAL = BAR_OFF + NBARS * 8 + 16 + 1024;
};

template <int KC>
__global__ void __launch_bounds__(64 + 32 * FF_EW, 1)
ff_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmOut, const FfParams p) {
  using L = FfSmem<KC>;
  constexpr int C = KC * 64;
  constexpr int NH = 4 * KC;        // hidden chunks of 64
  constexpr int NHALF = C / 2;      // O is accumulated as two column halves (MMA N <= 256)
  constexpr int W1_BYTES = 128 * 128, W2_BYTES = NHALF * 128;
  constexpr int O_COL = 128;
  constexpr int ODLY = 2;           // O(c) is issued after S(c + ODLY)
  constexpr int OCH = C / 32;       // 32-column chunks of the output tile

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem + L::A_OFF;
  uint8_t* sRing = smem + L::RING_OFF;
  uint8_t* sP = smem + L::P_OFF;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + FF_ST;
  uint64_t* a_full = empty_bar + FF_ST;   // TMA: activation tile landed
  uint64_t* a_ready = a_full + 1;         // epilogue warps: LayerNorm applied in place (256 arrivals)
  uint64_t* a_empty = a_ready + 1;        // MMA: last S MMA of the tile has read A
  uint64_t* s_full = a_empty + 1;         // [2] MMA: S(i) complete, i & 1 selects the epilogue warp set it is for
  uint64_t* s_free = s_full + 2;          // [2] epilogue set: S(i) read out of TMEM (128 arrivals)
  uint64_t* p_full = s_free + 2;          // [2] epilogue set: P(i) written (128 arrivals)
  uint64_t* p_free = p_full + 2;          // [2] MMA: O(i) has read P(i)
  uint64_t* o_full = p_free + 2;          // MMA: all O MMAs of the tile complete
  uint64_t* o_free = o_full + 1;          // epilogue: O read out of TMEM (256 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 1);
  float* sB1 = reinterpret_cast<float*>(smem + L::B1_OFF);
  float* sB2 = sB1 + 8 * KC * 64;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());      // 0 / 1 inside the CTA pair
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int pair_tiles = (p.m_tiles + 1) >> 1;                // both CTAs of a pair run the same number of tiles; a tile
                                                              // index >= m_tiles is a phantom (all-OOB loads and stores)
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmOut);
    for (int s = 0; s < FF_ST; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 2);   // the MMA warps of both CTAs of the pair
    }
    mbar_init(a_full, 1);
    // epilogue-side barriers count WARPS (one elected arrival after __syncwarp), not threads: hundreds of arrivals on
    // one mbarrier serialise on the same shared-memory word
    mbar_init(a_ready, FF_EW);
    mbar_init(a_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], FF_EW / 2);
      mbar_init(&p_full[i], FF_EW / 2);
      mbar_init(&p_free[i], 1);
    }
    mbar_init(o_full, 1);
    mbar_init(o_free, FF_EW);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  for (int i = threadIdx.x; i < 8 * KC * 64; i += blockDim.x) sB1[i] = p.b1[i];
  for (int i = threadIdx.x; i < KC * 64; i += blockDim.x) sB2[i] = p.b2[i];
  tc_fence_before();
  __syncthreads();
  cluster_sync();   // the partner's barriers are initialised before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_dep_sync();   // the set-up above overlapped the previous kernel's tail (programmatic dependent launch)

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int it = 0, ti = 0;
      constexpr int W1_HALF_ROWS = 64, W2_HALF_ROWS = NHALF / 2;
      for (int tp = pair; tp < pair_tiles; tp += n_pairs, ++ti) {
        const int m0 = (2 * tp + rank) * 128;
        // weight stream in the order the MMA warp consumes it: W1(0) | W1(1) W2(0) | ... | W1(NH-1) W2(NH-2) | W2(NH-1);
        // the activation tile is requested after W1(0), so the weights of the next tile are already on their way while
        // the last MMAs of this one still read A
        for (int i = 0; i < NH + ODLY; ++i) {
          if (i < NH) {
            for (int kc = 0; kc < KC; ++kc, ++it) {
              const int s = it % FF_ST;
              mbar_wait_relaxed(&empty_bar[s], ((it / FF_ST) & 1) ^ 1);
              mbar_expect_tx(&full_bar[s], W1_BYTES);   // my half + the partner's
              tma_load_2d_mc(sRing + s * FF_STAGE_BYTES + rank * (W1_HALF_ROWS * 128), &tmW1, &full_bar[s], kc * 64,
                             i * 128 + rank * W1_HALF_ROWS, 3);
            }
          }
          if (i == 0) {
            mbar_wait_relaxed(a_empty, (ti & 1) ^ 1);   // previous tile's S MMAs are done with A
            mbar_expect_tx(a_full, KC * FF_A_CHUNK);
#pragma unroll
            for (int kc = 0; kc < KC; ++kc) tma_load_2d(sA + kc * FF_A_CHUNK, &tmX, a_full, kc * 64, m0);
          }
          if (i >= ODLY) {
            for (int h = 0; h < 2; ++h, ++it) {
              const int s = it % FF_ST;
              mbar_wait_relaxed(&empty_bar[s], ((it / FF_ST) & 1) ^ 1);
              mbar_expect_tx(&full_bar[s], W2_BYTES);
              tma_load_2d_mc(sRing + s * FF_STAGE_BYTES + rank * (W2_HALF_ROWS * 128), &tmW2, &full_bar[s], (i - ODLY) * 64,
                             h * NHALF + rank * W2_HALF_ROWS, 3);
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t idesc_s = p.act_f16 ? umma_idesc_f16(128, 128, 0, 0) : umma_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_o = p.act_f16 ? umma_idesc_f16(128, NHALF, 0, 0) : umma_idesc_bf16(128, NHALF, 0, 0);
      int it = 0, ti = 0;
      uint32_t ci = 0;   // global chunk counter (phases of s_full / s_free / p_full / p_free)
      for (int tp = pair; tp < pair_tiles; tp += n_pairs, ++ti) {
        mbar_wait(a_ready, ti & 1);
        tc_fence_after();
        for (int i = 0; i < NH + ODLY; ++i) {
          if (i < NH) {
            // S(i): the epilogue set of chunk i-1 must have read S(i-1) out of TMEM
            const uint32_t c = ci + i;
            if (c > 0) {
              mbar_wait(&s_free[(c - 1) & 1], ((c - 1) >> 1) & 1);
              tc_fence_after();
            }
            for (int kc = 0; kc < KC; ++kc, ++it) {
              const int s = it % FF_ST;
              mbar_wait(&full_bar[s], (it / FF_ST) & 1);
              tc_fence_after();
              const uint64_t da = umma_desc_sw128(smem_u32(sA + kc * FF_A_CHUNK));
              const uint64_t db = umma_desc_sw128(smem_u32(sRing + s * FF_STAGE_BYTES));
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16(tmem_base, da + 2 * k, db + 2 * k, idesc_s, (kc | k) != 0 ? 1u : 0u);
              umma_commit_mc(&empty_bar[s], 3);
            }
            umma_commit(&s_full[c & 1]);
            FF_TRACE(0, i);
            if (i == NH - 1) umma_commit(a_empty);   // the activation tile may be overwritten
          }
          if (i >= ODLY) {
            // O += P(i-ODLY) . W2(i-ODLY)^T -- issued ODLY chunks late: by then the GELU epilogue of that chunk is long
            // done, so this wait never holds up the S MMAs queued behind it (the first version issued O(i-1) here and the
            // tensor pipe idled for the epilogue latency of every chunk: tools/ff_trace.py)
            const uint32_t c = ci + i - ODLY;
            const int pb = c & 1;
            mbar_wait(&p_full[pb], (c >> 1) & 1);
            if (i == ODLY) mbar_wait(o_free, (ti & 1) ^ 1);   // previous tile's O drained by the final epilogue
            tc_fence_after();
            const uint64_t dp = umma_desc_sw128(smem_u32(sP + pb * FF_P_BYTES));
            for (int h = 0; h < 2; ++h, ++it) {
              const int s = it % FF_ST;
              mbar_wait(&full_bar[s], (it / FF_ST) & 1);
              tc_fence_after();
              const uint64_t dw = umma_desc_sw128(smem_u32(sRing + s * FF_STAGE_BYTES));
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16(tmem_base + O_COL + h * NHALF, dp + 2 * k, dw + 2 * k, idesc_o, (i > ODLY || k != 0) ? 1u : 0u);
              umma_commit_mc(&empty_bar[s], 3);
            }
            umma_commit(&p_free[pb]);
            FF_TRACE(1, i - ODLY);
          }
        }
        umma_commit(o_full);
        ci += NH;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int ew = warp - 2;
    const int quad = warp & 3;
    const int half = (ew >> 2) & 1;   // chunk set: hidden chunks half, half + 2, ...
    const int ch = ew >> 3;           // column half of a chunk: hidden columns 32 ch .. 32 ch + 31
    const int r = quad * 32 + lane;   // tile row == TMEM lane
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    uint8_t* stage_slot = sP + ew * 2048;   // both P tiles double as the TMA-store staging of the final epilogue
    const int sw64 = (lane >> 1) & 3;
    const bool f16 = p.act_f16 != 0;
    uint32_t ci = 0;
    int ti = 0;
    for (int tp = pair; tp < pair_tiles; tp += n_pairs, ++ti) {
      const int m0 = (2 * tp + rank) * 128;
      // ---- LayerNorm of the activation tile, in place.  Four threads per row inside one warp: warp ew owns rows
      // 8 ew .. 8 ew + 7; thread (row, hh) owns the 16-byte units 2 hh, 2 hh + 1 of every 64-channel chunk
      // (conflict-free under the 128-byte swizzle); the quarters meet with two shuffles.
      mbar_wait(a_full, ti & 1);
      {
        const int lr = ew * 8 + (lane >> 2), hh = lane & 3;
        const long long grow = static_cast<long long>(m0) + lr;
        const bool in = grow < p.M;
        const float* av = (p.addv && in) ? p.addv + (grow / p.rows_per_group) * C : nullptr;
        uint8_t* rowp = sA + lr * 128;
        float sum = 0.f;
#pragma unroll
        for (int kc = 0; kc < KC; ++kc)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int u = 2 * hh + j;
            uint8_t* q = rowp + kc * FF_A_CHUNK + ((u ^ (lr & 7)) << 4);
            uint4 w = *reinterpret_cast<const uint4*>(q);
            uint32_t ww[4] = {w.x, w.y, w.z, w.w};
            if (av) {   // x + addv, rounded to bf16: the value every consumer of the stream sees
              const float4 a0 = __ldg(reinterpret_cast<const float4*>(av + kc * 64 + u * 8));
              const float4 a1 = __ldg(reinterpret_cast<const float4*>(av + kc * 64 + u * 8 + 4));
              const float aa[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = unpack16x2(ww[e], f16);
                ww[e] = pack16x2(f.x + aa[2 * e], f.y + aa[2 * e + 1], f16);
              }
              *reinterpret_cast<uint4*>(q) = make_uint4(ww[0], ww[1], ww[2], ww[3]);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = unpack16x2(ww[e], f16);
              sum += f.x + f.y;
            }
          }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        const float mean = sum / static_cast<float>(C);
        float var = 0.f;
#pragma unroll
        for (int kc = 0; kc < KC; ++kc)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int u = 2 * hh + j;
            const uint4 w = *reinterpret_cast<const uint4*>(rowp + kc * FF_A_CHUNK + ((u ^ (lr & 7)) << 4));
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = unpack16x2(ww[e], f16);
              const float d0 = f.x - mean, d1 = f.y - mean;
              var += d0 * d0 + d1 * d1;
            }
          }
        var += __shfl_xor_sync(0xffffffffu, var, 1);
        var += __shfl_xor_sync(0xffffffffu, var, 2);
        const float rstd = rsqrtf(var / static_cast<float>(C) + p.eps);
#pragma unroll
        for (int kc = 0; kc < KC; ++kc)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int u = 2 * hh + j;
            uint8_t* q = rowp + kc * FF_A_CHUNK + ((u ^ (lr & 7)) << 4);
            const uint4 w = *reinterpret_cast<const uint4*>(q);
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
            const int chn = kc * 64 + u * 8;
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma + chn));
            const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.gamma + chn + 4));
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.beta + chn));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.beta + chn + 4));
            const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = unpack16x2(ww[e], f16);
              o[e] = pack16x2(fmaf((f.x - mean) * rstd, g[2 * e], bb[2 * e]),
                              fmaf((f.y - mean) * rstd, g[2 * e + 1], bb[2 * e + 1]), f16);
            }
            *reinterpret_cast<uint4*>(q) = make_uint4(o[0], o[1], o[2], o[3]);
          }
        fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's operand reads
        __syncwarp();
        if (lane == 0) mbar_arrive(a_ready);
      }

      // ---- hidden chunks: P(i) = GEGLU(S(i) + b1).  The two warp sets (epilogue warps 0-3 / 4-7, one warp per TMEM lane
      // quadrant each) alternate chunks: while set A runs the GELU of chunk i, the tensor core computes S(i+1) for set B,
      // so a set has two chunk periods for its dependent chain (TMEM read -> GELU -> swizzled store -> proxy fence).
      for (int i = half; i < NH; i += 2) {
        const uint32_t c = ci + i;       // c & 1 == half (NH is even)
        mbar_wait(&s_full[half], (c >> 1) & 1);
        tc_fence_after();
        if (ch == 0 && quad == 0 && lane == 0) FF_TRACE(2, i);
        uint32_t u[32], g[32];
        tmem_ld_32x32(lane_addr + 32 * ch, u);          // this warp's 32 value columns
        tmem_ld_32x32(lane_addr + 64 + 32 * ch, g);     // and 32 gate columns
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[half]);   // the tensor core may start S(i+1)
        // P buffer free?  (O(i-2) has read it; the first chunk of a tile reuses the buffer of the previous tile)
        if (c >= 2) mbar_wait(&p_free[half], ((c >> 1) & 1) ^ 1);
        uint8_t* prow = sP + half * FF_P_BYTES + r * 128;
        const float4* bv = reinterpret_cast<const float4*>(sB1 + i * 128 + 32 * ch);
        const float4* bg = reinterpret_cast<const float4*>(sB1 + i * 128 + 64 + 32 * ch);
#pragma unroll
        for (int j = 0; j < 4; ++j) {   // 8 hidden columns = one 16-byte unit of the P row
          const float4 a0 = bv[2 * j], a1 = bv[2 * j + 1], b0 = bg[2 * j], b1 = bg[2 * j + 1];
          const uint64_t g0 = add2(pack2u(g[8 * j], g[8 * j + 1]), pack2(b0.x, b0.y));
          const uint64_t g1 = add2(pack2u(g[8 * j + 2], g[8 * j + 3]), pack2(b0.z, b0.w));
          const uint64_t g2 = add2(pack2u(g[8 * j + 4], g[8 * j + 5]), pack2(b1.x, b1.y));
          const uint64_t g3 = add2(pack2u(g[8 * j + 6], g[8 * j + 7]), pack2(b1.z, b1.w));
          const uint64_t v0 = mul2(add2(pack2u(u[8 * j], u[8 * j + 1]), pack2(a0.x, a0.y)), gelu2(g0));
          const uint64_t v1 = mul2(add2(pack2u(u[8 * j + 2], u[8 * j + 3]), pack2(a0.z, a0.w)), gelu2(g1));
          const uint64_t v2 = mul2(add2(pack2u(u[8 * j + 4], u[8 * j + 5]), pack2(a1.x, a1.y)), gelu2(g2));
          const uint64_t v3 = mul2(add2(pack2u(u[8 * j + 6], u[8 * j + 7]), pack2(a1.z, a1.w)), gelu2(g3));
          *reinterpret_cast<uint4*>(prow + (((4 * ch + j) ^ (r & 7)) << 4)) =
              make_uint4(pack16x2_f2(v0, f16), pack16x2_f2(v1, f16), pack16x2_f2(v2, f16), pack16x2_f2(v3, f16));
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[half]);
        if (ch == 0 && quad == 0 && lane == 0) FF_TRACE(3, i);
      }
      ci += NH;

      // ---- final epilogue: O + b2 + residual (+ blend) -> bf16 -> TMA store
      mbar_wait(o_full, ti & 1);
      tc_fence_after();
      // staging lives in the two P tiles: their last readers, O(NH-2) and O(NH-1), completed with o_full; the next tile
      // writes them only after S(0) of that tile, which waits for every epilogue thread's a_ready arrival
      const long long grow = static_cast<long long>(m0) + r;
      const bool in = grow < p.M;
      const long long grow_c = in ? grow : 0;
      const float* av = p.addv ? p.addv + (grow_c / p.rows_per_group) * C : nullptr;
      const uint64_t al2 = pack2(p.alpha, p.alpha), be2 = pack2(1.0f - p.alpha, 1.0f - p.alpha);
      if ((ew >> 2) >= OCH) {   // narrow tiles: this warp has no output chunk, the accumulator is not its business
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_free);
      }
      for (int oc = ew >> 2; oc < OCH; oc += 4) {
        uint32_t o[32];
        tmem_ld_32x32(lane_addr + O_COL + oc * 32, o);
        tmem_ld_wait();
        if (oc + 4 >= OCH) {   // last O chunk of this warp: release the accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(o_free);
        }
        uint64_t v2[16];
        const float4* b4 = reinterpret_cast<const float4*>(sB2 + oc * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = b4[j];
          v2[2 * j] = add2(pack2u(o[4 * j], o[4 * j + 1]), pack2(b.x, b.y));
          v2[2 * j + 1] = add2(pack2u(o[4 * j + 2], o[4 * j + 3]), pack2(b.z, b.w));
        }
        {   // residual: x (+ addv, rounded to bf16 as the LayerNorm input was)
          const uint4* x4 = reinterpret_cast<const uint4*>(p.x + grow_c * p.ldx_in + oc * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 q = __ldg(x4 + j);
            uint32_t ww[4] = {q.x, q.y, q.z, q.w};
            if (av) {
              const float4 a0 = __ldg(reinterpret_cast<const float4*>(av + oc * 32 + j * 8));
              const float4 a1 = __ldg(reinterpret_cast<const float4*>(av + oc * 32 + j * 8 + 4));
              const float aa[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = unpack16x2(ww[e], f16);
                ww[e] = pack16x2(f.x + aa[2 * e], f.y + aa[2 * e + 1], f16);
              }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) v2[4 * j + e] = add2(v2[4 * j + e], h16x2_to_f32x2(ww[e], f16));
          }
        }
        if (p.blend_x) {
          const uint4* x4 = reinterpret_cast<const uint4*>(p.blend_x + grow_c * p.ldbx + oc * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 q = __ldg(x4 + j);
            v2[4 * j + 0] = fma2(al2, h16x2_to_f32x2(q.x, f16), mul2(be2, v2[4 * j + 0]));
            v2[4 * j + 1] = fma2(al2, h16x2_to_f32x2(q.y, f16), mul2(be2, v2[4 * j + 1]));
            v2[4 * j + 2] = fma2(al2, h16x2_to_f32x2(q.z, f16), mul2(be2, v2[4 * j + 2]));
            v2[4 * j + 3] = fma2(al2, h16x2_to_f32x2(q.w, f16), mul2(be2, v2[4 * j + 3]));
          }
        }
        if (lane == 0) tma_store_wait_read<0>();   // the staging slot's previous store has read it
        __syncwarp();
        uint8_t* ro = stage_slot + lane * 64;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(ro + ((j ^ sw64) << 4)) =
              make_uint4(pack16x2_f2(v2[4 * j], f16), pack16x2_f2(v2[4 * j + 1], f16), pack16x2_f2(v2[4 * j + 2], f16),
                         pack16x2_f2(v2[4 * j + 3], f16));
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmOut, stage_slot, oc * 32, m0 + quad * 32);
          tma_store_commit();
        }
      }
      if (lane == 0) tma_store_wait_read<0>();   // the P tiles are written again by the first chunks of the next tile
      __syncwarp();
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync();   // neither CTA leaves while the partner may still multicast into its shared memory / barriers
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

template <int KC>
static int launch_ff(const CUtensorMap* tx, const CUtensorMap* tw1, const CUtensorMap* tw2, const CUtensorMap* to,
                     const FfParams& p, cudaStream_t stream) {
  using L = FfSmem<KC>;
  static_assert(L::TOTAL <= 232448, "shared memory budget");
  static DeviceOnce once;
  if (int rc = once.run([&]() -> int {
        SA_CUDA_CHECK(cudaFuncSetAttribute(ff_fused_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        return SA_OK;
      }))
    return rc;
  int pairs = num_sms() / 2;
  const int pair_tiles = (p.m_tiles + 1) / 2;
  if (pairs > pair_tiles) pairs = pair_tiles;
  SA_CUDA_CHECK(launch_k(ff_fused_kernel<KC>, dim3(2 * pairs), dim3(64 + 32 * FF_EW), L::TOTAL, stream, 2, true, *tx, *tw1, *tw2, *to, p));
  return SA_OK;
}

}  // namespace sa

#ifdef SA_FF_TRACE
extern "C" int sa_ff_set_trace(void* buf) {
  return cudaMemcpyToSymbol(sa::g_ff_trace, &buf, sizeof(buf)) == cudaSuccess ? 0 : -2;
}
#endif

extern "C" int sa_ff_fused(const sa_ff_args* a, void* stream_) {
  using namespace sa;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!a || !a->x || !a->gamma || !a->beta || !a->w1 || !a->b1 || !a->w2 || !a->b2 || !a->out)
    return fail(SA_ERR_INVALID, "sa_ff_fused: null pointer");
  const int C = a->c;
  if (C != 64 && C != 128 && C != 256 && C != 320)
    return fail(SA_ERR_INVALID, "sa_ff_fused: C=%d unsupported (64, 128, 256 or 320: the 128 x C fp32 accumulator lives in TMEM)", C);
  if (a->m <= 0) return fail(SA_ERR_INVALID, "sa_ff_fused: empty input");
  if (a->ldx < C || a->ldx % 8 || a->ldc < C || a->ldc % 8 || (a->blend_x && (a->ldbx < C || a->ldbx % 8)))
    return fail(SA_ERR_INVALID, "sa_ff_fused: bad leading dimension");
  if (a->addv && a->rows_per_group < 1) return fail(SA_ERR_INVALID, "sa_ff_fused: rows_per_group must be >= 1 with addv");

  FfParams p;
  memset(&p, 0, sizeof(p));
  p.M = static_cast<int>(a->m);
  p.C = C;
  p.x = static_cast<const __nv_bfloat16*>(a->x);
  p.ldx_in = a->ldx;
  p.addv = a->addv;
  p.rows_per_group = a->rows_per_group > 0 ? a->rows_per_group : 1;
  p.gamma = a->gamma;
  p.beta = a->beta;
  p.eps = a->eps;
  p.b1 = a->b1;
  p.b2 = a->b2;
  p.blend_x = static_cast<const __nv_bfloat16*>(a->blend_x);
  p.ldbx = a->ldbx;
  p.alpha = a->alpha;
  p.m_tiles = static_cast<int>((a->m + 127) / 128);
  p.act_f16 = a->act_f16 ? 1 : 0;

  // x [M, C] (box 64 ch x 128 rows), W1 [8C, C] (64 x 64 = half a stage), W2 [C, 4C] (64 x C/4), out [M, C] (32 x 32, SWIZZLE_64B)
  uint64_t xd[2] = {static_cast<uint64_t>(C), static_cast<uint64_t>(a->m)};
  uint64_t xs[1] = {static_cast<uint64_t>(a->ldx) * 2};
  uint32_t xb[2] = {64, 128};
  const CUtensorMap* tx = get_tmap_bf16(a->x, 2, xd, xs, xb);
  uint64_t w1d[2] = {static_cast<uint64_t>(C), static_cast<uint64_t>(8 * C)};
  uint64_t w1s[1] = {static_cast<uint64_t>(C) * 2};
  uint32_t w1b[2] = {64, 64};    // half a stage: the CTA pair loads the other half and both multicast
  const CUtensorMap* tw1 = get_tmap_bf16(a->w1, 2, w1d, w1s, w1b);
  uint64_t w2d[2] = {static_cast<uint64_t>(4 * C), static_cast<uint64_t>(C)};
  uint64_t w2s[1] = {static_cast<uint64_t>(4 * C) * 2};
  uint32_t w2b[2] = {64, static_cast<uint32_t>(C / 4)};
  const CUtensorMap* tw2 = get_tmap_bf16(a->w2, 2, w2d, w2s, w2b);
  uint64_t od[2] = {static_cast<uint64_t>(C), static_cast<uint64_t>(a->m)};
  uint64_t os[1] = {static_cast<uint64_t>(a->ldc) * 2};
  uint32_t ob[2] = {32, 32};
  const CUtensorMap* to = get_tmap_bf16(a->out, 2, od, os, ob, 64);
  if (!tx || !tw1 || !tw2 || !to) return SA_ERR_TMA;
  switch (C) {
    case 64: return launch_ff<1>(tx, tw1, tw2, to, p, stream);
    case 128: return launch_ff<2>(tx, tw1, tw2, to, p, stream);
    case 256: return launch_ff<4>(tx, tw1, tw2, to, p, stream);
    default: return launch_ff<5>(tx, tw1, tw2, to, p, stream);
  }
}
