// Tapped implicit GEMM on tcgen05 / TMEM, operands staged by TMA (SWIZZLE_128B), sm_100a.
//
// One CTA computes one 128 x BN output tile:
//   warp 0 (one elected lane) : TMA producer  -- per K block (64 channels of one tap) one 4-D box
//                                (64, b1, b2, b3) of A at the tap-shifted position (out-of-grid
//                                positions are zero-filled by TMA = conv zero padding) and one
//                                2-D box (64, BN) of W, both landing on the stage's mbarrier;
//   warp 1 (one elected lane) : MMA issuer    -- 4 x tcgen05.mma (K=16 each) per stage into the
//                                fp32 accumulator in TMEM, tcgen05.commit frees the stage;
//   warps 0-3                 : epilogue      -- tcgen05.ld 32 columns at a time (thread = row),
//                                scale / bias / GEGLU / residual / alpha-blend in fp32, bf16 store.
// Two CTAs are co-resident per SM (<= ~110 KB smem, <= 256 TMEM columns each) so one CTA's epilogue
// overlaps the other's main loop.
//
// Replaces: nn.Linear / nn.Conv2d 3x3 & 1x1 / nn.Conv3d (3,1,1) call sites of
// animation/modules/unet.py (see include/sa_b200.h for the per-site citations).
//
// Ablation knobs (compile-time, off in the product build; `nvcc -DSA_X_...` on this file only):
//   SA_X_NOEPI    epilogue warps only release the accumulator  -> main-loop (TMA + MMA) time
//   SA_X_LDONLY   ... after reading it with tcgen05.ld          -> + TMEM read
//   SA_X_NOSTORE  full epilogue without the TMA store           -> cost of the store path
//   SA_X_NOGELU / SA_X_HALFGELU  GEGLU without / with half of the GELUs -> MUFU / issue sensitivity
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace sa {

struct GemmKParams {
  int M, N;              // N = MMA output columns (before GEGLU halving)
  int kc0, kcpt;         // 64-chunks in source 0, chunks per tap (both sources)
  int ntaps;
  int d1, d2, d3;        // position grid
  int b1, b2, b3;        // box (b1*b2*b3 == 128)
  int t1, t2;            // tile counts along dims 1, 2
  int tap_off[27];
  const float* bias;
  const float* row_bias;
  int ldrb;
  const float* row_scale;
  int rows_per_group;
  const __nv_bfloat16* residual;
  int ldr;
  const __nv_bfloat16* blend_x;
  int ldx;
  float alpha;
  int geglu;
  void* out;
  int ldc;
  int out_f32;
  // persistent kernel (v2) only
  int n_tiles_n, total_tiles;
  float out_scale;  // v = (acc*rs + bias + row_bias [geglu]) * out_scale + residual ...
  int f16_from;     // > 0: output columns >= f16_from are stored as fp16 instead of bf16
  int cs;           // coordinate scale of the source grid along dims 1, 2 (2 = stride-2 convolution)
  int act_f16;      // 16-bit operands / outputs are IEEE fp16 instead of bf16
  double* gn_stats; // fused GroupNorm statistics of the output: fp64 [M / rows_per_stat, n_out / cpg, 2] or null
  int gn_groups;    // n_out / cpg
  FastDiv fd_ntn, fd_t1, fd_t2, fd_rpg, fd_gcpg, fd_grps;
};

template <int BN>
struct TmemCols {
  static constexpr int value = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(128)
gemm_tap_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                const __grid_constant__ CUtensorMap tmB, const GemmKParams p) {
  constexpr int A_BYTES = 128 * 128;
  constexpr int B_BYTES = BN * 128;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int TCOLS = TmemCols<BN>::value;

  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operands need 1024-byte aligned tiles
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int n0 = blockIdx.x * BN;
  const int mt = blockIdx.y;
  const int t1i = mt % p.t1;
  const int t2i = (mt / p.t1) % p.t2;
  const int t3i = mt / (p.t1 * p.t2);
  const int q1 = t1i * p.b1, q2 = t2i * p.b2, q3 = t3i * p.b3;  // tile origin in the grid
  const int KB = p.ntaps * p.kcpt;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TCOLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int tap = 0, ch = 0;
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
        uint8_t* sa_ = smem + s * STAGE_BYTES;
        const int o1 = p.tap_off[tap * 3 + 0], o2 = p.tap_off[tap * 3 + 1],
                  o3 = p.tap_off[tap * 3 + 2];
        if (ch < p.kc0)
          tma_load_4d(sa_, &tmA0, &full_bar[s], ch * 64, q1 * p.cs + o1, q2 * p.cs + o2, q3 + o3);
        else
          tma_load_4d(sa_, &tmA1, &full_bar[s], (ch - p.kc0) * 64, q1 * p.cs + o1, q2 * p.cs + o2, q3 + o3);
        tma_load_2d(sa_ + A_BYTES, &tmB, &full_bar[s], kb * 64, n0);
        if (++ch == p.kcpt) {
          ch = 0;
          ++tap;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = p.act_f16 ? umma_idesc_f16(128, BN, 0, 0) : umma_idesc_bf16(128, BN, 0, 0);
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
        const uint64_t da = umma_desc_sw128(a_addr);
        const uint64_t db = umma_desc_sw128(a_addr + A_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit(&empty_bar[s]);  // stage reusable once these MMAs have read it
      }
      umma_commit(acc_bar);  // accumulator complete
    }
    __syncwarp();
  }

  // ------------------------------------------------------------------ epilogue (all 4 warps)
  mbar_wait(acc_bar, 0);
  tc_fence_after();

  const int r = threadIdx.x;  // tile row == TMEM lane
  const int i1 = r % p.b1, i2 = (r / p.b1) % p.b2, i3 = r / (p.b1 * p.b2);
  const int p1 = q1 + i1, p2 = q2 + i2, p3 = q3 + i3;
  const bool valid = (p1 < p.d1) && (p2 < p.d2) && (p3 < p.d3);
  const long long row = (static_cast<long long>(p3) * p.d2 + p2) * p.d1 + p1;
  const int grp = valid ? static_cast<int>(row / p.rows_per_group) : 0;
  const float rs = (p.row_scale && valid) ? p.row_scale[grp] : 1.0f;
  const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  const float alpha = p.alpha, beta = 1.0f - p.alpha;
  const bool f16 = p.act_f16 != 0;

  // output-column geometry: plain -> BN columns at n0; GEGLU -> BN/2 columns at n0/2, MMA columns
  // come in 128-wide groups [64 value | 64 gate]
  const int n_out_total = p.geglu ? p.N / 2 : p.N;
  const int out_cols_tile = p.geglu ? BN / 2 : BN;
  const int nout0 = p.geglu ? n0 / 2 : n0;

  for (int c = 0; c < out_cols_tile; c += 32) {
    if (nout0 + c >= n_out_total) break;  // warp-uniform
    float v[32];
    {
      uint32_t u[32];
      const int mc = p.geglu ? ((c >> 6) * 128 + (c & 63)) : c;  // MMA column of this chunk
      tmem_ld_32x32(lane_addr + mc, u);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(u[j]) * rs;
      if (p.bias) {
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0 + mc);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (n0 + mc + 4 * j < p.N) {
            const float4 b = __ldg(b4 + j);
            v[4 * j + 0] += b.x;
            v[4 * j + 1] += b.y;
            v[4 * j + 2] += b.z;
            v[4 * j + 3] += b.w;
          }
        }
      }
      if (p.geglu) {
        uint32_t g[32];
        tmem_ld_32x32(lane_addr + mc + 64, g);
        tmem_ld_wait();
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0 + mc + 64);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.bias) b = __ldg(b4 + j);
          v[4 * j + 0] *= gelu_erf(__uint_as_float(g[4 * j + 0]) * rs + b.x);
          v[4 * j + 1] *= gelu_erf(__uint_as_float(g[4 * j + 1]) * rs + b.y);
          v[4 * j + 2] *= gelu_erf(__uint_as_float(g[4 * j + 2]) * rs + b.z);
          v[4 * j + 3] *= gelu_erf(__uint_as_float(g[4 * j + 3]) * rs + b.w);
        }
      }
    }
    if (!valid) continue;
    if (p.out_scale != 1.0f) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= p.out_scale;
    }
    const int ncol = nout0 + c;  // first output column of this chunk
    if (p.row_bias) {
      const float4* b4 =
          reinterpret_cast<const float4*>(p.row_bias + static_cast<long long>(grp) * p.ldrb + ncol);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (ncol + 4 * j < n_out_total) {
          const float4 b = __ldg(b4 + j);
          v[4 * j + 0] += b.x;
          v[4 * j + 1] += b.y;
          v[4 * j + 2] += b.z;
          v[4 * j + 3] += b.w;
        }
      }
    }
    if (p.residual) {
      const uint4* r4 = reinterpret_cast<const uint4*>(p.residual + row * p.ldr + ncol);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (ncol + 8 * j < n_out_total) {
          const uint4 q = __ldg(r4 + j);
          const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 f = unpack16x2(w[t], f16);
            v[8 * j + 2 * t] += f.x;
            v[8 * j + 2 * t + 1] += f.y;
          }
        }
      }
    }
    if (p.blend_x) {
      const uint4* x4 = reinterpret_cast<const uint4*>(p.blend_x + row * p.ldx + ncol);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (ncol + 8 * j < n_out_total) {
          const uint4 q = __ldg(x4 + j);
          const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 f = unpack16x2(w[t], f16);
            v[8 * j + 2 * t] = alpha * f.x + beta * v[8 * j + 2 * t];
            v[8 * j + 2 * t + 1] = alpha * f.y + beta * v[8 * j + 2 * t + 1];
          }
        }
      }
    }
    if (p.out_f32) {
      float4* o4 = reinterpret_cast<float4*>(static_cast<float*>(p.out) + row * p.ldc + ncol);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (ncol + 4 * j < n_out_total)
          o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
      uint4* o4 =
          reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + row * p.ldc + ncol);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (ncol + 8 * j < n_out_total)
          o4[j] = make_uint4(pack16x2(v[8 * j], v[8 * j + 1], f16), pack16x2(v[8 * j + 2], v[8 * j + 3], f16),
                             pack16x2(v[8 * j + 4], v[8 * j + 5], f16), pack16x2(v[8 * j + 6], v[8 * j + 7], f16));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TCOLS>(tmem_base);
}

// =====================================================================================================
// v2: persistent, warp-specialised variant (bf16 output).  One CTA per SM loops over output tiles:
//   warp 0      : TMA producer (A boxes + W boxes into a STAGES-deep ring that runs across tiles)
//   warp 1      : MMA issuer; accumulators double-buffered in TMEM (2 x BN columns) so the epilogue
//                 of tile i overlaps the main loop of tile i+1
//   warps 2..9  : epilogue.  Warp w owns the 32 tile rows of TMEM quadrant (w % 4) and every second
//                 32-column chunk ((w - 2) / 4 selects even / odd chunks); the eight warps run fully
//                 decoupled pipelines (no CTA-level barrier).  Per chunk: tcgen05.ld (prefetched one
//                 chunk ahead) -> fp32 math -> bf16 into the warp's SWIZZLE_64B staging tile
//                 [32 rows x 64 B] -> one TMA store of the (32 cols, s1, s2, s3) sub-box, which also
//                 clips ragged tiles.  residual / blend_x sub-boxes are TMA-loaded two chunks ahead
//                 into the same staging slots (a thread reads, then overwrites, only its own row), so
//                 every global access of the epilogue is a bulk copy.
// =====================================================================================================
constexpr int EPI_CHUNK = 32;             // output columns per chunk
constexpr int EPI_TILE_BYTES = 32 * 64;   // per-warp staging tile: [32 rows][32 bf16]
// Staging slots per epilogue warp: 3 output / residual slots (the residual sub-box is prefetched two chunks ahead) and
// 2 blend-operand slots.  SLIM (the GEGLU instantiations, which never carry a residual or a blend operand): 2 output
// slots only -- the 48 KB this frees with eight epilogue warps buy two more pipeline stages, which the short-K GEGLU
// GEMMs need to cover the TMA latency (4 stages x 32 KB in flight against ~125 B/clk of operand traffic per SM).
template <bool SLIM> struct EpiSlots { static constexpr int RO = SLIM ? 2 : 3, X = SLIM ? 0 : 2; };
// Accumulator sets of the fused GroupNorm statistics, indexed by (tile counter % GN_SETS).  The TMEM double buffer
// bounds the skew between epilogue warps to two tiles (a warp cannot read tile i+2 before every warp has released
// tile i, and cannot release tile i+1 before it is done with tile i), so three sets never alias.
constexpr int GN_SETS = 3;

template <int BN, int STAGES, int EW, bool SLIM = false, bool PAIR = false>
struct Gemm2Smem {
  static constexpr int EPI_WARP_BYTES = (EpiSlots<SLIM>::RO + EpiSlots<SLIM>::X) * EPI_TILE_BYTES;
  static constexpr int STAGE_BYTES = 128 * 128 + (PAIR ? BN / 2 : BN) * 128;   // PAIR: each CTA holds half of the weight tile
  static constexpr int EPI_OFF = STAGES * STAGE_BYTES;
  static constexpr int BAR_OFF = EPI_OFF + EW * EPI_WARP_BYTES;
  static constexpr int NBARS = 2 * STAGES + 4 + EW * (EpiSlots<SLIM>::RO + EpiSlots<SLIM>::X);
  // fused GroupNorm statistics (non-SLIM only): GN_SETS per-tile accumulator sets of [32 groups][2] fp64 + arrival counters
  static constexpr int GN_OFF = BAR_OFF + NBARS * 8 + 16;
  static constexpr int GN_BYTES = SLIM ? 0 : (GN_SETS * 64 * 8 + 16);
  static constexpr int TOTAL = GN_OFF + GN_BYTES + 1024;  // + alignment slack
};

// EW = number of epilogue warps: 8 (two per TMEM quadrant, alternating chunks) for short-K problems
// whose epilogue must keep up with a short main loop, 4 (less staging smem -> one more pipeline stage)
// for long-K problems.
// PLAIN: the epilogue is bias (+ GEGLU) only -- no row scale / row bias / output scale / residual / blend /
// fp16 block.  The short-K GEMMs are bound by the sub-partition issue port, so the instantiation without the
// per-chunk feature tests and the staging-slot machinery is worth having for the launches that need nothing else.
// PAIR: CTA pairs (cluster of two, tcgen05 cta_group::2).  The two CTAs take neighbouring row tiles of the same column
// tile and run as ONE 256-row MMA issued by the pair's leader; the weight operand of such an MMA is split between the two
// CTAs' shared memories (BN / 2 rows each), so a CTA streams half of the weight bytes from L2 (the long-K convolutions
// re-read the weights once per row tile: 1024x at the 64x64 level, L2 throughput 69 % of peak) and a stage shrinks from
// 16 + BN / 8 KB to 16 + BN / 16 KB -- two more pipeline stages in the same shared memory.  Protocol: each CTA's producer
// loads its A box and its half of the W box into its own ring slot with the bytes completing on the LEADER's full
// barrier; the leader's commits (stage free, accumulator full) are multicast to both CTAs; the epilogue warps of both
// CTAs release the accumulator on the leader's barrier.  Everything behind the accumulator is per CTA and unchanged.
template <int BN, int STAGES, bool GEGLU, int EW, bool PLAIN, bool PAIR = false>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
gemm_tap_persist_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                        const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
                        const __grid_constant__ CUtensorMap tmRes, const __grid_constant__ CUtensorMap tmBld,
                        const GemmKParams p) {
  constexpr bool SLIM = GEGLU || PLAIN;   // neither ever carries a residual / blend operand
  using L = Gemm2Smem<BN, STAGES, EW, SLIM, PAIR>;
  constexpr int EPI_RO_SLOTS = EpiSlots<SLIM>::RO, EPI_X_SLOTS = EpiSlots<SLIM>::X;
  constexpr int EPI_WARP_BYTES = L::EPI_WARP_BYTES;
  constexpr int CSTRIDE = EW / 4;  // chunk stride of one epilogue warp
  constexpr int A_BYTES = 128 * 128;
  constexpr int STAGE_BYTES = L::STAGE_BYTES;
  constexpr int TCOLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  constexpr int ACC_STRIDE = TCOLS / 2;  // column offset of the second accumulator

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_full = empty_bar + STAGES;   // [2]
  uint64_t* acc_empty = acc_full + 2;        // [2]
  uint64_t* epi_bar = acc_empty + 2;         // [EPI_WARPS][RO_SLOTS + X_SLOTS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(epi_bar + EW * (EPI_RO_SLOTS + EPI_X_SLOTS));
  double* gn_acc = reinterpret_cast<double*>(smem + L::GN_OFF);                    // [GN_SETS][32][2]
  uint32_t* gn_cnt = reinterpret_cast<uint32_t*>(smem + L::GN_OFF + GN_SETS * 64 * 8);  // [GN_SETS]

  const int warp = threadIdx.x >> 5;
  const int KB = p.ntaps * p.kcpt;
  const int out_cols_tile = GEGLU ? BN / 2 : BN;
  const int CPT = out_cols_tile / EPI_CHUNK;  // chunks per tile
  // tile walk: a "work item" is one output tile (PAIR: one pair of row-neighbouring tiles); items first, first + stride, ...
  const int rank = PAIR ? static_cast<int>(cluster_ctarank()) : 0;
  const int item_first = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int item_stride = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int item_count = p.total_tiles;   // PAIR: number of tile pairs

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], EW * (PAIR ? 2 : 1));   // one elected arrival per epilogue warp (PAIR: of both CTAs, on the leader)
    }
    for (int i = 0; i < EW * (EPI_RO_SLOTS + EPI_X_SLOTS); ++i) mbar_init(&epi_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) tmem_alloc_2cta<TCOLS>(tmem_slot);
    else tmem_alloc<TCOLS>(tmem_slot);
  }
  if constexpr (!SLIM) {
    if (p.gn_stats) {
      for (int i = threadIdx.x; i < GN_SETS * 64; i += blockDim.x) gn_acc[i] = 0.0;
      if (threadIdx.x < GN_SETS) gn_cnt[threadIdx.x] = 0u;
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync();   // the partner's barriers exist before anything is signalled on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_dep_sync();   // everything above overlapped the previous kernel's tail (programmatic dependent launch)

  auto tile_origin = [&](int tile, int& n0, int& q1, int& q2, int& q3) {
    uint32_t mt, nt, a, b, c;
    p.fd_ntn.divmod(static_cast<uint32_t>(tile), mt, nt);
    if constexpr (PAIR) mt = 2 * mt + rank;   // a row tile past the last one is a phantom: all-OOB loads and stores
    p.fd_t1.divmod(mt, a, b);   // b = mt % t1
    p.fd_t2.divmod(a, c, a);    // a = (mt / t1) % t2, c = mt / (t1*t2)
    n0 = static_cast<int>(nt) * BN;
    q1 = static_cast<int>(b) * p.b1;
    q2 = static_cast<int>(a) * p.b2;
    q3 = static_cast<int>(c) * p.b3;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int it = 0;
      constexpr int B_ROWS = PAIR ? BN / 2 : BN;
      for (int tile = item_first; tile < item_count; tile += item_stride) {
        int n0, q1, q2, q3;
        tile_origin(tile, n0, q1, q2, q3);
        int tap = 0, ch = 0;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait_relaxed(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
          uint8_t* sa_ = smem + s * STAGE_BYTES;
          const int o1 = p.tap_off[tap * 3 + 0], o2 = p.tap_off[tap * 3 + 1], o3 = p.tap_off[tap * 3 + 2];
          if constexpr (PAIR) {
            // both CTAs' boxes complete on the leader's barrier (it expects the bytes of the whole pair)
            if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * STAGE_BYTES);
            const uint32_t fb = mapa_u32(&full_bar[s], 0);
            if (ch < p.kc0)
              tma_load_4d_2cta(sa_, &tmA0, fb, ch * 64, q1 * p.cs + o1, q2 * p.cs + o2, q3 + o3);
            else
              tma_load_4d_2cta(sa_, &tmA1, fb, (ch - p.kc0) * 64, q1 * p.cs + o1, q2 * p.cs + o2, q3 + o3);
            tma_load_2d_2cta(sa_ + A_BYTES, &tmB, fb, kb * 64, n0 + rank * B_ROWS);
          } else {
            mbar_expect_tx(&full_bar[s], STAGE_BYTES);
            if (ch < p.kc0)
              tma_load_4d(sa_, &tmA0, &full_bar[s], ch * 64, q1 * p.cs + o1, q2 * p.cs + o2, q3 + o3);
            else
              tma_load_4d(sa_, &tmA1, &full_bar[s], (ch - p.kc0) * 64, q1 * p.cs + o1, q2 * p.cs + o2, q3 + o3);
            tma_load_2d(sa_ + A_BYTES, &tmB, &full_bar[s], kb * 64, n0);
          }
          if (++ch == p.kcpt) {
            ch = 0;
            ++tap;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (PAIR: the leader CTA only)
    if (rank == 0 && elect_one()) {
      const uint32_t idesc = p.act_f16 ? umma_idesc_f16(PAIR ? 256 : 128, BN, 0, 0) : umma_idesc_bf16(PAIR ? 256 : 128, BN, 0, 0);
      int it = 0, ti = 0;
      for (int tile = item_first; tile < item_count; tile += item_stride, ++ti) {
        const int ab = ti & 1;
        // epilogue drained this accumulator (PAIR: in both CTAs)
        if constexpr (PAIR) mbar_wait_cluster(&acc_empty[ab], ((ti >> 1) & 1) ^ 1);
        else mbar_wait(&acc_empty[ab], ((ti >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t acc = tmem_base + ab * ACC_STRIDE;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % STAGES;
          if constexpr (PAIR) mbar_wait_cluster(&full_bar[s], (it / STAGES) & 1);
          else mbar_wait(&full_bar[s], (it / STAGES) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
          const uint64_t da = umma_desc_sw128(a_addr);
          const uint64_t db = umma_desc_sw128(a_addr + A_BYTES);
          if constexpr (PAIR) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_2cta(acc, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_2cta(&empty_bar[s], 3);
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(acc, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit(&empty_bar[s]);
          }
        }
        if constexpr (PAIR) umma_commit_2cta(&acc_full[ab], 3);
        else umma_commit(&acc_full[ab]);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (EW decoupled warps)
    const int ew = warp - 2;              // 0..EW-1
    const int quad = warp & 3;            // TMEM lane quadrant this warp may access
    const int half = ew >> 2;             // first chunk of this warp (chunks half, half+CSTRIDE, ...)
    const int lane = threadIdx.x & 31;
    const int r = quad * 32 + lane;       // tile row == TMEM lane
    const int i1 = r % p.b1, i2 = (r / p.b1) % p.b2, i3 = r / (p.b1 * p.b2);
    // origin of this warp's 32-row sub-box inside the tile box (rows quad*32 .. quad*32+31)
    const int r0 = quad * 32;
    const int w1 = r0 % p.b1, w2 = (r0 / p.b1) % p.b2, w3 = r0 / (p.b1 * p.b2);
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const float alpha = p.alpha, beta = 1.0f - p.alpha, gs = p.out_scale;
    const int my_tiles = (item_count - item_first + item_stride - 1) / item_stride;
    const uint32_t acc_empty_l0 = PAIR ? mapa_u32(&acc_empty[0], 0) : 0u, acc_empty_l1 = PAIR ? mapa_u32(&acc_empty[1], 0) : 0u;
    auto release_acc = [&](int ab) {   // this warp has read its part of accumulator ab (tcgen05.wait::ld is warp-collective)
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR) mbar_arrive_cluster(ab ? acc_empty_l1 : acc_empty_l0);
        else mbar_arrive(&acc_empty[ab]);
      }
    };
    const int cpw = (CPT - half + CSTRIDE - 1) / CSTRIDE;  // chunks of a tile handled by this warp
    const int total_chunks = my_tiles * cpw;
    uint8_t* wbase = smem + L::EPI_OFF + ew * EPI_WARP_BYTES;
    uint8_t* sRO = wbase;
    uint8_t* sX = wbase + EPI_RO_SLOTS * EPI_TILE_BYTES;
    uint64_t* ro_ready = epi_bar + ew * (EPI_RO_SLOTS + EPI_X_SLOTS);
    uint64_t* x_ready = ro_ready + EPI_RO_SLOTS;
    const int sw = (lane >> 1) & 3;  // SWIZZLE_64B: 16-byte chunk index ^= (row >> 1) & 3
    uint8_t* my_ro = sRO + lane * 64;
    const uint8_t* my_x = sX + lane * 64;

    // lane 0 prepares the staging slots of this warp's chunks strictly in order (chunk 0, 1, 2, ...),
    // two chunks ahead of their use: TMA-load the residual / blend sub-boxes, or just mark the slot free
    int nx_g = 0, nx_tile = item_first, nx_c = half;
    int nx_n0 = 0, nx_q1 = 0, nx_q2 = 0, nx_q3 = 0;  // tile coordinates of nx_tile (div/mod chain once per tile)
    bool nx_fresh = true;
    auto issue_next = [&]() {
      if (nx_g >= total_chunks) return;
      if (nx_fresh) {
        tile_origin(nx_tile, nx_n0, nx_q1, nx_q2, nx_q3);
        nx_fresh = false;
      }
      const int col = (GEGLU ? nx_n0 / 2 : nx_n0) + nx_c * EPI_CHUNK;
      const int slot = nx_g % EPI_RO_SLOTS;
      if (p.residual) {
        mbar_expect_tx(&ro_ready[slot], EPI_TILE_BYTES);
        tma_load_4d(sRO + slot * EPI_TILE_BYTES, &tmRes, &ro_ready[slot], col, nx_q1 + w1, nx_q2 + w2, nx_q3 + w3);
      } else {
        mbar_arrive(&ro_ready[slot]);
      }
      if constexpr (EPI_X_SLOTS > 0) {
        if (p.blend_x) {
          const int xs = nx_g % EPI_X_SLOTS;
          mbar_expect_tx(&x_ready[xs], EPI_TILE_BYTES);
          tma_load_4d(sX + xs * EPI_TILE_BYTES, &tmBld, &x_ready[xs], col, nx_q1 + w1, nx_q2 + w2, nx_q3 + w3);
        }
      }
      ++nx_g;
      nx_c += CSTRIDE;
      if (nx_c >= CPT) {
        nx_c = half;
        nx_tile += item_stride;
        nx_fresh = true;
      }
    };
    // Without residual / blend operands nothing is prefetched into the staging slots: a slot is free as soon
    // as the store issued from it two chunks ago has read it (wait_group.read below), so the whole
    // issue_next / ro_ready machinery (two div/mod chains and a barrier round trip per chunk -- the short-K
    // GEMMs are issue-bound) is skipped.
    const bool staged = !SLIM && ((p.residual != nullptr) || (p.blend_x != nullptr));
    if (lane == 0 && staged) {
      issue_next();
      issue_next();
    }
    const bool need_grp = !PLAIN && ((p.row_scale != nullptr) || (p.row_bias != nullptr));

    int gc = 0, ti = 0;
    for (int tile = item_first; tile < item_count; tile += item_stride, ++ti) {
      int n0, q1, q2, q3;
      tile_origin(tile, n0, q1, q2, q3);
      int grp = 0;
      if (need_grp) {
        const int p1 = q1 + i1, p2 = q2 + i2, p3 = q3 + i3;
        if ((p1 < p.d1) && (p2 < p.d2) && (p3 < p.d3))
          grp = static_cast<int>(p.fd_rpg.div(static_cast<uint32_t>((p3 * p.d2 + p2) * p.d1 + p1)));
      }
      const float rs = (!PLAIN && p.row_scale) ? p.row_scale[grp] : 1.0f;
      // fused GroupNorm statistics: statistic index of this thread's row (-1 = row outside the grid); the fast path
      // below needs it uniform over the warp's 32 rows
      int gn_sid = -1;
      bool gn_uniform = false;
      // gn_tile: every row of the tile lies inside the grid and belongs to ONE statistic -> the cheap path: column
      // sums from the staged bf16 tile in packed fp32, warp-reduced to groups, kept in two registers per lane
      // (lane l = group g_first + l) over the tile, then one fp64 shared-memory add per lane and ONE set of fp64
      // global atomics per (tile, group) by the last epilogue warp to finish the tile
      bool gn_tile = false;
      int g_first = 0;
      float ga1 = 0.f, ga2 = 0.f;
      const int nout0 = GEGLU ? n0 / 2 : n0;
      if (!PLAIN && !GEGLU && p.gn_stats) {
        const int p1 = q1 + i1, p2 = q2 + i2, p3 = q3 + i3;
        if ((p1 < p.d1) && (p2 < p.d2) && (p3 < p.d3))
          gn_sid = static_cast<int>(p.fd_grps.div(static_cast<uint32_t>((p3 * p.d2 + p2) * p.d1 + p1)));
        const int e1 = q1 + p.b1 - 1, e2 = q2 + p.b2 - 1, e3 = q3 + p.b3 - 1;
        if (e1 < p.d1 && e2 < p.d2 && e3 < p.d3) {   // rows of the box span [row(q), row(e)]; the statistic index is monotonic
          const uint32_t smin = p.fd_grps.div(static_cast<uint32_t>((q3 * p.d2 + q2) * p.d1 + q1));
          const uint32_t smax = p.fd_grps.div(static_cast<uint32_t>((e3 * p.d2 + e2) * p.d1 + e1));
          gn_tile = smin == smax;
        }
        g_first = static_cast<int>(p.fd_gcpg.div(static_cast<uint32_t>(nout0)));
        if (!gn_tile)
          gn_uniform = __all_sync(0xffffffffu, gn_sid == __shfl_sync(0xffffffffu, gn_sid, 0)) && gn_sid >= 0;
      }
      const int ab = ti & 1;
      mbar_wait_relaxed(&acc_full[ab], (ti >> 1) & 1);
      tc_fence_after();
      const uint32_t acc = lane_addr + ab * ACC_STRIDE;
      auto mma_col = [&](int c) { return GEGLU ? ((c >> 1) * 128 + (c & 1) * 32) : c * EPI_CHUNK; };
#ifdef SA_X_NOEPI
      tc_fence_before();
      release_acc(ab);
      continue;
#endif
#ifdef SA_X_LDONLY
      {  // read the accumulator (as the real epilogue does) and release it, nothing else
        uint32_t uu[32];
        for (int c = half; c < CPT; c += CSTRIDE) {
          tmem_ld_32x32(acc + mma_col(c), uu);
          tmem_ld_wait();
          if (uu[3] == 0x7fc12345u && uu[7] == 0x1234567u) p.row_scale[0];
        }
        tc_fence_before();
        release_acc(ab);
        continue;
      }
#endif

      uint32_t u[32], g[GEGLU ? 32 : 1];
      if (half < CPT) {
        tmem_ld_32x32(acc + mma_col(half), u);
        if constexpr (GEGLU) tmem_ld_32x32(acc + mma_col(half) + 64, reinterpret_cast<uint32_t(&)[32]>(g));
      }
      for (int c = half; c < CPT; c += CSTRIDE, ++gc) {
        const int mc = mma_col(c);
        // all epilogue math runs on packed fp32x2 lanes (FFMA2): half the issue slots of scalar code
        uint64_t v2[16], g2[GEGLU ? 16 : 1];
        tmem_ld_wait();
        const uint64_t rs2 = pack2(rs, rs);
        if (p.bias) {
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0 + mc);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = __ldg(b4 + j);
            if constexpr (PLAIN) {
              v2[2 * j] = add2(pack2u(u[4 * j], u[4 * j + 1]), pack2(b.x, b.y));
              v2[2 * j + 1] = add2(pack2u(u[4 * j + 2], u[4 * j + 3]), pack2(b.z, b.w));
            } else {
              v2[2 * j] = fma2(pack2u(u[4 * j], u[4 * j + 1]), rs2, pack2(b.x, b.y));
              v2[2 * j + 1] = fma2(pack2u(u[4 * j + 2], u[4 * j + 3]), rs2, pack2(b.z, b.w));
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            v2[j] = PLAIN ? pack2u(u[2 * j], u[2 * j + 1]) : mul2(pack2u(u[2 * j], u[2 * j + 1]), rs2);
        }
        if constexpr (GEGLU) {
          if (p.bias) {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0 + mc + 64);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = __ldg(b4 + j);
              if constexpr (PLAIN) {
                g2[2 * j] = add2(pack2u(g[4 * j], g[4 * j + 1]), pack2(b.x, b.y));
                g2[2 * j + 1] = add2(pack2u(g[4 * j + 2], g[4 * j + 3]), pack2(b.z, b.w));
              } else {
                g2[2 * j] = fma2(pack2u(g[4 * j], g[4 * j + 1]), rs2, pack2(b.x, b.y));
                g2[2 * j + 1] = fma2(pack2u(g[4 * j + 2], g[4 * j + 3]), rs2, pack2(b.z, b.w));
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              g2[j] = PLAIN ? pack2u(g[2 * j], g[2 * j + 1]) : mul2(pack2u(g[2 * j], g[2 * j + 1]), rs2);
          }
        }
        if (c + CSTRIDE < CPT) {  // prefetch the next chunk's accumulator columns
          tmem_ld_32x32(acc + mma_col(c + CSTRIDE), u);
          if constexpr (GEGLU) tmem_ld_32x32(acc + mma_col(c + CSTRIDE) + 64, reinterpret_cast<uint32_t(&)[32]>(g));
        } else {            // accumulator fully read by this thread: hand it back to the MMA warp
          tc_fence_before();
          release_acc(ab);
        }
        if constexpr (GEGLU) {
#pragma unroll
#ifdef SA_X_NOGELU
          for (int j = 0; j < 16; ++j) v2[j] = mul2(v2[j], g2[j]);
#elif defined(SA_X_HALFGELU)
          for (int j = 0; j < 16; ++j) v2[j] = mul2(v2[j], (j & 1) ? g2[j] : gelu2(g2[j]));
#else
          for (int j = 0; j < 16; ++j) v2[j] = mul2(v2[j], gelu2(g2[j]));
#endif
        }
        const int ncol = nout0 + c * EPI_CHUNK;
        if (!PLAIN && p.row_bias) {
          const float4* b4 = reinterpret_cast<const float4*>(p.row_bias + static_cast<long long>(grp) * p.ldrb + ncol);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = __ldg(b4 + j);
            v2[2 * j] = add2(v2[2 * j], pack2(b.x, b.y));
            v2[2 * j + 1] = add2(v2[2 * j + 1], pack2(b.z, b.w));
          }
        }
        if (!PLAIN && gs != 1.0f) {
          const uint64_t gs2 = pack2(gs, gs);
#pragma unroll
          for (int j = 0; j < 16; ++j) v2[j] = mul2(v2[j], gs2);
        }
        // staging slot of this chunk: holds the residual sub-box (if any) and is free for our output
        const int slot = gc % EPI_RO_SLOTS;
        if (staged) mbar_wait(&ro_ready[slot], (gc / EPI_RO_SLOTS) & 1);
        uint8_t* ro = my_ro + slot * EPI_TILE_BYTES;
        if (!SLIM && p.residual) {
          if (p.act_f16) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 q = *reinterpret_cast<const uint4*>(ro + ((j ^ sw) << 4));
              v2[4 * j + 0] = add2(v2[4 * j + 0], h16x2_to_f32x2(q.x, true));
              v2[4 * j + 1] = add2(v2[4 * j + 1], h16x2_to_f32x2(q.y, true));
              v2[4 * j + 2] = add2(v2[4 * j + 2], h16x2_to_f32x2(q.z, true));
              v2[4 * j + 3] = add2(v2[4 * j + 3], h16x2_to_f32x2(q.w, true));
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 q = *reinterpret_cast<const uint4*>(ro + ((j ^ sw) << 4));
              v2[4 * j + 0] = add2(v2[4 * j + 0], bf16x2_to_f32x2(q.x));
              v2[4 * j + 1] = add2(v2[4 * j + 1], bf16x2_to_f32x2(q.y));
              v2[4 * j + 2] = add2(v2[4 * j + 2], bf16x2_to_f32x2(q.z));
              v2[4 * j + 3] = add2(v2[4 * j + 3], bf16x2_to_f32x2(q.w));
            }
          }
        }
        if constexpr (!PLAIN && EPI_X_SLOTS > 0) if (p.blend_x) {
          const int xs = gc % EPI_X_SLOTS;
          mbar_wait(&x_ready[xs], (gc / EPI_X_SLOTS) & 1);
          const uint8_t* xr = my_x + xs * EPI_TILE_BYTES;
          const uint64_t al2 = pack2(alpha, alpha), be2 = pack2(beta, beta);
#pragma unroll
          const bool xf16 = p.act_f16 != 0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 q = *reinterpret_cast<const uint4*>(xr + ((j ^ sw) << 4));
            v2[4 * j + 0] = fma2(al2, h16x2_to_f32x2(q.x, xf16), mul2(be2, v2[4 * j + 0]));
            v2[4 * j + 1] = fma2(al2, h16x2_to_f32x2(q.y, xf16), mul2(be2, v2[4 * j + 1]));
            v2[4 * j + 2] = fma2(al2, h16x2_to_f32x2(q.z, xf16), mul2(be2, v2[4 * j + 2]));
            v2[4 * j + 3] = fma2(al2, h16x2_to_f32x2(q.w, xf16), mul2(be2, v2[4 * j + 3]));
          }
        }
        if (p.act_f16 || (!PLAIN && p.f16_from > 0 && ncol >= p.f16_from)) {  // chunk-uniform: fp16 mode / attention V block
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(ro + ((j ^ sw) << 4)) =
                make_uint4(pack_f16x2_sat_f2(v2[4 * j]), pack_f16x2_sat_f2(v2[4 * j + 1]),
                           pack_f16x2_sat_f2(v2[4 * j + 2]), pack_f16x2_sat_f2(v2[4 * j + 3]));
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(ro + ((j ^ sw) << 4)) =
                make_uint4(pack_bf16x2_f2(v2[4 * j]), pack_bf16x2_f2(v2[4 * j + 1]), pack_bf16x2_f2(v2[4 * j + 2]),
                           pack_bf16x2_f2(v2[4 * j + 3]));
        }
        fence_proxy_async();  // staging rows visible to the TMA store
        __syncwarp();         // all 32 rows written (and this warp's reads of the X slot done)
        if (lane == 0) {
#ifndef SA_X_NOSTORE
          tma_store_4d(&tmOut, sRO + slot * EPI_TILE_BYTES, ncol, q1 + w1, q2 + w2, q3 + w3);
#endif
          tma_store_commit();
          // the slot of chunk gc-1 (== slot of chunk gc+2) is reusable once its store has read smem
          tma_store_wait_read<1>();
          if (staged) issue_next();
        }
        if constexpr (!PLAIN && !GEGLU) {
          if (p.gn_stats) {
            // Sum / Sum^2 of the bf16 values just staged (what a separate statistics pass would read back from HBM),
            // while the TMA store of the same tile is in flight.  Lane j takes COLUMN ncol + j of the 32 x 32 tile
            // (conflict-free: one 64-byte row per step), a segmented warp reduction merges the columns of a group,
            // the first lane of each group adds to the fp64 accumulators.
            const uint8_t* tile = sRO + slot * EPI_TILE_BYTES;
            const int gid = static_cast<int>(p.fd_gcpg.div(static_cast<uint32_t>(ncol + lane)));
            const bool gn_f16 = p.act_f16 != 0;
            if (gn_tile) {
              // lane = (row parity hf, 32-bit word w = column pair (2w, 2w+1)): 16 conflict-free word loads cover the
              // 32 x 32 tile; channels-per-group and chunk origins are even, so a column pair never straddles a group
              const int hf = lane >> 4, w = lane & 15;
              const uint8_t* colp = tile + hf * 64 + (w & 3) * 4;
              const int u = w >> 2;
              uint64_t sv = pack2(0.f, 0.f), qv = pack2(0.f, 0.f);
#pragma unroll
              for (int rr = 0; rr < 16; ++rr) {
                const uint64_t v = h16x2_to_f32x2(*reinterpret_cast<const uint32_t*>(colp + rr * 128 + ((u ^ (rr & 3)) << 4)), gn_f16);
                sv = add2(sv, v);
                qv = fma2(v, v, qv);
              }
              const float2 sf = unpack2(sv), qf = unpack2(qv);
              float s1 = sf.x + sf.y, s2 = qf.x + qf.y;
              s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
              s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
              // segmented suffix sums over the 16 column pairs (both half-warps mirror each other)
              const int cpg = static_cast<int>(p.fd_gcpg.d);
              const int gw = static_cast<int>(p.fd_gcpg.div(static_cast<uint32_t>(ncol + 2 * w)));
#pragma unroll
              for (int off = 1; off < 16; off <<= 1) {
                const float t1 = __shfl_down_sync(0xffffffffu, s1, off, 16);
                const float t2 = __shfl_down_sync(0xffffffffu, s2, off, 16);
                const int g2_ = __shfl_down_sync(0xffffffffu, gw, off, 16);
                if (w + off < 16 && g2_ == gw) {
                  s1 += t1;
                  s2 += t2;
                }
              }
              // lane l collects group g_first + l from the first column pair of that group inside this chunk
              const int lo = (g_first + lane) * cpg;
              const bool present = lo < ncol + EPI_CHUNK && lo + cpg > ncol;
              const int src = present ? ((lo > ncol ? lo : ncol) - ncol) >> 1 : 0;
              const float v1 = __shfl_sync(0xffffffffu, s1, src);
              const float v2 = __shfl_sync(0xffffffffu, s2, src);
              if (present) {
                ga1 += v1;
                ga2 += v2;
              }
            } else
            // bf16 values and their squares are exact in fp32; the sums run in fp64, so the statistics do not depend on
            // how rows fall into tiles / warps (B = 2 gives the bits of two B = 1 calls) and carry no fp32 partial-sum noise
            if (gn_uniform) {
              double s1 = 0.0, s2 = 0.0;
              const int cj = lane >> 3, ej = (lane & 7) * 2;
#pragma unroll 8
              for (int rr = 0; rr < 32; ++rr) {
                const float v = load16(tile + rr * 64 + ((cj ^ ((rr >> 1) & 3)) << 4) + ej, gn_f16);
                s1 += static_cast<double>(v);
                s2 += static_cast<double>(v * v);
              }
#pragma unroll
              for (int off = 1; off < 32; off <<= 1) {
                const double t1 = __shfl_down_sync(0xffffffffu, s1, off);
                const double t2 = __shfl_down_sync(0xffffffffu, s2, off);
                const int g2_ = __shfl_down_sync(0xffffffffu, gid, off);
                if (lane + off < 32 && g2_ == gid) {
                  s1 += t1;
                  s2 += t2;
                }
              }
              const int gprev = __shfl_up_sync(0xffffffffu, gid, 1);
              if (lane == 0 || gprev != gid) {
                double* dst = p.gn_stats + (static_cast<long long>(gn_sid) * p.gn_groups + gid) * 2;
                atomicAdd(dst, s1);
                atomicAdd(dst + 1, s2);
              }
            } else if (gn_sid >= 0) {
              // rows of this warp belong to different statistics (or some lie outside the grid): every lane walks
              // its own row and flushes at group boundaries
              const uint8_t* rowp = tile + lane * 64;
              double s1 = 0.0, s2 = 0.0;
              int gcur = static_cast<int>(p.fd_gcpg.div(static_cast<uint32_t>(ncol)));
              int left = (gcur + 1) * static_cast<int>(p.fd_gcpg.d) - ncol;  // columns until the group ends
              for (int j = 0; j < 32; ++j) {
                const float v = load16(rowp + (((j >> 3) ^ sw) << 4) + (j & 7) * 2, gn_f16);
                s1 += static_cast<double>(v);
                s2 += static_cast<double>(v * v);
                if (--left == 0 || j == 31) {
                  double* dst = p.gn_stats + (static_cast<long long>(gn_sid) * p.gn_groups + gcur) * 2;
                  atomicAdd(dst, s1);
                  atomicAdd(dst + 1, s2);
                  s1 = s2 = 0.0;
                  ++gcur;
                  left = static_cast<int>(p.fd_gcpg.d);
                }
              }
            }
          }
        }
        __syncwarp();
      }
      if constexpr (!SLIM) {
        if (gn_tile) {
          // tile done for this warp: fold the lane's group sums into the tile's shared fp64 set; the last of the EW
          // warps to arrive flushes the set to the global fp64 accumulators and clears it
          const int set = ti % GN_SETS;
          double* acc = gn_acc + set * 64;
          if (ga1 != 0.f || ga2 != 0.f) {
            atomicAdd(&acc[lane * 2], static_cast<double>(ga1));
            atomicAdd(&acc[lane * 2 + 1], static_cast<double>(ga2));
          }
          __threadfence_block();
          __syncwarp();
          uint32_t last = 0;
          if (lane == 0) last = (atomicAdd(&gn_cnt[set], 1u) == static_cast<uint32_t>(EW - 1)) ? 1u : 0u;
          last = __shfl_sync(0xffffffffu, last, 0);
          if (last) {
            __threadfence_block();
            volatile double* vacc = acc;
            const double t1 = vacc[lane * 2], t2 = vacc[lane * 2 + 1];
            vacc[lane * 2] = 0.0;
            vacc[lane * 2 + 1] = 0.0;
            if (g_first + lane < p.gn_groups && (t1 != 0.0 || t2 != 0.0)) {
              double* dst = p.gn_stats + (static_cast<long long>(gn_sid) * p.gn_groups + g_first + lane) * 2;
              atomicAdd(dst, t1);
              atomicAdd(dst + 1, t2);
            }
            __threadfence_block();
            __syncwarp();
            if (lane == 0) *reinterpret_cast<volatile uint32_t*>(&gn_cnt[set]) = 0u;
          }
        }
      }
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) {
    cluster_sync();   // neither CTA leaves while the partner may still signal its barriers
    if (warp == 1) tmem_dealloc_2cta<TCOLS>(tmem_base);
  } else {
    if (warp == 1) tmem_dealloc<TCOLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------ host side
template <int BN, int STAGES>
static int launch(const CUtensorMap* a0, const CUtensorMap* a1, const CUtensorMap* b,
                  const GemmKParams& p, int m_tiles, cudaStream_t stream) {
  constexpr int SMEM = STAGES * (128 * 128 + BN * 128) + 1024 + 256;
  static DeviceOnce once;
  if (int rc = once.run([&]() -> int {
        SA_CUDA_CHECK(cudaFuncSetAttribute(gemm_tap_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        return SA_OK;
      }))
    return rc;
  dim3 grid((p.N + BN - 1) / BN, m_tiles, 1);
  gemm_tap_kernel<BN, STAGES><<<grid, 128, SMEM, stream>>>(*a0, *a1, *b, p);
  SA_CUDA_CHECK(cudaGetLastError());
  return SA_OK;
}

template <int BN, int STAGES, bool GEGLU, int EW>
static int launch_persist(const CUtensorMap* a0, const CUtensorMap* a1, const CUtensorMap* b,
                          const CUtensorMap* o, const CUtensorMap* r, const CUtensorMap* x, GemmKParams p,
                          int m_tiles, cudaStream_t stream) {
  // the PLAIN instantiation has the slim staging layout: spend the shared memory it frees on pipeline stages
  constexpr int SLIM_MAX = (232448 - 2048 - EW * 2 * EPI_TILE_BYTES) / (128 * 128 + BN * 128);
  constexpr int STP = GEGLU ? STAGES : (SLIM_MAX < STAGES ? STAGES : (SLIM_MAX > 8 ? 8 : SLIM_MAX));
  using L = Gemm2Smem<BN, STAGES, EW, GEGLU>;
  using LP = Gemm2Smem<BN, STP, EW, true>;
  static_assert(L::TOTAL <= 232448 && LP::TOTAL <= 232448, "shared memory budget");
  static DeviceOnce once;
  if (int rc = once.run([&]() -> int {
        SA_CUDA_CHECK(cudaFuncSetAttribute(gemm_tap_persist_kernel<BN, STAGES, GEGLU, EW, false>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        SA_CUDA_CHECK(cudaFuncSetAttribute(gemm_tap_persist_kernel<BN, STP, GEGLU, EW, true>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, LP::TOTAL));
        return SA_OK;
      }))
    return rc;
  p.n_tiles_n = p.N / BN;
  p.total_tiles = m_tiles * p.n_tiles_n;
  p.fd_ntn.init(p.n_tiles_n);
  p.fd_t1.init(p.t1);
  p.fd_t2.init(p.t2);
  p.fd_rpg.init(p.rows_per_group);
  int grid = num_sms();
  if (grid > p.total_tiles) grid = p.total_tiles;
  const bool plain = !p.row_scale && !p.row_bias && p.out_scale == 1.0f && !p.residual && !p.blend_x && p.f16_from == 0 &&
                     !p.gn_stats;
  if (plain)
    SA_CUDA_CHECK(launch_k(gemm_tap_persist_kernel<BN, STP, GEGLU, EW, true>, dim3(grid), dim3(64 + 32 * EW), LP::TOTAL, stream, 1, true,
                           *a0, *a1, *b, *o, *r, *x, p));
  else
    SA_CUDA_CHECK(launch_k(gemm_tap_persist_kernel<BN, STAGES, GEGLU, EW, false>, dim3(grid), dim3(64 + 32 * EW), L::TOTAL, stream,
                           1, true, *a0, *a1, *b, *o, *r, *x, p));
  return SA_OK;
}

// CTA-pair launch (cta_group::2): non-GEGLU, EW epilogue warps; `b` is the weight map with a (64, BN / 2) box
template <int BN, int STAGES, int EW>
static int launch_pair(const CUtensorMap* a0, const CUtensorMap* a1, const CUtensorMap* b, const CUtensorMap* o,
                       const CUtensorMap* r, const CUtensorMap* x, GemmKParams p, int m_tiles, cudaStream_t stream) {
  using L = Gemm2Smem<BN, STAGES, EW, false, true>;
  using LP = Gemm2Smem<BN, STAGES + 1, EW, true, true>;
  static_assert(L::TOTAL <= 232448 && LP::TOTAL <= 232448, "shared memory budget");
  static DeviceOnce once;
  if (int rc = once.run([&]() -> int {
        SA_CUDA_CHECK(cudaFuncSetAttribute(gemm_tap_persist_kernel<BN, STAGES, false, EW, false, true>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        SA_CUDA_CHECK(cudaFuncSetAttribute(gemm_tap_persist_kernel<BN, STAGES + 1, false, EW, true, true>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, LP::TOTAL));
        return SA_OK;
      }))
    return rc;
  p.n_tiles_n = p.N / BN;
  p.total_tiles = ((m_tiles + 1) / 2) * p.n_tiles_n;   // work items = pairs of row-neighbouring tiles
  p.fd_ntn.init(p.n_tiles_n);
  p.fd_t1.init(p.t1);
  p.fd_t2.init(p.t2);
  p.fd_rpg.init(p.rows_per_group);
  int pairs = num_sms() / 2;
  if (pairs > p.total_tiles) pairs = p.total_tiles;
  const bool plain = !p.row_scale && !p.row_bias && p.out_scale == 1.0f && !p.residual && !p.blend_x && p.f16_from == 0 &&
                     !p.gn_stats;
  const dim3 grid(2 * pairs), block(64 + 32 * EW);
  if (plain)
    SA_CUDA_CHECK(launch_k(gemm_tap_persist_kernel<BN, STAGES + 1, false, EW, true, true>, grid, block, LP::TOTAL, stream, 2, true,
                           *a0, *a1, *b, *o, *r, *x, p));
  else
    SA_CUDA_CHECK(launch_k(gemm_tap_persist_kernel<BN, STAGES, false, EW, false, true>, grid, block, L::TOTAL, stream, 2, true,
                           *a0, *a1, *b, *o, *r, *x, p));
  return SA_OK;
}

static int pick_tile(int m, int n, int geglu, long long k = 0) {
  // cost model: an SM works through ceil(ctas / sms) tiles, each costing ~ (bn + overhead) columns
  // of MMA time; pick the exact divisor of n that minimises it (ties -> larger tile).
  const long long m_tiles = (m + 127) / 128;
  const int sms = num_sms();
  const int cand[5] = {256, 160, 128, 64, 32};
  int best = 0;
  long long best_cost = 0;
  for (int i = 0; i < 5; ++i) {
    const int bn = cand[i];
    if (n % bn != 0) continue;
    if (geglu && bn != (k >= 640 ? 256 : 128)) continue;  // measured (tools/gemm_bench.py): 256 wins from K = 640 up
    if (bn == 32 && n >= 64) continue;
    const long long ctas = m_tiles * (n / bn);
    const long long cost = ((ctas + sms - 1) / sms) * (bn + 16);
    if (best == 0 || cost < best_cost) {
      best = bn;
      best_cost = cost;
    }
  }
  if (best) return best;
  // no exact divisor: smallest tile that covers n with least padding
  if (n <= 32) return 32;
  if (n <= 64) return 64;
  return 128;
}

// choose a box (b1,b2,b3), b1*b2*b3 == 128, each a power of two <= 128, maximising the fraction of
// useful rows over the whole grid; ties go to the widest b1 (longest contiguous TMA rows).
static void pick_box(int d1, int d2, int d3, int* b1, int* b2, int* b3) {
  double best = -1.0;
  for (int e1 = 7; e1 >= 0; --e1)
    for (int e2 = 7 - e1; e2 >= 0; --e2) {
      const int x1 = 1 << e1, x2 = 1 << e2, x3 = 128 / (x1 * x2);
      const double cover = static_cast<double>((d1 + x1 - 1) / x1 * x1) *
                           ((d2 + x2 - 1) / x2 * x2) * ((d3 + x3 - 1) / x3 * x3);
      const double eff = static_cast<double>(d1) * d2 * d3 / cover;
      if (eff > best + 1e-9) {
        best = eff;
        *b1 = x1;
        *b2 = x2;
        *b3 = x3;
      }
    }
}

}  // namespace sa

extern "C" int sa_gemm_pick_tile(int32_t m, int32_t n, int32_t geglu) {
  return sa::pick_tile(m, n, geglu);
}

extern "C" int sa_gemm_tap(const sa_gemm_args* a, void* stream_) {
  using namespace sa;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!a || !a->a0 || !a->w || !a->out) return fail(SA_ERR_INVALID, "sa_gemm_tap: null pointer");
  if (a->c0 <= 0 || a->c0 % 64 || a->c1 < 0 || a->c1 % 64 || (a->c1 > 0 && !a->a1))
    return fail(SA_ERR_INVALID, "sa_gemm_tap: channels must be multiples of 64 (c0=%d c1=%d)", a->c0,
                a->c1);
  if (a->ntaps < 1 || a->ntaps > 9) return fail(SA_ERR_INVALID, "sa_gemm_tap: ntaps=%d", a->ntaps);
  if (a->n <= 0 || a->n % 8) return fail(SA_ERR_INVALID, "sa_gemm_tap: n=%d must be a multiple of 8", a->n);
  if (a->d1 <= 0 || a->d2 <= 0 || a->d3 <= 0) return fail(SA_ERR_INVALID, "sa_gemm_tap: empty grid");
  if (a->lda0 < a->c0 || a->lda0 % 8 || (a->c1 > 0 && (a->lda1 < a->c1 || a->lda1 % 8)))
    return fail(SA_ERR_INVALID, "sa_gemm_tap: bad lda");
  if (a->geglu && (a->n % 128)) return fail(SA_ERR_INVALID, "sa_gemm_tap: geglu needs n %% 128 == 0");
  const int n_out = a->geglu ? a->n / 2 : a->n;
  if (a->ldc < n_out || a->ldc % (a->out_f32 ? 4 : 8)) return fail(SA_ERR_INVALID, "sa_gemm_tap: bad ldc");
  if (a->residual && (a->ldr < n_out || a->ldr % 8)) return fail(SA_ERR_INVALID, "sa_gemm_tap: bad ldr");
  if (a->blend_x && (a->ldx < n_out || a->ldx % 8)) return fail(SA_ERR_INVALID, "sa_gemm_tap: bad ldx");
  if (a->row_bias && a->ldrb % 4) return fail(SA_ERR_INVALID, "sa_gemm_tap: ldrb must be a multiple of 4");
  const long long M = static_cast<long long>(a->d1) * a->d2 * a->d3;
  if (M > 0x7fffffffLL) return fail(SA_ERR_INVALID, "sa_gemm_tap: M too large");

  GemmKParams p;
  memset(&p, 0, sizeof(p));
  p.M = static_cast<int>(M);
  p.N = a->n;
  p.kc0 = a->c0 / 64;
  p.kcpt = (a->c0 + a->c1) / 64;
  p.ntaps = a->ntaps;
  p.d1 = a->d1;
  p.d2 = a->d2;
  p.d3 = a->d3;
  pick_box(a->d1, a->d2, a->d3, &p.b1, &p.b2, &p.b3);
  p.cs = a->in_stride > 1 ? a->in_stride : 1;
  if (p.cs > 2) return fail(SA_ERR_INVALID, "sa_gemm_tap: in_stride must be 0, 1 or 2");
  p.t1 = (a->d1 + p.b1 - 1) / p.b1;
  p.t2 = (a->d2 + p.b2 - 1) / p.b2;
  const int t3 = (a->d3 + p.b3 - 1) / p.b3;
  const long long m_tiles = static_cast<long long>(p.t1) * p.t2 * t3;
  if (m_tiles > 65535) return fail(SA_ERR_INVALID, "sa_gemm_tap: too many row tiles (%lld)", m_tiles);
  for (int i = 0; i < a->ntaps * 3; ++i) p.tap_off[i] = a->tap_off[i];
  p.bias = a->bias;
  p.row_bias = a->row_bias;
  p.ldrb = a->ldrb > 0 ? a->ldrb : (a->geglu ? a->n / 2 : a->n);
  p.row_scale = a->row_scale;
  p.rows_per_group = a->rows_per_group > 0 ? a->rows_per_group : 1;
  p.residual = static_cast<const __nv_bfloat16*>(a->residual);
  p.ldr = a->ldr;
  p.blend_x = static_cast<const __nv_bfloat16*>(a->blend_x);
  p.ldx = a->ldx;
  p.alpha = a->alpha;
  p.geglu = a->geglu;
  p.out = a->out;
  p.ldc = a->ldc;
  p.out_f32 = a->out_f32;
  p.act_f16 = a->act_f16 ? 1 : 0;

  const long long ktot_pick = static_cast<long long>(a->ntaps) * (a->c0 + a->c1);
  int bn = a->tile_n ? a->tile_n : pick_tile(p.M, p.N, p.geglu, ktot_pick);
  if (p.geglu && bn != 128 && bn != 256) return fail(SA_ERR_INVALID, "sa_gemm_tap: geglu tile must be 128/256");
  if (p.geglu && (a->residual || a->blend_x))
    return fail(SA_ERR_INVALID, "sa_gemm_tap: the geglu epilogue takes no residual / blend_x operand");

  // tensor maps: A sources as rank-4 (C, d1, d2, d3); W as rank-2 (K, N)
  const CUtensorMap* tma[2] = {nullptr, nullptr};
  for (int s = 0; s < 2; ++s) {
    const void* base = s == 0 ? a->a0 : a->a1;
    const int c = s == 0 ? a->c0 : a->c1;
    const long long ld = s == 0 ? a->lda0 : a->lda1;
    if (s == 1 && c == 0) {
      tma[1] = tma[0];
      break;
    }
    // source grid: (d1, d2, d3) positions, or (cs*d1, cs*d2, d3) sampled with TMA element strides (cs, cs)
    const uint64_t e1 = static_cast<uint64_t>(a->d1) * p.cs, e2 = static_cast<uint64_t>(a->d2) * p.cs;
    uint64_t dims[4] = {static_cast<uint64_t>(c), e1, e2, static_cast<uint64_t>(a->d3)};
    uint64_t strides[3] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(ld) * 2 * e1,
                           static_cast<uint64_t>(ld) * 2 * e1 * e2};
    uint32_t box[4] = {64, static_cast<uint32_t>(p.b1 * p.cs), static_cast<uint32_t>(p.b2 * p.cs),
                       static_cast<uint32_t>(p.b3)};
    uint32_t estr[4] = {1, static_cast<uint32_t>(p.cs), static_cast<uint32_t>(p.cs), 1};
    tma[s] = get_tmap_bf16(base, 4, dims, strides, box, 128, p.cs > 1 ? estr : nullptr);
    if (!tma[s]) return SA_ERR_TMA;
  }
  const long long ktot = static_cast<long long>(a->ntaps) * (a->c0 + a->c1);
  uint64_t wdims[2] = {static_cast<uint64_t>(ktot), static_cast<uint64_t>(a->n)};
  uint64_t wstr[1] = {static_cast<uint64_t>(ktot) * 2};
  uint32_t wbox[2] = {64, static_cast<uint32_t>(bn)};
  const CUtensorMap* tmb = get_tmap_bf16(a->w, 2, wdims, wstr, wbox);
  if (!tmb) return SA_ERR_TMA;

  const int mt = static_cast<int>(m_tiles);
  p.out_scale = a->out_scale != 0.0f ? a->out_scale : 1.0f;
  p.f16_from = a->f16_from_col;
  if (a->f16_from_col < 0 || a->f16_from_col % 32) return fail(SA_ERR_INVALID, "sa_gemm_tap: f16_from_col must be a multiple of 32");
  static const bool force_v1 = getenv("SA_GEMM_V1") != nullptr;
  const bool v2_ok = !force_v1 && !a->out_f32 && (a->n % bn == 0) && bn >= 64;
  if (a->gn_stats) {
    if (!v2_ok || a->geglu || a->f16_from_col > 0)
      return fail(SA_ERR_INVALID, "sa_gemm_tap: gn_stats needs the persistent bf16 path without geglu / f16 columns");
    if (a->gn_cpg < 1 || a->n % a->gn_cpg || a->gn_rows_per_stat < 1 || M % a->gn_rows_per_stat)
      return fail(SA_ERR_INVALID, "sa_gemm_tap: gn_cpg must divide n and gn_rows_per_stat must divide M");
    p.gn_stats = static_cast<double*>(a->gn_stats);
    p.gn_groups = a->n / a->gn_cpg;
    p.fd_gcpg.init(static_cast<uint32_t>(a->gn_cpg));
    p.fd_grps.init(static_cast<uint32_t>(a->gn_rows_per_stat));
    if (!a->gn_stats_zeroed)
      SA_CUDA_CHECK(cudaMemsetAsync(a->gn_stats, 0, sizeof(double) * 2 * p.gn_groups * (M / a->gn_rows_per_stat), stream));
  }
  if (v2_ok) {
    // epilogue tensor maps: [n_out cols, d1, d2, d3]; box = (32 columns) x the 32-row sub-box
    // (s1, s2, s3) of the tile box that one epilogue warp owns, SWIZZLE_64B
    const int s1 = p.b1 < 32 ? p.b1 : 32;
    const int s2 = p.b2 < 32 / s1 ? p.b2 : 32 / s1;
    const int s3 = 32 / (s1 * s2);
    auto epi_map = [&](const void* base, long long ld) -> const CUtensorMap* {
      uint64_t dims[4] = {static_cast<uint64_t>(n_out), static_cast<uint64_t>(a->d1),
                          static_cast<uint64_t>(a->d2), static_cast<uint64_t>(a->d3)};
      uint64_t strides[3] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(ld) * 2 * a->d1,
                             static_cast<uint64_t>(ld) * 2 * a->d1 * a->d2};
      uint32_t box[4] = {32, static_cast<uint32_t>(s1), static_cast<uint32_t>(s2), static_cast<uint32_t>(s3)};
      return get_tmap_bf16(base, 4, dims, strides, box, 64);
    };
    const CUtensorMap* tmo;
    if (a->ldc2 || a->ldc3) {  // strided output grid (Upsample2D phases)
      const long long s2_ = a->ldc2 ? a->ldc2 : static_cast<long long>(a->ldc) * a->d1;
      const long long s3_ = a->ldc3 ? a->ldc3 : s2_ * a->d2;
      if (a->residual || a->blend_x || a->gn_stats || s2_ % 8 || s3_ % 8 || s2_ < a->ldc || s3_ < s2_)
        return fail(SA_ERR_INVALID, "sa_gemm_tap: strided output (ldc2/ldc3) excludes residual / blend_x / gn_stats");
      uint64_t dims[4] = {static_cast<uint64_t>(n_out), static_cast<uint64_t>(a->d1),
                          static_cast<uint64_t>(a->d2), static_cast<uint64_t>(a->d3)};
      uint64_t strides[3] = {static_cast<uint64_t>(a->ldc) * 2, static_cast<uint64_t>(s2_) * 2,
                             static_cast<uint64_t>(s3_) * 2};
      uint32_t box[4] = {32, static_cast<uint32_t>(s1), static_cast<uint32_t>(s2), static_cast<uint32_t>(s3)};
      tmo = get_tmap_bf16(a->out, 4, dims, strides, box, 64);
    } else {
      tmo = epi_map(a->out, a->ldc);
    }
    const CUtensorMap* tmr = a->residual ? epi_map(a->residual, a->ldr) : tmo;
    const CUtensorMap* tmx = a->blend_x ? epi_map(a->blend_x, a->ldx) : tmo;
    if (!tmo || !tmr || !tmx) return SA_ERR_TMA;
    // long main loop: 4 epilogue warps, one more pipeline stage (SA_GEMM_LONGK=<k> moves the threshold, for A/B runs)
    static const long long long_k_min = getenv("SA_GEMM_LONGK") ? atoll(getenv("SA_GEMM_LONGK")) : 1280;
    const bool long_k = ktot >= long_k_min;
    // CTA pairs (cta_group::2) for the long-K family (SA_GEMM_PAIR=0: single-CTA kernels, for A/B runs)
    static const bool use_pair = !(getenv("SA_GEMM_PAIR") && atoi(getenv("SA_GEMM_PAIR")) == 0);
    if (use_pair && long_k && !p.geglu && (bn == 160 || bn == 128) && m_tiles >= 2) {
      uint32_t hbox[2] = {64, static_cast<uint32_t>(bn / 2)};
      const CUtensorMap* tmbh = get_tmap_bf16(a->w, 2, wdims, wstr, hbox);
      if (!tmbh) return SA_ERR_TMA;
      if (bn == 160) return launch_pair<160, 7, 4>(tma[0], tma[1], tmbh, tmo, tmr, tmx, p, mt, stream);
      return launch_pair<128, 7, 4>(tma[0], tma[1], tmbh, tmo, tmr, tmx, p, mt, stream);
    }
#define SA_LP(BN_, ST_, G_, EW_) return launch_persist<BN_, ST_, G_, EW_>(tma[0], tma[1], tmb, tmo, tmr, tmx, p, mt, stream)
    if (p.geglu) {
      switch (bn) {
        case 128: if (long_k) SA_LP(128, 6, true, 4); else SA_LP(128, 6, true, 8);
        case 256: SA_LP(256, 4, true, 8);
        default: break;
      }
    } else
    switch (bn) {
      case 64: SA_LP(64, 6, false, 8);
      case 128: if (long_k) SA_LP(128, 5, false, 4); else SA_LP(128, 4, false, 8);
      case 160: if (long_k) SA_LP(160, 5, false, 4); else SA_LP(160, 4, false, 8);
      case 256: SA_LP(256, 3, false, 8);
      default: break;
    }
#undef SA_LP
  }
  if (a->f16_from_col > 0)
    return fail(SA_ERR_INVALID, "sa_gemm_tap: f16_from_col needs the persistent bf16 path (n %% tile == 0, tile >= 64)");
  if (a->ldc2 || a->ldc3)
    return fail(SA_ERR_INVALID, "sa_gemm_tap: strided output needs the persistent bf16 path (n %% tile == 0, tile >= 64)");
  switch (bn) {
    case 32: return launch<32, 4>(tma[0], tma[1], tmb, p, mt, stream);
    case 64: return launch<64, 4>(tma[0], tma[1], tmb, p, mt, stream);
    case 128: return launch<128, 3>(tma[0], tma[1], tmb, p, mt, stream);
    case 160: return launch<160, 3>(tma[0], tma[1], tmb, p, mt, stream);
    case 256: return launch<256, 2>(tma[0], tma[1], tmb, p, mt, stream);
    default: return fail(SA_ERR_INVALID, "sa_gemm_tap: unsupported tile_n=%d", bn);
  }
}
