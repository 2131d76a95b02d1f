// Spatial self-attention (head_dim 64, non-causal, no mask) as a flash-attention kernel on
// tcgen05 / TMEM with TMA-staged operands, sm_100a.
//
// Replaces xformers.ops.memory_efficient_attention(q, k, v) in AnimationAttnProcessor.__call__
// (animation/modules/unet.py:1893-1905).  Reads Q, K, V straight out of the fused projection
// output [frames*S, 3C] (column blocks q | k | v, head h at columns h*64) through ONE 3-D tensor
// map (cols, S, frames), so head_to_batch_dim / .contiguous() (:1893-1895) and batch_to_head_dim
// (:1905) never materialise; rows past S are zero-filled by TMA and masked to -inf.
//
// CTA = 2 query tiles of 128 rows of one (frame, head); 640 threads:
//   warps 0-15 : softmax.  Warp w serves TMEM lane quadrant (w & 3) of tile ((w >> 2) & 1) and the 64-column half
//                (w >> 3) of each 128-key block: TWO threads share a query row, so an SM sub-partition holds four
//                softmax warps (round 1: two, each owning whole rows -- ncu: MUFU pipe 69 %, issue 35 %, tensor 34 %:
//                latency-exposed on the ex2 chain with too few warps to cover it)
//   warp 16    : TMA producer (Q tiles once; K/V blocks of 128 keys, 4-stage ring shared by both tiles)
//   warp 17    : MMA issuer + TMEM allocator      (warps 18, 19 idle: setmaxnreg moves registers per warpgroup)
// The two tiles ping-pong on the tensor core: while the softmax of S_A(j) runs the MMA pipe computes P_B V(j-1),
// S_B(j), ...  Per tile and key block:
//   S = Q K^T (128x128 fp32 in TMEM) -> softmax thread: one tcgen05.ld of its half row, max, exp2, P as
//   fp16 back into TMEM -> O += P V (TS-form MMA, accumulating in TMEM).
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace sa {

constexpr int FA_BM = 128, FA_BN = 128, FA_D = 64;
constexpr int FA_TILE = 128 * 128;  // bytes of a [128][64] 16-bit tile

// -----------------------------------------------------------------------------------------------------
// Single-pass softmax, S / P / O resident in TMEM:
//   * the softmax thread reads its half of the S row (64 fp32) with ONE tcgen05.ld and immediately hands the
//     S columns back (s_free), so the tensor core computes S(j+1) WHILE the row's softmax of block j runs;
//   * row max with 3-input FMNMX; the two threads of a row exchange their half maxima through shared memory
//     (one 64-thread named barrier per warp pair and block), so both take identical rescale decisions;
//   * exp2 on packed fp32x2 lanes (FFMA2 / FADD2).  At head_dim 64 one ex2 per score costs twice the tensor-pipe
//     time of the score's two MMAs (16 MUFU results / clk / SM), so every POLY-th group of four columns (default: an
//     eighth of them -- measured optimum, SA_FA_POLY) takes its 2^x from the FMA pipe instead (Cody-Waite split +
//     degree-3 minimax polynomial, |rel err| <= 7.5e-5, below the fp16 rounding of P): the MUFU pipe, not the
//     tensor pipe, is this kernel's roofline;
//   * P goes back into TMEM (its own 64 columns per tile, tcgen05.st) as fp16 pairs: P is in (0, 256], so half
//     precision holds it with an 11-bit significand (bf16: 8).  kind::f16 needs A and B in the same format (a
//     mixed fp16 x bf16 descriptor traps), so the V block of qkv is fp16 too: the q|k|v projection writes it
//     that way (sa_gemm_tap f16_from_col = 2C);
//   * O += P V is a TS-form tcgen05.mma (A = P from TMEM, B = V MN-major from the TMA tile) that
//     ACCUMULATES into O in TMEM over all key blocks -- no per-block fold through registers;
//   * the running max is updated lazily: O and l are rescaled (TMEM read-modify-write, each thread of the row
//     its 32 O columns) only when a block's max exceeds the reference max by more than 2^8, so P <= 256 and the
//     common case has no correction work at all.  The decision is warp-uniform (any-lane vote) and identical in
//     both warps of a pair.
// TMEM columns (all 512): S_A [0,128) S_B [128,256) O_A [256,320) O_B [320,384) P_A [384,448) P_B [448,512).
// smem: Q 2x16K | K 4x16K | V 4x16K = 160 KB + barriers + 4 KB exchange -> one CTA per SM.
// -----------------------------------------------------------------------------------------------------
// NT = query tiles per CTA.  NT = 2: one CTA per SM, the two tiles ping-pong on the tensor core and share the K / V ring.
// NT = 1 (SA_FA_NT=1): two co-resident CTAs per SM of one tile each (256 TMEM columns, 80 KB of shared memory, 8 softmax
// warps): the prologue / epilogue of one CTA (TMEM allocation, Q load, first S; final O read-out) runs under the other
// CTA's main loop instead of leaving the SM idle, at the price of loading K / V once per tile.
template <int NT> struct Fa3Cfg {
  static constexpr int ST = NT == 2 ? 4 : 2;          // K / V ring stages
  static constexpr int SW = 8 * NT;                   // softmax warps
  // NT = 2: + one warpgroup (TMA, MMA, two idle warps: setmaxnreg is per warpgroup); NT = 1: + TMA and MMA warps only --
  // 2 x 320 threads leave each thread 96 registers, which the softmax code fits without rebalancing
  static constexpr int THREADS = (SW + 4) * 32;
  static constexpr bool REBALANCE = true;
  static constexpr int TCOLS = 256 * NT;              // S | O | P
  static constexpr int O_COL = 128 * NT, P_COL = 192 * NT;
  static constexpr int REGS0 = NT == 2 ? 96 : 80;     // registers per thread the CTA is launched with
  static constexpr int REGS_CTRL = 24, REGS_SOFTMAX = NT == 2 ? 112 : 104;
  static_assert(!REBALANCE || SW * 32 * REGS_SOFTMAX + 4 * 32 * REGS_CTRL <= THREADS * REGS0, "register pool of the CTA");
};
constexpr int FA3_ST = 4;                       // (barrier arrays are sized for the larger ring)
constexpr int FA3_SW = 16;
constexpr int FA3_THREADS = (FA3_SW + 4) * 32;
constexpr float FA3_TAU = 8.0f;  // lazy-rescale threshold (log2 units)
// columns whose 2^x comes from the FMA-pipe polynomial: every POLY-th group of 4 columns (template parameter; 0 = none)

struct Fa3Bars {
  uint64_t q_full;
  uint64_t k_full[FA3_ST], v_full[FA3_ST], kv_empty[FA3_ST];
  uint64_t s_full[2], s_free[2], o_full[2], p_full[2];
  uint32_t tmem_slot;
  float xch[2][8][2][32];  // [block parity][warp pair][column half][lane]: half-row maxima / final row sums
};
template <int NT> constexpr int fa3_smem() {
  return FA_TILE * (NT + 2 * Fa3Cfg<NT>::ST) + 1024 + static_cast<int>(sizeof(Fa3Bars)) + 64;
}

// 2^x for two lanes on the FMA pipe: x = n + f, n = round(x), f in [-0.5, 0.5]; 2^f by a degree-3 minimax
// polynomial (relative error <= 7.5e-5), 2^n by adding n to the exponent field.  x is clamped at -126 (2^-126 ~ 1e-38
// stands in for the zeros of masked columns); no MUFU instruction.
__device__ __forceinline__ float exp2_poly(float x) {
  // scalar form: every constant is an instruction immediate (the packed form keeps seven 64-bit constants in
  // registers, which this kernel cannot afford: 120 registers per softmax thread)
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;            // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = x - (t - 12582912.0f);      // [-0.5, 0.5]
  float p = fmaf(5.517166771e-02f, f, 2.426111221e-01f);
  p = fmaf(p, f, 6.932609857e-01f);
  p = fmaf(p, f, 9.999280736e-01f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));
}
__device__ __forceinline__ void exp2_poly2(uint64_t x, float& e0, float& e1) {
  const float2 xf = unpack2(x);
  e0 = exp2_poly(xf.x);
  e1 = exp2_poly(xf.y);
}

template <int POLY, int NT>
__global__ void __launch_bounds__(Fa3Cfg<NT>::THREADS, NT == 2 ? 1 : 2)
attn_flash_kernel(const __grid_constant__ CUtensorMap tm, int S, int C, __nv_bfloat16* __restrict__ out,
                  int ldo, int f16) {   // f16: q, k and out are IEEE fp16 like v (the fp16 activation mode)
  using Cfg = Fa3Cfg<NT>;
  constexpr int ST = Cfg::ST, SW = Cfg::SW, THREADS = Cfg::THREADS, O_COL = Cfg::O_COL, P_COL = Cfg::P_COL;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                          // NT tiles
  uint8_t* sK = smem + NT * FA_TILE;           // ST stages
  uint8_t* sV = sK + ST * FA_TILE;             // ST stages
  Fa3Bars* bars = reinterpret_cast<Fa3Bars*>(sV + ST * FA_TILE);

  const int warp = threadIdx.x >> 5;
  const int q0 = blockIdx.x * NT * FA_BM;
  const int head = blockIdx.y;
  const int frame = blockIdx.z;
  const int nblk = (S + FA_BN - 1) / FA_BN;
  const int ntiles = (NT == 2 && q0 + FA_BM < S) ? 2 : 1;

  if (warp == SW && elect_one()) {
    tma_prefetch_desc(&tm);
    mbar_init(&bars->q_full, 1);
    for (int i = 0; i < ST; ++i) {
      mbar_init(&bars->k_full[i], 1);
      mbar_init(&bars->v_full[i], 1);
      mbar_init(&bars->kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->s_full[i], 1);
      mbar_init(&bars->o_full[i], 1);
      mbar_init(&bars->s_free[i], 8);   // one elected arrival per softmax warp of the tile (8 warps), not one per thread:
      mbar_init(&bars->p_full[i], 8);   // 256 arrivals on one mbarrier serialise on the same shared-memory word
    }
    fence_barrier_init();
  }
  if (warp == SW + 1) tmem_alloc<Cfg::TCOLS>(&bars->tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;
  grid_dep_sync();   // programmatic dependent launch: the set-up above overlapped the previous kernel's tail
  // Register rebalancing: the two branches below never re-join (each ends with its own CTA-wide barrier and return),
  // so ptxas allocates them under different budgets.
  if (warp >= SW) {
    if constexpr (Cfg::REBALANCE) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::REGS_CTRL));
  if (warp == SW) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_expect_tx(&bars->q_full, ntiles * FA_TILE);
      for (int t = 0; t < ntiles; ++t)
        tma_load_3d(sQ + t * FA_TILE, &tm, &bars->q_full, head * FA_D, q0 + t * FA_BM, frame);
      for (int j = 0; j < nblk; ++j) {
        const int st = j % ST;
        mbar_wait(&bars->kv_empty[st], ((j / ST) & 1) ^ 1);
        mbar_expect_tx(&bars->k_full[st], FA_TILE);
        tma_load_3d(sK + st * FA_TILE, &tm, &bars->k_full[st], C + head * FA_D, j * FA_BN, frame);
        mbar_expect_tx(&bars->v_full[st], FA_TILE);
        tma_load_3d(sV + st * FA_TILE, &tm, &bars->v_full[st], 2 * C + head * FA_D, j * FA_BN, frame);
      }
    }
    __syncwarp();
  } else if (warp == SW + 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t idesc_s = f16 ? umma_idesc_f16(128, FA_BN, 0, 0) : umma_idesc_bf16(128, FA_BN, 0, 0);  // S = Q K^T
      constexpr uint32_t idesc_o = umma_idesc_f16(128, FA_D, 0, 1);    // O += P V, fp16 x fp16 (V MN-major)
      auto issue_s = [&](int t, int j) {
        const uint64_t dq = umma_desc_sw128(smem_u32(sQ + t * FA_TILE));
        const uint64_t dk = umma_desc_sw128(smem_u32(sK + (j % ST) * FA_TILE));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem_base + t * 128, dq + 2 * k, dk + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&bars->s_full[t]);
      };
      mbar_wait(&bars->q_full, 0);
      mbar_wait(&bars->k_full[0], 0);
      tc_fence_after();
      for (int t = 0; t < ntiles; ++t) issue_s(t, 0);
      for (int j = 0; j < nblk; ++j) {
        const int st = j % ST;
        if (j + 1 < nblk) {
          // S(j+1) as soon as the softmax threads hold S(j) in registers
          mbar_wait(&bars->k_full[(j + 1) % ST], ((j + 1) / ST) & 1);
          for (int t = 0; t < ntiles; ++t) {
            mbar_wait(&bars->s_free[t], j & 1);
            tc_fence_after();
            issue_s(t, j + 1);
          }
        }
        mbar_wait(&bars->v_full[st], (j / ST) & 1);
        for (int t = 0; t < ntiles; ++t) {
          mbar_wait(&bars->p_full[t], j & 1);  // P_t(j) in TMEM, O_t corrected if it had to be
          tc_fence_after();
          const uint32_t p_tmem = tmem_base + P_COL + t * 64;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint64_t dv = umma_desc_sw128(smem_u32(sV + st * FA_TILE + kk * 2048));
            umma_f16_ts(tmem_base + O_COL + t * 64, p_tmem + kk * 8, dv, idesc_o, (j | kk) != 0 ? 1u : 0u);
          }
          umma_commit(&bars->o_full[t]);
        }
        umma_commit(&bars->kv_empty[st]);  // K_j (read by S(j)) and V_j (read by PV(j)) are free
      }
    }
    __syncwarp();
  }
    tc_fence_before();
    named_bar_sync(15, THREADS);   // every softmax thread is done with TMEM
    if (warp == SW + 1) tmem_dealloc<Cfg::TCOLS>(tmem_base);
    return;
  }
  if constexpr (Cfg::REBALANCE) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::REGS_SOFTMAX));
  {
    // ------------------------------------------------------------ softmax warps (two threads per query row)
    const int quad = warp & 3;               // TMEM lane quadrant
    const int t = (warp >> 2) & (NT - 1);    // tile handled by this warp
    const int hc = warp >> (NT == 2 ? 3 : 2);   // column half of the key block
    const int pair = warp & (4 * NT - 1);    // (tile, quadrant): the two warps sharing these 32 rows
    const int lane = threadIdx.x & 31;
    if (t < ntiles) {
      const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
      const uint32_t s_addr = lane_addr + t * 128 + hc * 64;
      const uint32_t o_addr = lane_addr + O_COL + t * 64 + hc * 32;   // this thread's half of the row's O columns
      const uint32_t p_addr = lane_addr + P_COL + t * 64 + hc * 32;
      const float sl2 = 0.125f * 1.4426950408889634f;  // attn.scale * log2(e)
      const uint64_t sl22 = pack2(sl2, sl2);
      float m_ref = -INFINITY;                 // reference max the exponentials are taken against
      uint64_t l2 = pack2(0.f, 0.f);           // partial row sum over this thread's columns (two interleaved sums)

      for (int j = 0; j < nblk; ++j) {
        const int kv_left = S - j * FA_BN - hc * 64;   // valid columns of this thread's half
        mbar_wait(&bars->s_full[t], j & 1);
        tc_fence_after();
        uint32_t s[64];
        if constexpr (NT == 2) {
          tmem_ld_32x64(s_addr, s);
        } else {   // two 32-column loads: the 64-register destination block of the x64 form does not fit next to the rest
          tmem_ld_32x32(s_addr, reinterpret_cast<uint32_t(&)[32]>(s[0]));
          tmem_ld_32x32(s_addr + 32, reinterpret_cast<uint32_t(&)[32]>(s[32]));
        }
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->s_free[t]);  // the tensor core may overwrite S_t with S_t(j+1) now
        if (kv_left < 64) {
#pragma unroll
          for (int c = 0; c < 64; ++c)
            if (c >= kv_left) s[c] = 0xff800000u;  // -inf
        }
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < 64; c += 8) {
          mx[0] = max3(mx[0], __uint_as_float(s[c]), __uint_as_float(s[c + 1]));
          mx[1] = max3(mx[1], __uint_as_float(s[c + 2]), __uint_as_float(s[c + 3]));
          mx[2] = max3(mx[2], __uint_as_float(s[c + 4]), __uint_as_float(s[c + 5]));
          mx[3] = max3(mx[3], __uint_as_float(s[c + 6]), __uint_as_float(s[c + 7]));
        }
        float bm = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * sl2;
        // the row's other half: exchange through shared memory (double-buffered by block parity)
        bars->xch[j & 1][pair][hc][lane] = bm;
        named_bar_sync(1 + pair, 64);
        bm = fmaxf(bm, bars->xch[j & 1][pair][hc ^ 1][lane]);
        const bool need = bm > m_ref + FA3_TAU;   // always true on the first block (m_ref = -inf)
        if (__any_sync(0xffffffffu, need)) {
          const float m_new = need ? bm : m_ref;
          const float alpha = ex2_approx(m_ref - m_new);  // 1 for lanes that keep their max, 0 on block 0
          m_ref = m_new;
          const uint64_t a2 = pack2(alpha, alpha);
          l2 = mul2(l2, a2);
          if (j > 0) {
            // rescale O in TMEM: P V(j-1) must have completed; P V(j) is gated by our p_full arrival
            mbar_wait(&bars->o_full[t], (j - 1) & 1);
            tc_fence_after();
            uint32_t o[32];
            tmem_ld_32x32(o_addr, o);
            tmem_ld_wait();
#pragma unroll
            for (int d = 0; d < 32; d += 2) {
              const float2 f = unpack2(mul2(pack2u(o[d], o[d + 1]), a2));
              o[d] = __float_as_uint(f.x);
              o[d + 1] = __float_as_uint(f.y);
            }
            tmem_st_32x32(o_addr, o);
          }
        }
        const uint64_t nm2 = pack2(-m_ref, -m_ref);
        uint64_t la = pack2(0.f, 0.f), lb = pack2(0.f, 0.f);
        if (j > 0) {  // P V(j-1) must be done reading P_t before it is overwritten (normally long since)
          mbar_wait(&bars->o_full[t], (j - 1) & 1);
          tc_fence_after();
        }
        // two halves of 32 columns, each stored to TMEM as soon as it is packed (keeps the live registers down)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t pw[16];
#pragma unroll
          for (int cc = 0; cc < 32; cc += 4) {
            const int c = hh * 32 + cc;
            const uint64_t x0 = fma2(pack2u(s[c], s[c + 1]), sl22, nm2);
            const uint64_t x1 = fma2(pack2u(s[c + 2], s[c + 3]), sl22, nm2);
            float e0, e1, e2, e3;
            if (POLY > 0 && ((c >> 2) % (POLY > 0 ? POLY : 1)) == POLY - 1) {   // compile-time after unrolling
              exp2_poly2(x0, e0, e1);
              exp2_poly2(x1, e2, e3);
            } else {
              const float2 f0 = unpack2(x0), f1 = unpack2(x1);
              e0 = ex2_approx(f0.x), e1 = ex2_approx(f0.y);
              e2 = ex2_approx(f1.x), e3 = ex2_approx(f1.y);
            }
            la = add2(la, pack2(e0, e1));
            lb = add2(lb, pack2(e2, e3));
            pw[cc >> 1] = pack_f16x2(e0, e1);        // P <= 2^8 (lazy max): fp16 keeps 11 significant bits
            pw[(cc >> 1) + 1] = pack_f16x2(e2, e3);
          }
          tmem_st_32x16(p_addr + hh * 16, pw);
        }
        l2 = add2(l2, add2(la, lb));
        tmem_st_wait();              // P (and a rescaled O) are in TMEM
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->p_full[t]);
      }
      // all P V accumulated: total row sum = this thread's half + the partner's
      const float2 lf = unpack2(l2);
      const float lh = lf.x + lf.y;
      bars->xch[nblk & 1][pair][hc][lane] = lh;
      named_bar_sync(1 + pair, 64);
      const float inv = 1.0f / (lh + bars->xch[nblk & 1][pair][hc ^ 1][lane]);
      mbar_wait(&bars->o_full[t], (nblk - 1) & 1);
      tc_fence_after();
      uint32_t o[32];
      tmem_ld_32x32(o_addr, o);
      tmem_ld_wait();
      const int qrow = q0 + t * FA_BM + quad * 32 + lane;
      if (qrow < S) {
        uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<long long>(frame) * S + qrow) * ldo + head * FA_D + hc * 32);
        const bool h16 = f16 != 0;
#pragma unroll
        for (int g = 0; g < 4; ++g)
          dst[g] = make_uint4(pack16x2(__uint_as_float(o[8 * g]) * inv, __uint_as_float(o[8 * g + 1]) * inv, h16),
                              pack16x2(__uint_as_float(o[8 * g + 2]) * inv, __uint_as_float(o[8 * g + 3]) * inv, h16),
                              pack16x2(__uint_as_float(o[8 * g + 4]) * inv, __uint_as_float(o[8 * g + 5]) * inv, h16),
                              pack16x2(__uint_as_float(o[8 * g + 6]) * inv, __uint_as_float(o[8 * g + 7]) * inv, h16));
      }
    }
  }

  tc_fence_before();
  named_bar_sync(15, THREADS);
}

}  // namespace sa

using namespace sa;

// qkv: bf16 [frames*S, ld] with q | k | v column blocks of width C = heads*64; out: bf16 [frames*S, ldo]
extern "C" int sa_attn_spatial(const void* qkv, int32_t ld, int32_t frames, int32_t S, int32_t C,
                               int32_t heads, void* out, int32_t ldo, int32_t act_f16, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!qkv || !out) return fail(SA_ERR_INVALID, "sa_attn_spatial: null pointer");
  if (C != heads * 64 || ld < 3 * C || ld % 8 || ldo % 8)
    return fail(SA_ERR_INVALID, "sa_attn_spatial: head_dim must be 64 and ld >= 3C");
  if (frames < 1 || frames > 65535 || S < 1) return fail(SA_ERR_INVALID, "sa_attn_spatial: bad sizes");
  uint64_t dims[3] = {static_cast<uint64_t>(3 * C), static_cast<uint64_t>(S), static_cast<uint64_t>(frames)};
  uint64_t strides[2] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(ld) * 2 * S};
  uint32_t box[3] = {64, 128, 1};
  const CUtensorMap* tm = get_tmap_bf16(qkv, 3, dims, strides, box);
  if (!tm) return SA_ERR_TMA;
  // SA_FA_POLY=<n>: every n-th group of four score columns takes its exp2 from the FMA pipe (0 = MUFU only); A/B knob.
  // SA_FA_NT=1: one query tile per CTA, two CTAs per SM (see Fa3Cfg); default 2.
  static const int poly = getenv("SA_FA_POLY") ? atoi(getenv("SA_FA_POLY")) : 8;
  static const int nt = getenv("SA_FA_NT") ? atoi(getenv("SA_FA_NT")) : 2;
  static DeviceOnce once;
  if (int rc = once.run([&]() -> int {
        SA_CUDA_CHECK(cudaFuncSetAttribute(attn_flash_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, fa3_smem<2>()));
        SA_CUDA_CHECK(cudaFuncSetAttribute(attn_flash_kernel<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, fa3_smem<2>()));
        SA_CUDA_CHECK(cudaFuncSetAttribute(attn_flash_kernel<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, fa3_smem<2>()));
        SA_CUDA_CHECK(cudaFuncSetAttribute(attn_flash_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, fa3_smem<1>()));
        SA_CUDA_CHECK(cudaFuncSetAttribute(attn_flash_kernel<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, fa3_smem<1>()));
        return SA_OK;
      }))
    return rc;
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
  if (nt == 1) {
    dim3 grid((S + FA_BM - 1) / FA_BM, heads, frames);
    if (poly == 0) SA_CUDA_CHECK(launch_k(attn_flash_kernel<0, 1>, grid, dim3(Fa3Cfg<1>::THREADS), fa3_smem<1>(), stream, 1, true, *tm, S, C, o, ldo, act_f16));
    else SA_CUDA_CHECK(launch_k(attn_flash_kernel<8, 1>, grid, dim3(Fa3Cfg<1>::THREADS), fa3_smem<1>(), stream, 1, true, *tm, S, C, o, ldo, act_f16));
  } else {
    dim3 grid((S + 2 * FA_BM - 1) / (2 * FA_BM), heads, frames);
    switch (poly) {
      case 0: SA_CUDA_CHECK(launch_k(attn_flash_kernel<0, 2>, grid, dim3(Fa3Cfg<2>::THREADS), fa3_smem<2>(), stream, 1, true, *tm, S, C, o, ldo, act_f16)); break;
      case 4: SA_CUDA_CHECK(launch_k(attn_flash_kernel<4, 2>, grid, dim3(Fa3Cfg<2>::THREADS), fa3_smem<2>(), stream, 1, true, *tm, S, C, o, ldo, act_f16)); break;
      default: SA_CUDA_CHECK(launch_k(attn_flash_kernel<8, 2>, grid, dim3(Fa3Cfg<2>::THREADS), fa3_smem<2>(), stream, 1, true, *tm, S, C, o, ldo, act_f16)); break;
    }
  }
  return SA_OK;
}
