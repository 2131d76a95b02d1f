// Spatial self-attention (head_dim 64, non-causal, no mask) as a flash-attention kernel on
// tcgen05 / TMEM with TMA-staged operands, sm_100a.
//
// Replaces xformers.ops.memory_efficient_attention(q, k, v) in AnimationAttnProcessor.__call__
// (animation/modules/unet.py:1893-1905).  Reads Q, K, V straight out of the fused projection
// output [frames*S, 3C] (column blocks q | k | v, head h at columns h*64) through ONE 3-D tensor
// map (cols, S, frames), so head_to_batch_dim / .contiguous() (:1893-1895) and batch_to_head_dim
// (:1905) never materialise; rows past S are zero-filled by TMA and masked to -inf.
//
// CTA = 2 query tiles of 128 rows of one (frame, head); 320 threads:
//   warps 0-3 / 4-7 : softmax warpgroup of tile A / tile B, one thread per query row (TMEM lane)
//   warp 8          : TMA producer (Q tiles once; K/V blocks of 128 keys, 4-stage ring shared by
//                     both tiles)
//   warp 9          : MMA issuer + TMEM allocator
// The two tiles ping-pong on the tensor core: while warpgroup A runs the softmax of S_A(j) the MMA
// pipe computes P_B V(j-1), S_B(j), ...  Per tile and key block:
//   S = Q K^T (128x128 fp32 in TMEM) -> softmax thread: one tcgen05.ld of its row, max, exp2, P as
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
//   * the softmax thread reads its whole S row (128 fp32) with ONE tcgen05.ld and immediately hands the
//     S columns back (s_free), so the tensor core computes S(j+1) WHILE the row's softmax of block j
//     runs -- the softmax warps never wait for Q K^T in steady state;
//   * row max with 3-input FMNMX, exp2 on packed fp32x2 lanes (FFMA2 / FADD2); P goes back into TMEM
//     (its own 64 columns per tile, tcgen05.st) as fp16 pairs: P is in (0, 256], so half precision holds
//     it with an 11-bit significand (bf16: 8).  kind::f16 needs A and B in the same format (a mixed
//     fp16 x bf16 descriptor traps), so the V block of qkv is fp16 too: the q|k|v projection writes it
//     that way (sa_gemm_tap f16_from_col = 2C);
//   * O += P V is a TS-form tcgen05.mma (A = P from TMEM, B = V MN-major from the TMA tile) that
//     ACCUMULATES into O in TMEM over all key blocks -- no per-block fold through registers;
//   * the running max is updated lazily: O and l are rescaled (TMEM read-modify-write by the row's own
//     thread) only when a block's max exceeds the reference max by more than 2^8, so P <= 256 and the
//     common case has no correction work at all.  The decision is warp-uniform (any-lane vote);
//   * the exp2 phase saturates a sub-partition's MUFU pipe on its own (1024 cycles per warp and block;
//     measured: tools/probes/exp_loop.cu).  Left alone, the two softmax warps of a sub-partition (tile A,
//     tile B) fall into lock step -- both in the exp phase at half rate, then both in the MUFU-free phase
//     (tcgen05.ld, max, barriers) with the pipe idle -- so they take turns through a pair of mbarriers:
//     A(j), B(j), A(j+1), ... puts one warp's MUFU-free work under the other's exp phase (worth 4 %;
//     ablations on B200 at S=4096: 0.90 ms as is, 0.64 ms with the MUFU instructions removed, i.e. the
//     kernel runs at 70 % MUFU-pipe utilisation and 34 % tensor-pipe utilisation -- at head_dim 64 one
//     exp2 per score costs twice the tensor-pipe time of its two MMAs, so MUFU is the roofline here).
// TMEM columns (all 512): S_A [0,128) S_B [128,256) O_A [256,320) O_B [320,384) P_A [384,448) P_B [448,512).
// smem: Q 2x16K | K 4x16K | V 4x16K = 160 KB + barriers -> one CTA per SM.
// -----------------------------------------------------------------------------------------------------
constexpr int FA3_ST = 4;
constexpr int FA3_SMEM = FA_TILE * (2 + 2 * FA3_ST) + 1024 + 256;
constexpr float FA3_TAU = 8.0f;  // lazy-rescale threshold (log2 units)

struct Fa3Bars {
  uint64_t q_full;
  uint64_t k_full[FA3_ST], v_full[FA3_ST], kv_empty[FA3_ST];
  uint64_t s_full[2], s_free[2], o_full[2], p_full[2];
  uint64_t order[2][4];  // MUFU turn-taking between the two softmax warps of an SM sub-partition
  uint32_t tmem_slot;
};

__global__ void __launch_bounds__(320, 1)
attn_flash_kernel(const __grid_constant__ CUtensorMap tm, int S, int C, __nv_bfloat16* __restrict__ out,
                  int ldo) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                          // 2 tiles
  uint8_t* sK = smem + 2 * FA_TILE;            // FA3_ST stages
  uint8_t* sV = sK + FA3_ST * FA_TILE;         // FA3_ST stages
  Fa3Bars* bars = reinterpret_cast<Fa3Bars*>(sV + FA3_ST * FA_TILE);

  const int warp = threadIdx.x >> 5;
  const int q0 = blockIdx.x * 2 * FA_BM;
  const int head = blockIdx.y;
  const int frame = blockIdx.z;
  const int nblk = (S + FA_BN - 1) / FA_BN;
  const int ntiles = (q0 + FA_BM < S) ? 2 : 1;

  if (warp == 8 && elect_one()) {
    tma_prefetch_desc(&tm);
    mbar_init(&bars->q_full, 1);
    for (int i = 0; i < FA3_ST; ++i) {
      mbar_init(&bars->k_full[i], 1);
      mbar_init(&bars->v_full[i], 1);
      mbar_init(&bars->kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->s_full[i], 1);
      mbar_init(&bars->o_full[i], 1);
      mbar_init(&bars->s_free[i], 128);
      mbar_init(&bars->p_full[i], 128);
      for (int q = 0; q < 4; ++q) mbar_init(&bars->order[i][q], 32);
    }
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc<512>(&bars->tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;

  if (warp == 8) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_expect_tx(&bars->q_full, ntiles * FA_TILE);
      for (int t = 0; t < ntiles; ++t)
        tma_load_3d(sQ + t * FA_TILE, &tm, &bars->q_full, head * FA_D, q0 + t * FA_BM, frame);
      for (int j = 0; j < nblk; ++j) {
        const int st = j % FA3_ST;
        mbar_wait(&bars->kv_empty[st], ((j / FA3_ST) & 1) ^ 1);
        mbar_expect_tx(&bars->k_full[st], FA_TILE);
        tma_load_3d(sK + st * FA_TILE, &tm, &bars->k_full[st], C + head * FA_D, j * FA_BN, frame);
        mbar_expect_tx(&bars->v_full[st], FA_TILE);
        tma_load_3d(sV + st * FA_TILE, &tm, &bars->v_full[st], 2 * C + head * FA_D, j * FA_BN, frame);
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, FA_BN, 0, 0);  // S = Q K^T (bf16, both K-major)
      constexpr uint32_t idesc_o = umma_idesc_f16(128, FA_D, 0, 1);    // O += P V, fp16 x fp16 (V MN-major)
      auto issue_s = [&](int t, int j) {
        const uint64_t dq = umma_desc_sw128(smem_u32(sQ + t * FA_TILE));
        const uint64_t dk = umma_desc_sw128(smem_u32(sK + (j % FA3_ST) * FA_TILE));
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
        const int st = j % FA3_ST;
        if (j + 1 < nblk) {
          // S(j+1) as soon as the softmax threads hold S(j) in registers
          mbar_wait(&bars->k_full[(j + 1) % FA3_ST], ((j + 1) / FA3_ST) & 1);
          for (int t = 0; t < ntiles; ++t) {
            mbar_wait(&bars->s_free[t], j & 1);
            tc_fence_after();
            issue_s(t, j + 1);
          }
        }
        mbar_wait(&bars->v_full[st], (j / FA3_ST) & 1);
        for (int t = 0; t < ntiles; ++t) {
          mbar_wait(&bars->p_full[t], j & 1);  // P_t(j) in TMEM, O_t corrected if it had to be
          tc_fence_after();
          const uint32_t p_tmem = tmem_base + 384 + t * 64;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint64_t dv = umma_desc_sw128(smem_u32(sV + st * FA_TILE + kk * 2048));
            umma_f16_ts(tmem_base + 256 + t * 64, p_tmem + kk * 8, dv, idesc_o, (j | kk) != 0 ? 1u : 0u);
          }
          umma_commit(&bars->o_full[t]);
        }
        umma_commit(&bars->kv_empty[st]);  // K_j (read by S(j)) and V_j (read by PV(j)) are free
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ softmax warpgroups (row = thread)
    const int t = warp >> 2;                 // tile handled by this warpgroup
    const int r = threadIdx.x & 127;         // row in tile == TMEM lane
    if (t < ntiles) {
      const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
      const uint32_t s_addr = lane_addr + t * 128;
      const uint32_t o_addr = lane_addr + 256 + t * 64;
      const uint32_t p_addr = lane_addr + 384 + t * 64;
      const float sl2 = 0.125f * 1.4426950408889634f;  // attn.scale * log2(e)
      const uint64_t sl22 = pack2(sl2, sl2);
      float m_ref = -INFINITY;                 // reference max the exponentials are taken against
      uint64_t l2 = pack2(0.f, 0.f);           // row sum (two interleaved partial sums)
      uint64_t* my_turn = &bars->order[t][warp & 3];
      uint64_t* their_turn = &bars->order[t ^ 1][warp & 3];
      const bool take_turns = ntiles == 2;

      for (int j = 0; j < nblk; ++j) {
        const int kv_left = S - j * FA_BN;
        mbar_wait(&bars->s_full[t], j & 1);
        tc_fence_after();
        uint32_t s[128];
        tmem_ld_32x128(s_addr, s);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&bars->s_free[t]);  // the tensor core may overwrite S_t with S_t(j+1) now
        if (kv_left < FA_BN) {
#pragma unroll
          for (int c = 0; c < FA_BN; ++c)
            if (c >= kv_left) s[c] = 0xff800000u;  // -inf
        }
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < FA_BN; c += 8) {
          mx[0] = max3(mx[0], __uint_as_float(s[c]), __uint_as_float(s[c + 1]));
          mx[1] = max3(mx[1], __uint_as_float(s[c + 2]), __uint_as_float(s[c + 3]));
          mx[2] = max3(mx[2], __uint_as_float(s[c + 4]), __uint_as_float(s[c + 5]));
          mx[3] = max3(mx[3], __uint_as_float(s[c + 6]), __uint_as_float(s[c + 7]));
        }
        const float bm = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * sl2;
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
            uint32_t o[64];
            tmem_ld_32x64(o_addr, o);
            tmem_ld_wait();
#pragma unroll
            for (int d = 0; d < FA_D; d += 2) {
              const float2 f = unpack2(mul2(pack2u(o[d], o[d + 1]), a2));
              o[d] = __float_as_uint(f.x);
              o[d + 1] = __float_as_uint(f.y);
            }
            tmem_st_32x64(o_addr, o);
          }
        }
        const uint64_t nm2 = pack2(-m_ref, -m_ref);
        uint32_t pw[64];
        uint64_t la = pack2(0.f, 0.f), lb = pack2(0.f, 0.f);
        // my turn on the MUFU pipe: tile A's block j follows tile B's block j-1, tile B's follows tile A's
        if (take_turns && (t == 1 || j > 0)) mbar_wait(my_turn, (t == 1 ? j : j - 1) & 1);
#pragma unroll
        for (int c = 0; c < FA_BN; c += 4) {
          const float2 x0 = unpack2(fma2(pack2u(s[c], s[c + 1]), sl22, nm2));
          const float2 x1 = unpack2(fma2(pack2u(s[c + 2], s[c + 3]), sl22, nm2));
          const float e0 = ex2_approx(x0.x), e1 = ex2_approx(x0.y);
          const float e2 = ex2_approx(x1.x), e3 = ex2_approx(x1.y);
          la = add2(la, pack2(e0, e1));
          lb = add2(lb, pack2(e2, e3));
          pw[c >> 1] = pack_f16x2(e0, e1);        // P <= 2^8 (lazy max): fp16 keeps 11 significant bits
          pw[(c >> 1) + 1] = pack_f16x2(e2, e3);
        }
        if (take_turns) mbar_arrive(their_turn);  // the other tile's warp may start its exp phase
        l2 = add2(l2, add2(la, lb));
        if (j > 0) {  // P V(j-1) must be done reading P_t before it is overwritten (normally long since)
          mbar_wait(&bars->o_full[t], (j - 1) & 1);
          tc_fence_after();
        }
        tmem_st_32x64(p_addr, pw);
        tmem_st_wait();              // P (and a rescaled O) are in TMEM
        tc_fence_before();
        mbar_arrive(&bars->p_full[t]);
      }
      // all P V accumulated
      mbar_wait(&bars->o_full[t], (nblk - 1) & 1);
      tc_fence_after();
      uint32_t o[64];
      tmem_ld_32x64(o_addr, o);
      tmem_ld_wait();
      const int qrow = q0 + t * FA_BM + r;
      if (qrow < S) {
        const float2 lf = unpack2(l2);
        const float inv = 1.0f / (lf.x + lf.y);
        uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<long long>(frame) * S + qrow) * ldo + head * FA_D);
#pragma unroll
        for (int g = 0; g < 8; ++g)
          dst[g] = make_uint4(pack_bf16x2(__uint_as_float(o[8 * g]) * inv, __uint_as_float(o[8 * g + 1]) * inv),
                              pack_bf16x2(__uint_as_float(o[8 * g + 2]) * inv, __uint_as_float(o[8 * g + 3]) * inv),
                              pack_bf16x2(__uint_as_float(o[8 * g + 4]) * inv, __uint_as_float(o[8 * g + 5]) * inv),
                              pack_bf16x2(__uint_as_float(o[8 * g + 6]) * inv, __uint_as_float(o[8 * g + 7]) * inv));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem_base);
}

}  // namespace sa

using namespace sa;

// qkv: bf16 [frames*S, ld] with q | k | v column blocks of width C = heads*64; out: bf16 [frames*S, ldo]
extern "C" int sa_attn_spatial(const void* qkv, int32_t ld, int32_t frames, int32_t S, int32_t C,
                               int32_t heads, void* out, int32_t ldo, void* stream_) {
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
  static DeviceOnce once;
  if (int rc = once.run([&]() -> int {
        SA_CUDA_CHECK(cudaFuncSetAttribute(attn_flash_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA3_SMEM));
        return SA_OK;
      }))
    return rc;
  dim3 grid((S + 2 * FA_BM - 1) / (2 * FA_BM), heads, frames);
  attn_flash_kernel<<<grid, 320, FA3_SMEM, stream>>>(*tm, S, C, static_cast<__nv_bfloat16*>(out), ldo);
  SA_CUDA_CHECK(cudaGetLastError());
  return SA_OK;
}
