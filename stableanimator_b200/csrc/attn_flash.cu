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
//   warp 8          : TMA producer (Q tiles once; K/V blocks of 128 keys, 3-stage ring shared by
//                     both tiles)
//   warp 9          : MMA issuer + TMEM allocator
// The two tiles ping-pong on the tensor core: while warpgroup A runs the softmax of S_A(j) the MMA
// pipe computes S_B(j), P_A V(j-1), ...  Per tile and key block:
//   S = Q K^T (128x128 fp32 in TMEM) -> softmax thread reads its row twice (max, then exp2), writes
//   P as bf16 into 128B-swizzled smem (A operand) -> O_part = P V (128x64 fp32 in TMEM, V consumed
//   MN-major straight from the TMA tile) -> folded into the thread's fp32 O registers with the
//   running-max correction (O = O * alpha + O_part), so TMEM is never read-modify-written.
// TMEM columns: S_A [0,128) S_B [128,256) O_A [256,320) O_B [320,384).
// smem: Q 2x16K | K 3x16K | V 3x16K | P 2x32K = 192 KB -> one CTA per SM.
#include <math.h>

#include "common.cuh"
#include "ptx.cuh"

namespace sa {

constexpr int FA_BM = 128, FA_BN = 128, FA_D = 64, FA_ST = 3;
constexpr int FA_TILE = 128 * 128;  // bytes of a [128][64] bf16 tile
constexpr int FA_SMEM = FA_TILE * (2 + 2 * FA_ST + 4) + 1024 + 256;

struct FaBars {
  uint64_t q_full;
  uint64_t k_full[FA_ST], v_full[FA_ST], kv_empty[FA_ST];
  uint64_t s_full[2], o_full[2], p_full[2];
  uint32_t tmem_slot;
};

// one key block of one query row: MASKED handles the last, partial block (keys >= kv_left are -inf)
template <bool MASKED>
__device__ __forceinline__ void fa_softmax_block(uint32_t s_addr, uint32_t o_addr, bool fold, uint64_t* o_full_bar,
                                                 uint32_t o_phase, int kv_left, float sl2, float& m, float& l,
                                                 float& alpha_prev, float (&o)[FA_D], uint8_t* prow, int sw) {
  // pass 1: row max
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < FA_BN; c += 64) {
    uint32_t u[64];
    tmem_ld_32x64(s_addr + c, u);
    tmem_ld_wait();
#pragma unroll
    for (int t = 0; t < 64; ++t) {
      if (!MASKED || c + t < kv_left) mx = fmaxf(mx, __uint_as_float(u[t]));
    }
  }
  const float m_new = fmaxf(m, mx * sl2);
  const float alpha = ex2_approx(m - m_new);  // 0 on the first block (m = -inf)
  // fold the previous block's P V into the O registers (PV_{j-1} must have completed)
  if (fold) {
    mbar_wait(o_full_bar, o_phase);
    tc_fence_after();
    uint32_t u[64];
    tmem_ld_32x64(o_addr, u);
    tmem_ld_wait();
#pragma unroll
    for (int t = 0; t < FA_D; ++t) o[t] = o[t] * alpha_prev + __uint_as_float(u[t]);
  }
  alpha_prev = alpha;
  l *= alpha;
  m = m_new;
  // pass 2: P = exp2(s * sl2 - m) -> bf16 -> swizzled smem (A operand of the PV MMA)
#pragma unroll
  for (int c = 0; c < FA_BN; c += 64) {
    uint32_t u[64];
    tmem_ld_32x64(s_addr + c, u);
    tmem_ld_wait();
    uint8_t* base = prow + (c >> 6) * FA_TILE;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      float p[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const float e = ex2_approx(fmaf(__uint_as_float(u[8 * g + t]), sl2, -m_new));
        p[t] = (!MASKED || c + 8 * g + t < kv_left) ? e : 0.f;
        l += p[t];
      }
      *reinterpret_cast<uint4*>(base + ((g ^ sw) << 4)) =
          make_uint4(pack_bf16x2(p[0], p[1]), pack_bf16x2(p[2], p[3]), pack_bf16x2(p[4], p[5]),
                     pack_bf16x2(p[6], p[7]));
    }
  }
}

__global__ void __launch_bounds__(320, 1)
attn_flash_kernel(const __grid_constant__ CUtensorMap tm, int S, int C, __nv_bfloat16* __restrict__ out,
                  int ldo) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                          // 2 tiles
  uint8_t* sK = smem + 2 * FA_TILE;            // FA_ST stages
  uint8_t* sV = sK + FA_ST * FA_TILE;          // FA_ST stages
  uint8_t* sP = sV + FA_ST * FA_TILE;          // 2 tiles x 2 halves of 64 keys
  FaBars* bars = reinterpret_cast<FaBars*>(sP + 4 * FA_TILE);

  const int warp = threadIdx.x >> 5;
  const int q0 = blockIdx.x * 2 * FA_BM;
  const int head = blockIdx.y;
  const int frame = blockIdx.z;
  const int nblk = (S + FA_BN - 1) / FA_BN;
  const int ntiles = (q0 + FA_BM < S) ? 2 : 1;

  if (warp == 8 && elect_one()) {
    tma_prefetch_desc(&tm);
    mbar_init(&bars->q_full, 1);
    for (int i = 0; i < FA_ST; ++i) {
      mbar_init(&bars->k_full[i], 1);
      mbar_init(&bars->v_full[i], 1);
      mbar_init(&bars->kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->s_full[i], 1);
      mbar_init(&bars->o_full[i], 1);
      mbar_init(&bars->p_full[i], 128);
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
        const int st = j % FA_ST;
        mbar_wait(&bars->kv_empty[st], ((j / FA_ST) & 1) ^ 1);
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
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, FA_BN, 0, 0);  // S = Q K^T (both K-major)
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, FA_D, 0, 1);   // O = P V   (V is MN-major)
      auto issue_s = [&](int t, int j) {
        const uint64_t dq = umma_desc_sw128(smem_u32(sQ + t * FA_TILE));
        const uint64_t dk = umma_desc_sw128(smem_u32(sK + (j % FA_ST) * FA_TILE));
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
        const int st = j % FA_ST;
        for (int t = 0; t < ntiles; ++t) {
          mbar_wait(&bars->p_full[t], j & 1);  // P_t(j) in smem, S_t(j) and O_t(j-1) drained
          if (t == 0) mbar_wait(&bars->v_full[st], (j / FA_ST) & 1);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint64_t dp = umma_desc_sw128(smem_u32(sP + (2 * t + (kk >> 2)) * FA_TILE)) + 2 * (kk & 3);
            const uint64_t dv = umma_desc_sw128(smem_u32(sV + st * FA_TILE + kk * 2048));
            umma_f16(tmem_base + 256 + t * 64, dp, dv, idesc_o, kk != 0 ? 1u : 0u);
          }
          umma_commit(&bars->o_full[t]);
          if (j + 1 < nblk) {
            if (t == 0) {
              mbar_wait(&bars->k_full[(j + 1) % FA_ST], ((j + 1) / FA_ST) & 1);
              tc_fence_after();
            }
            issue_s(t, j + 1);
          }
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
      const float sl2 = 0.125f * 1.4426950408889634f;  // attn.scale * log2(e)
      float o[FA_D];
#pragma unroll
      for (int d = 0; d < FA_D; ++d) o[d] = 0.f;
      float m = -INFINITY, l = 0.f, alpha_prev = 1.f;
      uint8_t* prow = sP + 2 * t * FA_TILE + r * 128;
      const int sw = r & 7;

      for (int j = 0; j < nblk; ++j) {
        const int kv_left = S - j * FA_BN;
        mbar_wait(&bars->s_full[t], j & 1);
        tc_fence_after();
        if (kv_left >= FA_BN)
          fa_softmax_block<false>(s_addr, o_addr, j >= 1, &bars->o_full[t], (j - 1) & 1, kv_left, sl2, m, l,
                                  alpha_prev, o, prow, sw);
        else
          fa_softmax_block<true>(s_addr, o_addr, j >= 1, &bars->o_full[t], (j - 1) & 1, kv_left, sl2, m, l,
                                 alpha_prev, o, prow, sw);
        fence_proxy_async();  // P visible to the tensor core's async-proxy reads
        tc_fence_before();    // our TMEM reads of S_t(j) / O_t(j-1) precede the MMAs that overwrite them
        mbar_arrive(&bars->p_full[t]);
      }
      // last block's P V
      mbar_wait(&bars->o_full[t], (nblk - 1) & 1);
      tc_fence_after();
      {
        uint32_t u[64];
        tmem_ld_32x64(o_addr, u);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < FA_D; ++d) o[d] = o[d] * alpha_prev + __uint_as_float(u[d]);
      }
      const int qrow = q0 + t * FA_BM + r;
      if (qrow < S) {
        const float inv = 1.0f / l;
        uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<long long>(frame) * S + qrow) * ldo + head * FA_D);
#pragma unroll
        for (int g = 0; g < 8; ++g)
          dst[g] = make_uint4(pack_bf16x2(o[8 * g] * inv, o[8 * g + 1] * inv),
                              pack_bf16x2(o[8 * g + 2] * inv, o[8 * g + 3] * inv),
                              pack_bf16x2(o[8 * g + 4] * inv, o[8 * g + 5] * inv),
                              pack_bf16x2(o[8 * g + 6] * inv, o[8 * g + 7] * inv));
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
  static bool configured = false;
  if (!configured) {
    SA_CUDA_CHECK(cudaFuncSetAttribute(attn_flash_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
    configured = true;
  }
  dim3 grid((S + 2 * FA_BM - 1) / (2 * FA_BM), heads, frames);
  attn_flash_kernel<<<grid, 320, FA_SMEM, stream>>>(*tm, S, C, static_cast<__nv_bfloat16*>(out), ldo);
  SA_CUDA_CHECK(cudaGetLastError());
  return SA_OK;
}
