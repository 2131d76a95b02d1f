// Spatial self-attention (head_dim 64, non-causal, no mask) as a flash-attention kernel on
// tcgen05 / TMEM with TMA-staged operands, sm_100a.
//
// Replaces xformers.ops.memory_efficient_attention(q, k, v) in AnimationAttnProcessor.__call__
// (animation/modules/unet.py:1893-1905).  Reads Q, K, V straight out of the fused projection
// output [frames*S, 3C] (column blocks q | k | v, head h at columns h*64) through ONE 3-D tensor
// map (cols, S, frames), so head_to_batch_dim / .contiguous() (:1893-1895) and batch_to_head_dim
// (:1905) never materialise; rows past S are zero-filled by TMA and masked to -inf.
//
// CTA = 128 query rows of one (frame, head); 192 threads:
//   warps 0-3 : softmax + O accumulation, one thread per query row (TMEM lane == row)
//   warp 4    : TMA producer (Q once; K/V blocks of 128 keys, 2 stages)
//   warp 5    : MMA issuer + TMEM allocator
// TMEM (256 columns): two 128-column S buffers; after the softmax threads have drained S_j the PV
// product of block j lands in the first 64 columns of the same buffer and is folded into the fp32
// O registers of the row's thread (O = O * alpha + P_j V_j), so no TMEM read-modify-write.
// smem: Q 16K | K 2x16K | V 2x16K | P 32K  (~112 KB -> two CTAs per SM).
#include <math.h>

#include "common.cuh"
#include "ptx.cuh"

namespace sa {

constexpr int FA_BM = 128, FA_BN = 128, FA_D = 64;
constexpr int FA_TILE = 128 * 128;  // bytes of a [128][64] bf16 tile
constexpr int FA_SMEM = FA_TILE * 7 + 1024 + 256;

__global__ void __launch_bounds__(192, 2)
attn_flash_kernel(const __grid_constant__ CUtensorMap tm, int S, int C, __nv_bfloat16* __restrict__ out,
                  int ldo) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sK = smem + FA_TILE;      // 2 stages
  uint8_t* sV = smem + 3 * FA_TILE;  // 2 stages
  uint8_t* sP = smem + 5 * FA_TILE;  // 2 halves of 64 keys
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 7 * FA_TILE);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;     // [2]
  uint64_t* v_full = bars + 3;     // [2]
  uint64_t* kv_empty = bars + 5;   // [2]
  uint64_t* s_full = bars + 7;     // [2]
  uint64_t* o_full = bars + 9;     // [2]
  uint64_t* p_full = bars + 11;    // 128 arrivals
  uint64_t* o_cons = bars + 12;    // 128 arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5;
  const int q0 = blockIdx.x * FA_BM;
  const int head = blockIdx.y;
  const int frame = blockIdx.z;
  const int nblk = (S + FA_BN - 1) / FA_BN;

  if (warp == 4 && elect_one()) {
    tma_prefetch_desc(&tm);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&o_full[i], 1);
    }
    mbar_init(p_full, 128);
    mbar_init(o_cons, 128);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_expect_tx(q_full, FA_TILE);
      tma_load_3d(sQ, &tm, q_full, head * FA_D, q0, frame);
      for (int j = 0; j < nblk; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&k_full[st], FA_TILE);
        tma_load_3d(sK + st * FA_TILE, &tm, &k_full[st], C + head * FA_D, j * FA_BN, frame);
        mbar_expect_tx(&v_full[st], FA_TILE);
        tma_load_3d(sV + st * FA_TILE, &tm, &v_full[st], 2 * C + head * FA_D, j * FA_BN, frame);
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, FA_BN, 0, 0);  // S = Q K^T (both K-major)
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, FA_D, 0, 1);   // O = P V   (V is MN-major)
      const uint64_t dq = umma_desc_sw128(smem_u32(sQ));
      auto issue_s = [&](int j) {
        const int st = j & 1;
        mbar_wait(&k_full[st], (j >> 1) & 1);
        tc_fence_after();
        const uint64_t dk = umma_desc_sw128(smem_u32(sK + st * FA_TILE));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem_base + st * 128, dq + 2 * k, dk + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&s_full[st]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < nblk; ++j) {
        const int st = j & 1;
        if (j + 1 < nblk) {
          if (j >= 1) mbar_wait(o_cons, (j - 1) & 1);  // O_{j-1} drained: its S buffer is free
          issue_s(j + 1);
        }
        mbar_wait(p_full, j & 1);
        mbar_wait(&v_full[st], (j >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const uint64_t dp = umma_desc_sw128(smem_u32(sP + (t >> 2) * FA_TILE)) + 2 * (t & 3);
          const uint64_t dv = umma_desc_sw128(smem_u32(sV + st * FA_TILE + t * 2048));
          umma_f16(tmem_base + st * 128, dp, dv, idesc_o, t != 0 ? 1u : 0u);
        }
        umma_commit(&o_full[st]);
        umma_commit(&kv_empty[st]);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ softmax / accumulate (row = thread)
    const int r = threadIdx.x;  // 0..127
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const float sl2 = 0.125f * 1.4426950408889634f;  // attn.scale * log2(e)
    float o[FA_D];
#pragma unroll
    for (int d = 0; d < FA_D; ++d) o[d] = 0.f;
    float m = -INFINITY, l = 0.f, alpha_prev = 1.f;
    uint8_t* prow = sP + r * 128;
    const int sw = r & 7;

    for (int j = 0; j < nblk; ++j) {
      const int st = j & 1;
      const uint32_t s_addr = lane_addr + st * 128;
      const int kv_left = S - j * FA_BN;  // valid keys in this block (>= 1)
      mbar_wait(&s_full[st], (j >> 1) & 1);
      tc_fence_after();
      // pass 1: row max
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < FA_BN; c += 32) {
        uint32_t u[32];
        tmem_ld_32x32(s_addr + c, u);
        tmem_ld_wait();
#pragma unroll
        for (int t = 0; t < 32; ++t)
          if (c + t < kv_left) mx = fmaxf(mx, __uint_as_float(u[t]));
      }
      const float m_new = fmaxf(m, mx * sl2);
      const float alpha = exp2f(m - m_new);  // 0 on the first block (m = -inf)
      // fold the previous block's P V into O (needs PV_{j-1} done; also frees sP and its S buffer)
      if (j >= 1) {
        const int pst = (j - 1) & 1;
        mbar_wait(&o_full[pst], ((j - 1) >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < FA_D; c += 32) {
          uint32_t u[32];
          tmem_ld_32x32(lane_addr + pst * 128 + c, u);
          tmem_ld_wait();
#pragma unroll
          for (int t = 0; t < 32; ++t) o[c + t] = o[c + t] * alpha_prev + __uint_as_float(u[t]);
        }
        tc_fence_before();
        mbar_arrive(o_cons);
      }
      alpha_prev = alpha;
      l *= alpha;
      m = m_new;
      // pass 2: P = exp2(s * sl2 - m) -> bf16 -> swizzled smem (A operand of the PV MMA)
#pragma unroll
      for (int c = 0; c < FA_BN; c += 32) {
        uint32_t u[32];
        tmem_ld_32x32(s_addr + c, u);
        tmem_ld_wait();
        float p[32];
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          p[t] = (c + t < kv_left) ? exp2f(__uint_as_float(u[t]) * sl2 - m_new) : 0.f;
          l += p[t];
        }
        uint8_t* base = prow + (c >> 6) * FA_TILE;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int chunk = ((c & 63) >> 3) + g;
          *reinterpret_cast<uint4*>(base + ((chunk ^ sw) << 4)) =
              make_uint4(pack_bf16x2(p[8 * g], p[8 * g + 1]), pack_bf16x2(p[8 * g + 2], p[8 * g + 3]),
                         pack_bf16x2(p[8 * g + 4], p[8 * g + 5]), pack_bf16x2(p[8 * g + 6], p[8 * g + 7]));
        }
      }
      fence_proxy_async();  // P visible to the tensor core's async-proxy reads
      tc_fence_before();    // our TMEM reads of S_j are ordered before the PV MMA overwrites it
      mbar_arrive(p_full);
    }
    // last block's P V
    {
      const int pst = (nblk - 1) & 1;
      mbar_wait(&o_full[pst], ((nblk - 1) >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < FA_D; c += 32) {
        uint32_t u[32];
        tmem_ld_32x32(lane_addr + pst * 128 + c, u);
        tmem_ld_wait();
#pragma unroll
        for (int t = 0; t < 32; ++t) o[c + t] = o[c + t] * alpha_prev + __uint_as_float(u[t]);
      }
    }
    if (q0 + r < S) {
      const float inv = 1.0f / l;
      uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<long long>(frame) * S + q0 + r) * ldo + head * FA_D);
#pragma unroll
      for (int g = 0; g < 8; ++g)
        dst[g] = make_uint4(pack_bf16x2(o[8 * g] * inv, o[8 * g + 1] * inv),
                            pack_bf16x2(o[8 * g + 2] * inv, o[8 * g + 3] * inv),
                            pack_bf16x2(o[8 * g + 4] * inv, o[8 * g + 5] * inv),
                            pack_bf16x2(o[8 * g + 6] * inv, o[8 * g + 7] * inv));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<256>(tmem_base);
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
  dim3 grid((S + FA_BM - 1) / FA_BM, heads, frames);
  attn_flash_kernel<<<grid, 192, FA_SMEM, stream>>>(*tm, S, C, static_cast<__nv_bfloat16*>(out), ldo);
  SA_CUDA_CHECK(cudaGetLastError());
  return SA_OK;
}
