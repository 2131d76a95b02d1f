// Probe: tcgen05.ld throughput (bytes/clk) for 1, 2, 4 warps (one per TMEM lane quadrant) and 8 warps
// (two per quadrant), x32 and x64 shapes.  One CTA per SM, 148 CTAs.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../stableanimator_b200/csrc/ptx.cuh"
using namespace sa;

template <int X>
__global__ void __launch_bounds__(256, 1) k(long long* out, int iters, int nwarps) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  __syncthreads();
  if (warp < nwarps) {
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (X == 32) {
        uint32_t u[32];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tmem_ld_32x32(base + ((it & 1) * 128 + c * 32), u);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) acc ^= u[j];
        }
      } else {
        uint32_t u[64];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          tmem_ld_32x64(base + ((it & 1) * 128 + c * 64), u);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 64; ++j) acc ^= u[j];
        }
      }
    }
    t1 = clock64();
  }
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 0x12345u) out[1] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(slot);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  const int iters = 2000;
  for (int x : {32, 64})
    for (int nw : {1, 2, 4, 8}) {
      if (x == 32) k<32><<<148, 256>>>(d, iters, nw); else k<64><<<148, 256>>>(d, iters, nw);
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      double bytes = double(iters) * 128 * 32 * 4 * nw;  // per CTA: each warp reads 32 lanes x 128 cols x 4 B per iter
      printf("x%d warps=%d: %lld clk, %.1f B/clk/SM (%.1f B/clk/warp)\n", x, nw, h[0], bytes / h[0], bytes / h[0] / nw);
    }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
