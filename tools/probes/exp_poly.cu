// Probe: softmax exp loop with a fraction of the exponentials evaluated on the FMA pipe (round-to-nearest
// Cody-Waite split + degree-3 polynomial + exponent insertion) instead of MUFU.EX2.  8 warps per CTA
// (2 per SM sub-partition), 148 CTAs.  Prints cycles per 128-element row-block per warp and the max
// relative error of the polynomial path against exp2f.
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>
#include "../../stableanimator_b200/csrc/ptx.cuh"
using namespace sa;

__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// 2^x for a packed pair, x clamped to >= -125: t = x + 1.5*2^23 (integer part in the low mantissa bits),
// f = x - (t - magic) in [-0.5, 0.5], p(f) ~ 2^f (rel err 7.5e-5), result = p * 2^int via exponent add.
__device__ __forceinline__ void exp2_poly2(uint64_t x, float& e0, float& e1) {
  const uint64_t magic = pack2(12582912.f, 12582912.f);
  const float2 xc = unpack2(x);
  const uint64_t xx = pack2(fmaxf(xc.x, -125.f), fmaxf(xc.y, -125.f));
  const uint64_t t = add2(xx, magic);
  const uint64_t f = sub2(xx, sub2(t, magic));
  uint64_t p = fma2(pack2(0.05517090f, 0.05517090f), f, pack2(0.24260951f, 0.24260951f));
  p = fma2(p, f, pack2(0.69326096f, 0.69326096f));
  p = fma2(p, f, pack2(0.99992817f, 0.99992817f));
  const float2 pf = unpack2(p), tf = unpack2(t);
  e0 = __uint_as_float(__float_as_uint(pf.x) + (__float_as_uint(tf.x) << 23));
  e1 = __uint_as_float(__float_as_uint(pf.y) + (__float_as_uint(tf.y) << 23));
}

// POLY = number of pairs (out of 4 per 8 elements) evaluated with the polynomial
template <int POLY>
__global__ void __launch_bounds__(256, 1) k(uint32_t* out, long long* clk, int iters, float seed) {
  uint32_t s[128];
  for (int i = 0; i < 128; ++i) s[i] = __float_as_uint(-seed * (i + 1) * 0.05f - threadIdx.x * 1e-3f);
  const uint64_t sl22 = pack2(0.18f, 0.18f);
  uint64_t nm2 = pack2(-seed, -seed);
  uint64_t la = pack2(0.f, 0.f), lb = pack2(0.f, 0.f);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 128; c += 8) {
      float e[8];
#pragma unroll
      for (int pr = 0; pr < 4; ++pr) {
        const uint64_t x = fma2(pack2u(s[c + 2 * pr], s[c + 2 * pr + 1]), sl22, nm2);
        if (pr < POLY) {
          exp2_poly2(x, e[2 * pr], e[2 * pr + 1]);
        } else {
          const float2 xf = unpack2(x);
          e[2 * pr] = ex2_approx(xf.x);
          e[2 * pr + 1] = ex2_approx(xf.y);
        }
      }
      la = add2(la, add2(pack2(e[0], e[1]), pack2(e[2], e[3])));
      lb = add2(lb, add2(pack2(e[4], e[5]), pack2(e[6], e[7])));
      acc ^= pack_f16x2(e[0], e[1]) ^ pack_f16x2(e[2], e[3]) ^ pack_f16x2(e[4], e[5]) ^ pack_f16x2(e[6], e[7]);
    }
    nm2 = add2(nm2, pack2(1e-3f, 1e-3f));
  }
  const long long t1 = clock64();
  const float2 l = unpack2(add2(la, lb));
  if ((acc ^ __float_as_uint(l.x + l.y)) == 0x1234567u) out[0] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

__global__ void err_kernel(float* maxerr) {
  float worst = 0.f;
  for (int i = threadIdx.x; i < 4000000; i += blockDim.x) {
    const float x = -130.f + i * (138.f / 4000000.f);
    float e0, e1;
    exp2_poly2(pack2(x, x), e0, e1);
    const float ref = exp2f(fmaxf(x, -125.f));
    worst = fmaxf(worst, fabsf(e0 - ref) / ref);
  }
  atomicMax(reinterpret_cast<int*>(maxerr), __float_as_int(worst));
}

template <int POLY>
void run() {
  uint32_t* d; long long* c; cudaMalloc(&d, 4); cudaMalloc(&c, 8);
  const int iters = 200;
  k<POLY><<<148, 256>>>(d, c, iters, 0.5f);
  long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("poly pairs per 4: %d (%3d %% of the exponentials on the FMA pipe): %7.1f clk per 128-element row-block per warp\n",
         POLY, POLY * 25, double(h) / iters);
  cudaFree(d); cudaFree(c);
}

int main() {
  run<0>(); run<1>(); run<2>(); run<3>();
  float* e; cudaMalloc(&e, 4); cudaMemset(e, 0, 4);
  err_kernel<<<1, 256>>>(e);
  float h; cudaMemcpy(&h, e, 4, cudaMemcpyDeviceToHost);
  printf("polynomial exp2 max relative error on [-130, 8]: %.3e\n", h);
  cudaError_t st = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(st));
  return st != cudaSuccess;
}
