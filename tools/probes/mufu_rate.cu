// Probe: per-SM throughput of ex2.approx in f32 / f16x2 / bf16x2 form, and of cvt/pack helpers.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/_bin/mufu_rate tools/probes/mufu_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(uint32_t* out, int iters, float seed) {
  // 8 independent chains per thread
  float f[8];
  uint32_t h[8];
  for (int i = 0; i < 8; ++i) { f[i] = seed + i * 0.01f + threadIdx.x * 1e-4f; h[i] = 0x3c003c00u + i + threadIdx.x; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 3) asm volatile("ex2.approx.f16 %0, %0;" : "+h"(*reinterpret_cast<uint16_t*>(&h[i])));
      if (MODE == 4) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(f[i]));
      if (MODE == 5) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %1;" : "+r"(h[i]) : "f"(f[i]));
      if (MODE == 6) asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(seed));
      if (MODE == 7) asm volatile("add.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(seed));
      if (MODE == 8) asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(*reinterpret_cast<uint64_t*>(&f[i & 6])));
    }
  }
  uint32_t acc = 0;
  for (int i = 0; i < 8; ++i) acc ^= h[i] ^ __float_as_uint(f[i]);
  if (acc == 0x12345678u) out[0] = acc;
}

template <int MODE>
void run(const char* name, double ops_per_instr) {
  uint32_t* d; cudaMalloc(&d, 4);
  const int iters = 4096, threads = 512, blocks = 148 * 2;
  k<MODE><<<blocks, threads>>>(d, 16, 0.5f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(d, iters, 0.5f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double instr = double(blocks) * threads * iters * 8;
  double per_sm_per_clk = instr / 148.0 / (ms * 1e-3 * clk_khz * 1e3);
  printf("%-28s %8.3f ms  %7.2f thread-instr/clk/SM (at max clock %d MHz)  -> %7.2f elem/clk/SM\n", name, ms,
         per_sm_per_clk, clk_khz / 1000, per_sm_per_clk * ops_per_instr);
  cudaFree(d);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.f16x2", 2);
  run<2>("ex2.approx.ftz.bf16x2", 2);
  run<3>("ex2.approx.f16", 1);
  run<4>("fma.rn.f32", 1);
  run<5>("cvt.rn.bf16x2.f32", 2);
  run<6>("max.f32", 1);
  run<7>("add.f32", 1);
  run<8>("fma.rn.f32x2", 2);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
