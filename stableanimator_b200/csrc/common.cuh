// Shared host-side helpers of the C-ABI library: error reporting, TMA tensor-map encode + cache.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <utility>

#include "../../include/sa_b200.h"

namespace sa {

// sets the thread-local error text returned by sa_last_error(); returns `code`
int fail(int code, const char* fmt, ...);

#define SA_CUDA_CHECK(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      return ::sa::fail(SA_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                        __FILE__, __LINE__);                                             \
  } while (0)

// bf16 tensor map of rank `rank` (<= 5), SWIZZLE_128B (or 64B), zero OOB fill, cached by content.
// dims/strides follow cuTensorMapEncodeTiled (dim 0 innermost; strides in BYTES for dims 1..).
// elem_strides (optional, per dim): traversal stride -- box[i] elements are traversed, every elem_strides[i]-th loaded.
// Returns nullptr (and sets the error text) on failure.  The cache is a bounded LRU (SA_TMAP_CACHE_CAP entries):
// a returned pointer stays valid until that many OTHER maps have been requested after it -- entry points copy the
// map into the kernel's parameter space (__grid_constant__) at launch, a handful of requests later.
const CUtensorMap* get_tmap_bf16(const void* base, int rank, const uint64_t* dims,
                                 const uint64_t* strides_bytes, const uint32_t* box,
                                 int swizzle_bytes = 128, const uint32_t* elem_strides = nullptr);

constexpr int SA_TMAP_CACHE_CAP = 8192;

// SM count of the CURRENT device (cached per device)
int num_sms();

// One-time, per-DEVICE kernel configuration (cudaFuncSetAttribute applies to the current device only), safe under
// concurrent callers: `static DeviceOnce once; if (int rc = once.run([&]() -> int { ...; return SA_OK; })) return rc;`
struct DeviceOnce {
  std::atomic<uint64_t> done{0};  // bit d = device d configured
  std::mutex mu;
  template <class F>
  int run(F&& f) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return fail(SA_ERR_CUDA, "cudaGetDevice failed");
    const uint64_t bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_acquire) & bit) return SA_OK;
    std::lock_guard<std::mutex> lock(mu);
    if (done.load(std::memory_order_relaxed) & bit) return SA_OK;
    const int rc = f();
    if (rc == SA_OK) done.fetch_or(bit, std::memory_order_release);
    return rc;
  }
};


// Programmatic dependent launch (ptx.cuh `grid_dep_sync` / `grid_dep_sync_light`).  SA_PDL = 2 (default): every kernel of the
// denoising call; 1: only the kernels with a long set-up (barrier rings, TMEM allocation, tensor-map prefetch: tapped GEMM,
// fused feed-forward, flash attention); 0: plain launches.  The heavy kernels release their successor right after their own
// set-up, the light (memory-bound) ones only when their CTAs exit -- released early they got SLOWER (+1.6 us per launch in a
// chain of LayerNorms, `tools/pdl_probe.py`); with the late release the chain gains 0.2 us per launch and the UNet call
// 0.4 ms over level 1 (0.7 ms over plain launches).
int pdl_level();

// One launch path for the kernels of the denoising call: optional cluster (cluster_x > 1) and the programmatic-launch
// attribute (`heavy` = released at level 1 already).  ONLY for kernels whose every global-memory access follows a `grid_dep_wait()`.
template <class... P, class... A>
inline cudaError_t launch_k(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, unsigned cluster_x,
                            bool heavy, A&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  unsigned n = 0;
  if (cluster_x > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = cluster_x;
    at[n].val.clusterDim.y = 1;
    at[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_level() >= (heavy ? 1 : 2)) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = at;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<A>(args)...);
}

}  // namespace sa
