// Shared host-side helpers of the C-ABI library: error reporting, TMA tensor-map encode + cache.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

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
// Returns nullptr (and sets the error text) on failure.  The returned pointer stays valid for the
// life of the process (the cache never evicts).
const CUtensorMap* get_tmap_bf16(const void* base, int rank, const uint64_t* dims,
                                 const uint64_t* strides_bytes, const uint32_t* box,
                                 int swizzle_bytes = 128, const uint32_t* elem_strides = nullptr);

int num_sms();

}  // namespace sa
