// Error text, TMA tensor-map encode (driver entry point resolved at run time, no -lcuda) + cache.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <list>
#include <mutex>
#include <string>
#include <unordered_map>

#include "common.cuh"

namespace sa {

static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct TmapKey {
  uint64_t w[18];
  bool operator==(const TmapKey& o) const { return memcmp(w, o.w, sizeof(w)) == 0; }
};
struct TmapHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (uint64_t x : k.w) {
      h ^= x;
      h *= 1099511628211ull;
    }
    return static_cast<size_t>(h);
  }
};

const CUtensorMap* get_tmap_bf16(const void* base, int rank, const uint64_t* dims,
                                 const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes,
                                 const uint32_t* elem_strides) {
  // bounded LRU: `lru` front = most recently used; the map points at list nodes (key, storage)
  struct Node {
    TmapKey key;
    CUtensorMap* tm;
  };
  static std::mutex mu;
  static std::list<Node> lru;
  static std::unordered_map<TmapKey, std::list<Node>::iterator, TmapHash> cache;
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.w[0] = reinterpret_cast<uint64_t>(base);
  key.w[1] = static_cast<uint64_t>(rank) | (static_cast<uint64_t>(swizzle_bytes) << 8);
  for (int i = 0; i < rank; ++i) {
    key.w[2 + i] = dims[i];
    key.w[7 + i] = (i == 0) ? 0 : strides_bytes[i - 1];
    key.w[12 + i] = box[i];
    if (elem_strides) key.w[17] |= static_cast<uint64_t>(elem_strides[i]) << (8 * i);
  }
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) {
    lru.splice(lru.begin(), lru, it->second);
    return it->second->tm;
  }

  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    fail(SA_ERR_TMA, "cuTensorMapEncodeTiled entry point not available");
    return nullptr;
  }
  // 64-byte aligned storage as the driver requires; at capacity the least recently used entry's storage is recycled
  // (eager callers whose activation addresses churn would otherwise grow the cache without bound)
  CUtensorMap* tm = nullptr;
  if (cache.size() >= static_cast<size_t>(SA_TMAP_CACHE_CAP)) {
    tm = lru.back().tm;
    cache.erase(lru.back().key);
    lru.pop_back();
  } else if (posix_memalign(reinterpret_cast<void**>(&tm), 64, sizeof(CUtensorMap)) != 0 || !tm) {
    fail(SA_ERR_TMA, "out of host memory for a tensor map");
    return nullptr;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  if (elem_strides)
    for (int i = 0; i < rank; ++i) estr[i] = elem_strides[i];
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank),
                  const_cast<void*>(base), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fail(SA_ERR_TMA,
         "cuTensorMapEncodeTiled failed (%d): base=%p rank=%d dims=[%llu,%llu,%llu,%llu] "
         "strides=[%llu,%llu,%llu] box=[%u,%u,%u,%u]",
         static_cast<int>(r), base, rank, (unsigned long long)dims[0],
         (unsigned long long)(rank > 1 ? dims[1] : 0), (unsigned long long)(rank > 2 ? dims[2] : 0),
         (unsigned long long)(rank > 3 ? dims[3] : 0),
         (unsigned long long)(rank > 1 ? strides_bytes[0] : 0),
         (unsigned long long)(rank > 2 ? strides_bytes[1] : 0),
         (unsigned long long)(rank > 3 ? strides_bytes[2] : 0), box[0], rank > 1 ? box[1] : 0,
         rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    free(tm);
    return nullptr;
  }
  lru.push_front(Node{key, tm});
  cache.emplace(key, lru.begin());
  return tm;
}

int num_sms() {
  static std::atomic<int> cache[64];  // zero-initialised; per device
  int dev = 0;
  cudaGetDevice(&dev);
  int n = cache[dev & 63].load(std::memory_order_relaxed);
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    cache[dev & 63].store(n, std::memory_order_relaxed);
  }
  return n;
}

int pdl_level() {
  const char* e = getenv("SA_PDL");   // read per launch (captured graphs keep what they were captured with)
  return e ? (e[0] == '0' ? 0 : e[0] == '1' ? 1 : 2) : 2;
}

}  // namespace sa

extern "C" int sa_version(void) { return 100; }
extern "C" const char* sa_last_error(void) { return sa::g_err; }
