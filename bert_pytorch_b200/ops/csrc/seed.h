// Dropout seed of a launch: a host value plus an optional device-resident step counter, so that a captured
// CUDA graph draws fresh masks on every replay (the counter is advanced inside the graph).
#pragma once

namespace b200 {

struct Seed {
  unsigned long long base;
  const unsigned long long* step;     // may be null
#ifdef __CUDACC__
  __device__ __forceinline__ unsigned long long value() const {
    return step ? base + __ldg(step) * 0x9E3779B97F4A7C15ull : base;
  }
#endif
};

}  // namespace b200
