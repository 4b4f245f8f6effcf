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

// Optional fp8 copy of a kernel's bf16 output (delayed scaling: quantise with meta[1], record |x| max in meta[0]),
// so that the tensor does not have to be read again by a separate quantise pass.
struct Fp8Out {
  unsigned char* q = nullptr;         // nullptr: disabled
  float* meta = nullptr;              // {amax, scale, inv_scale, _}
  int e5m2 = 0;
};

}  // namespace b200
