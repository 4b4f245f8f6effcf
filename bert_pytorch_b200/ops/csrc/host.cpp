// Host-side native helpers (plain C++17, no CUDA): the data pipeline's two hot loops.
//   * h5_inflate_rows -- zlib-inflate the chunks of a row-chunked HDF5 dataset straight into the output
//     array, one thread per chunk (the reference does this inside libhdf5 through h5py).
//   * mask_batch_i32  -- dynamic MLM masking of a whole micro-batch (semantics of the reference's
//     per-sample Python loop, src/dataset.py:277-296, see data/dataset.py::mask_batch), threaded over rows
//     with a counter-based RNG so results do not depend on the thread count.
// zlib is linked as libz.so.1 with hand-declared prototypes (the image ships no zlib.h).
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

extern "C" {
int uncompress(unsigned char* dest, unsigned long* destLen, const unsigned char* source, unsigned long sourceLen);
}

namespace {

inline int pick_threads(int requested, int64_t work) {
  int hw = (int)std::thread::hardware_concurrency();
  if (hw <= 0) hw = 4;
  int t = requested > 0 ? requested : std::min(hw, 16);
  return (int)std::max<int64_t>(1, std::min<int64_t>(t, work));
}

template <typename F>
void parallel_for(int64_t n, int threads, F&& fn) {
  threads = pick_threads(threads, n);
  if (threads <= 1) {
    for (int64_t i = 0; i < n; ++i) fn(i);
    return;
  }
  std::atomic<int64_t> next{0};
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([&] {
      for (;;) {
        int64_t i = next.fetch_add(1);
        if (i >= n) break;
        fn(i);
      }
    });
  for (auto& th : pool) th.join();
}

// splitmix64 -> stateless counter RNG
inline uint64_t mix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
inline double u01(uint64_t seed, uint64_t row, uint64_t draw, uint64_t which) {
  return (double)(mix(seed ^ mix(row * 0x100000001B3ull + draw * 8 + which)) >> 11) * (1.0 / 9007199254740992.0);
}

}  // namespace

extern "C" {

int h5_inflate_rows(const unsigned char* file, int64_t file_len, const int64_t* offs, const int64_t* sizes,
                    const int64_t* rows0, int64_t nchunks, int64_t chunk_rows, int64_t total_rows, int64_t row_bytes,
                    unsigned char* out, int threads) {
  std::atomic<int> rc{0};
  parallel_for(nchunks, threads, [&](int64_t c) {
    if (offs[c] < 0 || offs[c] + sizes[c] > file_len) { rc = -100; return; }
    const int64_t r0 = rows0[c];
    const int64_t rows = std::min(chunk_rows, total_rows - r0);
    unsigned long want = (unsigned long)(chunk_rows * row_bytes);
    if (rows == chunk_rows) {
      unsigned long got = want;
      int z = uncompress(out + r0 * row_bytes, &got, file + offs[c], (unsigned long)sizes[c]);
      if (z != 0 || got != want) rc = z != 0 ? z : -101;
    } else {  // edge chunk is stored full size: inflate to scratch, copy the live rows
      std::vector<unsigned char> tmp(want);
      unsigned long got = want;
      int z = uncompress(tmp.data(), &got, file + offs[c], (unsigned long)sizes[c]);
      if (z != 0 || got != want) { rc = z != 0 ? z : -101; return; }
      if (rows > 0) std::memcpy(out + r0 * row_bytes, tmp.data(), (size_t)(rows * row_bytes));
    }
  });
  return rc.load();
}

void mask_batch_i32(const int32_t* ids, const int32_t* sp, int32_t* out_ids, int32_t* labels, int64_t B, int64_t S,
                    int64_t nsp, int32_t mask_token, int32_t max_pred, double p_mask, int32_t vocab, double p_keep,
                    double p_rand, uint64_t seed, int threads) {
  parallel_for(B, threads, [&](int64_t b) {
    const int32_t* in = ids + b * S;
    int32_t* o = out_ids + b * S;
    int32_t* lab = labels + b * S;
    std::memcpy(o, in, sizeof(int32_t) * S);
    std::fill(lab, lab + S, -1);
    int32_t special[8];
    const int ns = (int)std::min<int64_t>(nsp, 8);
    for (int i = 0; i < ns; ++i) special[i] = sp[b * nsp + i];
    const int32_t last = special[ns - 1];
    std::sort(special, special + ns - 1);
    const int64_t ncand = std::max<int64_t>((int64_t)last - (ns - 1), 0);
    if (ncand <= 0) return;
    int64_t count = std::min<int64_t>(max_pred, std::max<int64_t>(1, (int64_t)((double)ncand * p_mask)));
    for (int64_t j = 0; j < count; ++j) {
      int64_t k = (int64_t)(u01(seed, (uint64_t)b, (uint64_t)j, 0) * (double)ncand);
      if (k >= ncand) k = ncand - 1;
      int64_t pos = k;
      for (int s = 0; s < ns - 1; ++s) pos += (pos >= special[s]) ? 1 : 0;
      lab[pos] = in[pos];
      const double a = u01(seed, (uint64_t)b, (uint64_t)j, 1);
      if (a < p_keep) continue;
      if (a < p_keep + p_rand) {
        const int64_t hi = std::max<int64_t>((int64_t)vocab - 1, 1);
        o[pos] = (int32_t)std::min<int64_t>((int64_t)(u01(seed, (uint64_t)b, (uint64_t)j, 2) * (double)hi), hi - 1);
      } else {
        o[pos] = mask_token;
      }
    }
  });
}

}  // extern "C"
