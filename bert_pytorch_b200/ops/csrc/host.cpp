// Host-side native helpers (plain C++17, no CUDA): the data pipeline's two hot loops.
//   * h5_inflate_rows -- zlib-inflate the chunks of a row-chunked HDF5 dataset straight into the output
//     array, one thread per chunk (the reference does this inside libhdf5 through h5py).
//   * mask_batch_i32  -- dynamic MLM masking of a whole micro-batch (semantics of the reference's
//     per-sample Python loop, src/dataset.py:277-296, see data/dataset.py::mask_batch), threaded over rows
//     with a counter-based RNG so results do not depend on the thread count.
//   * wp_*            -- WordPiece tokenisation (BasicTokenizer + greedy longest-match sub-words, the semantics of
//     data/tokenization.py, i.e. of the reference's src/tokenization.py:60-229; the reference gets its speed from
//     the Rust `tokenizers` package), batched and threaded; character classes / case folding come from a table the
//     Python side builds from its own rules, text outside the table stays on the fallback path.
// zlib is linked as libz.so.1 with hand-declared prototypes (the image ships no zlib.h).
#include <algorithm>
#include <atomic>
#include <climits>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
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

// ------------------------------------------------------------------------------------------------
// WordPiece.  Character handling is table driven: the Python side (data/tokenization.py) evaluates ITS OWN
// character rules (whitespace / control / punctuation classes, lower-casing + accent stripping) for every code point
// below `ncp` and hands the result over, so both implementations agree by construction; a text containing a code
// point outside the table is reported back and tokenised by the fallback.
// ------------------------------------------------------------------------------------------------
namespace {

enum : uint8_t { CP_NORMAL = 0, CP_SPACE = 1, CP_DROP = 2, CP_UNSUPPORTED = 255 };

struct WordPiece {
  std::unordered_map<std::string, int32_t> vocab;
  int32_t unk = 0;
  int max_chars = 100;
  // per code point: class, and for CP_NORMAL the replacement characters as records {flags | nbytes, bytes...}
  // (flags bit 7: the replacement character is punctuation)
  std::vector<uint8_t> cls;
  std::vector<int32_t> map_off;
  std::vector<uint8_t> map_blob;
};

// greedy longest-match-first split of one word given as UTF-8 bytes + the byte offsets of its characters
void wordpiece_word(const WordPiece& wp, const std::string& word, const std::vector<uint32_t>& bounds,
                    std::vector<int32_t>& out) {
  const size_t nchar = bounds.size() - 1;
  if ((int)nchar > wp.max_chars) { out.push_back(wp.unk); return; }
  const size_t mark = out.size();
  size_t start = 0;
  std::string sub;
  while (start < nchar) {
    size_t end = nchar;
    int32_t id = -1;
    while (start < end) {
      sub.assign(start > 0 ? "##" : "");
      sub.append(word, bounds[start], bounds[end] - bounds[start]);
      auto it = wp.vocab.find(sub);
      if (it != wp.vocab.end()) { id = it->second; break; }
      --end;
    }
    if (id < 0) { out.resize(mark); out.push_back(wp.unk); return; }
    out.push_back(id);
    start = end;
  }
}

// returns false when the text contains something the tables do not cover (caller falls back)
bool encode_text(const WordPiece& wp, const unsigned char* s, int64_t n, std::vector<int32_t>& out) {
  static const char* kNever[] = {"[UNK]", "[SEP]", "[PAD]", "[CLS]", "[MASK]"};
  const int64_t ncp = (int64_t)wp.cls.size();
  std::string raw, piece;
  std::vector<uint32_t> raw_cp, bounds;
  int64_t i = 0;
  while (i < n) {
    // ---- next whitespace-delimited token: raw bytes + its code points (control characters vanish)
    raw.clear();
    raw_cp.clear();
    while (i < n) {
      uint32_t cp;
      int len;
      const unsigned char c = s[i];
      if (c < 0x80) { cp = c; len = 1; }
      else if ((c >> 5) == 6 && i + 1 < n) { cp = ((c & 31u) << 6) | (s[i + 1] & 63u); len = 2; }
      else if ((c >> 4) == 14 && i + 2 < n) { cp = ((c & 15u) << 12) | ((s[i + 1] & 63u) << 6) | (s[i + 2] & 63u); len = 3; }
      else return false;                                   // 4-byte sequences / malformed input: not covered
      if ((int64_t)cp >= ncp) return false;
      const uint8_t k = wp.cls[cp];
      if (k == CP_UNSUPPORTED) return false;
      if (k == CP_SPACE) { i += len; if (!raw.empty()) break; else continue; }
      if (k == CP_DROP) { i += len; continue; }
      raw.append(reinterpret_cast<const char*>(s + i), (size_t)len);
      raw_cp.push_back(cp);
      i += len;
    }
    if (raw.empty()) continue;
    bool never = false;
    for (const char* k : kNever) never |= raw == k;
    if (never) {
      auto it = wp.vocab.find(raw);
      out.push_back(it != wp.vocab.end() ? it->second : wp.unk);
      continue;
    }
    // ---- replacement characters (lower-cased, accents stripped) with punctuation marks as their own words
    piece.clear();
    bounds.assign(1, 0u);
    auto flush = [&]() {
      if (!piece.empty()) { wordpiece_word(wp, piece, bounds, out); piece.clear(); bounds.assign(1, 0u); }
    };
    for (uint32_t cp : raw_cp) {
      const uint8_t* r = wp.map_blob.data() + wp.map_off[cp];
      const uint8_t* e = wp.map_blob.data() + wp.map_off[cp + 1];
      while (r < e) {
        const int nb = *r & 0x7F;
        const bool punct = (*r & 0x80) != 0;
        ++r;
        if (punct) {
          flush();
          piece.assign(reinterpret_cast<const char*>(r), (size_t)nb);
          bounds.assign({0u, (uint32_t)nb});
          flush();
        } else {
          piece.append(reinterpret_cast<const char*>(r), (size_t)nb);
          bounds.push_back((uint32_t)piece.size());
        }
        r += nb;
      }
    }
    flush();
  }
  return true;
}


// ------------------------------------------------------------------------------------------------
// byte-level BPE (GPT-2 / RoBERTa style, the `tokenizers.ByteLevelBPETokenizer` the reference offers with
// --tokenizer bpe): pre-tokenisation by the GPT-2 pattern
//   's|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+
// over a caller-supplied category table (BMP), bytes -> printable code points, then rank-ordered merges.
// ------------------------------------------------------------------------------------------------
enum : uint8_t { CAT_OTHER = 0, CAT_LETTER = 1, CAT_NUMBER = 2, CAT_SPACE = 3 };

struct Bpe {
  std::unordered_map<std::string, int32_t> vocab;   // token (in the byte-level alphabet, UTF-8) -> id
  std::unordered_map<std::string, int32_t> ranks;   // left + '\x01' + right -> merge rank
  std::vector<uint8_t> cat;                         // category per BMP code point
  std::string byte_str[256];                        // byte -> UTF-8 of its printable stand-in
};

static void append_utf8(std::string& s, uint32_t cp) {
  if (cp < 0x80) s.push_back((char)cp);
  else if (cp < 0x800) { s.push_back((char)(0xC0 | (cp >> 6))); s.push_back((char)(0x80 | (cp & 63))); }
  else { s.push_back((char)(0xE0 | (cp >> 12))); s.push_back((char)(0x80 | ((cp >> 6) & 63))); s.push_back((char)(0x80 | (cp & 63))); }
}

// merges of one pre-token given as its byte-level symbols
static bool bpe_word(const Bpe& bpe, std::vector<std::string>& sym, std::vector<int32_t>& out) {
  std::string key;
  while (sym.size() > 1) {
    int32_t best = INT32_MAX;
    for (size_t i = 0; i + 1 < sym.size(); ++i) {
      key.assign(sym[i]); key.push_back('\x01'); key.append(sym[i + 1]);
      auto it = bpe.ranks.find(key);
      if (it != bpe.ranks.end() && it->second < best) best = it->second;
    }
    if (best == INT32_MAX) break;
    std::vector<std::string> next;
    next.reserve(sym.size());
    for (size_t i = 0; i < sym.size();) {                    // merge every occurrence of the best pair, left to right
      if (i + 1 < sym.size()) {
        key.assign(sym[i]); key.push_back('\x01'); key.append(sym[i + 1]);
        auto it = bpe.ranks.find(key);
        if (it != bpe.ranks.end() && it->second == best) { next.push_back(sym[i] + sym[i + 1]); i += 2; continue; }
      }
      next.push_back(sym[i]);
      ++i;
    }
    sym.swap(next);
  }
  for (const auto& t : sym) {
    auto it = bpe.vocab.find(t);
    if (it == bpe.vocab.end()) return false;
    out.push_back(it->second);
  }
  return true;
}

static bool bpe_encode_text(const Bpe& bpe, const unsigned char* s, int64_t n, std::vector<int32_t>& out) {
  // decode to (code point, byte offset) so that the pattern can be applied on characters
  std::vector<uint32_t> cp;
  std::vector<int64_t> off;
  for (int64_t i = 0; i < n;) {
    const unsigned char c = s[i];
    uint32_t v; int len;
    if (c < 0x80) { v = c; len = 1; }
    else if ((c >> 5) == 6 && i + 1 < n) { v = ((c & 31u) << 6) | (s[i + 1] & 63u); len = 2; }
    else if ((c >> 4) == 14 && i + 2 < n) { v = ((c & 15u) << 12) | ((s[i + 1] & 63u) << 6) | (s[i + 2] & 63u); len = 3; }
    else return false;                                       // outside the BMP / malformed: caller falls back
    if (v >= bpe.cat.size()) return false;
    cp.push_back(v); off.push_back(i); i += len;
  }
  off.push_back(n);
  const size_t m = cp.size();
  auto cat = [&](size_t k) { return bpe.cat[cp[k]]; };
  std::vector<std::string> sym;
  auto emit = [&](size_t a, size_t b) -> bool {             // characters [a, b) form one pre-token
    sym.clear();
    for (int64_t k = off[a]; k < off[b]; ++k) sym.push_back(bpe.byte_str[s[k]]);
    return bpe_word(bpe, sym, out);
  };
  size_t i = 0;
  while (i < m) {
    // 's|'t|'re|'ve|'m|'ll|'d
    if (cp[i] == '\'' && i + 1 < m) {
      const uint32_t a = cp[i + 1], b = i + 2 < m ? cp[i + 2] : 0;
      size_t len = 0;
      if (a == 's' || a == 't' || a == 'm' || a == 'd') len = 2;
      else if ((a == 'r' && b == 'e') || (a == 'v' && b == 'e') || (a == 'l' && b == 'l')) len = 3;
      if (len) { if (!emit(i, i + len)) return false; i += len; continue; }
    }
    //  ?\p{L}+ |  ?\p{N}+ |  ?[^\s\p{L}\p{N}]+
    {
      const size_t j = (cp[i] == ' ' && i + 1 < m) ? i + 1 : i;
      if (j < m && cat(j) != CAT_SPACE) {
        const uint8_t kind = cat(j);
        size_t e = j;
        while (e < m && cat(e) == kind) ++e;
        if (!emit(i, e)) return false;
        i = e;
        continue;
      }
    }
    // \s+(?!\S) | \s+      (here cp[i] is white space)
    {
      size_t e = i;
      while (e < m && cat(e) == CAT_SPACE) ++e;
      if (e < m && e - i > 1) --e;                           // leave the last one for the next token's optional space
      if (!emit(i, e)) return false;
      i = e;
    }
  }
  return true;
}
}  // namespace

extern "C" {

// vocab: tokens separated by '\n', id = line number.  cls[ncp], map_off[ncp + 1], map_blob: see WordPiece.
void* wp_create(const char* blob, int64_t len, const uint8_t* cls, const int32_t* map_off, int64_t ncp,
                const uint8_t* map_blob, int64_t map_len) {
  auto* wp = new WordPiece();
  int32_t id = 0;
  int64_t a = 0;
  for (int64_t i = 0; i <= len; ++i) {
    if (i == len || blob[i] == '\n') {
      int64_t b = i;
      if (i < len || b > a) wp->vocab.emplace(std::string(blob + a, (size_t)(b - a)), id++);
      a = i + 1;
    }
  }
  auto it = wp->vocab.find("[UNK]");
  wp->unk = it != wp->vocab.end() ? it->second : 0;
  wp->cls.assign(cls, cls + ncp);
  wp->map_off.assign(map_off, map_off + ncp + 1);
  wp->map_blob.assign(map_blob, map_blob + map_len);
  return wp;
}

void wp_destroy(void* h) { delete static_cast<WordPiece*>(h); }

// texts: buf[offs[i] .. offs[i+1]) (UTF-8).  Writes the ids of text i to out[out_offs[i] .. out_offs[i+1]) and
// ok[i] = 1, or ok[i] = 0 (no ids) when the text is not covered by the tables.  Returns the total number of ids; if
// that exceeds `cap` nothing is copied and the caller retries with a larger buffer.
int64_t wp_encode_batch(void* h, const char* buf, const int64_t* offs, int64_t n, int32_t* out, int64_t cap,
                        int64_t* out_offs, uint8_t* ok, int threads) {
  const WordPiece& wp = *static_cast<const WordPiece*>(h);
  std::vector<std::vector<int32_t>> res((size_t)n);
  parallel_for(n, threads, [&](int64_t i) {
    auto& r = res[(size_t)i];
    ok[i] = encode_text(wp, reinterpret_cast<const unsigned char*>(buf) + offs[i], offs[i + 1] - offs[i], r) ? 1 : 0;
    if (!ok[i]) r.clear();
  });
  int64_t total = 0;
  for (int64_t i = 0; i < n; ++i) { out_offs[i] = total; total += (int64_t)res[(size_t)i].size(); }
  out_offs[n] = total;
  if (total > cap) return total;
  for (int64_t i = 0; i < n; ++i)
    if (!res[(size_t)i].empty()) std::memcpy(out + out_offs[i], res[(size_t)i].data(), res[(size_t)i].size() * sizeof(int32_t));
  return total;
}


// vocab_blob: tokens in id order separated by '\n' (byte-level alphabet, UTF-8); merges_blob: "left right" lines in rank
// order; cat[ncp]: CAT_* per code point.
void* bpe_create(const char* vocab_blob, int64_t vocab_len, const char* merges_blob, int64_t merges_len, const uint8_t* cat,
                 int64_t ncp) {
  auto* bpe = new Bpe();
  int32_t id = 0;
  int64_t a = 0;
  for (int64_t i = 0; i <= vocab_len; ++i)
    if (i == vocab_len || vocab_blob[i] == '\n') {
      if (i < vocab_len || i > a) bpe->vocab.emplace(std::string(vocab_blob + a, (size_t)(i - a)), id++);
      a = i + 1;
    }
  int32_t rank = 0;
  a = 0;
  for (int64_t i = 0; i <= merges_len; ++i)
    if (i == merges_len || merges_blob[i] == '\n') {
      if (i > a) {
        std::string line(merges_blob + a, (size_t)(i - a));
        const size_t sp = line.find(' ');
        if (sp != std::string::npos) {
          line[sp] = '\x01';
          bpe->ranks.emplace(line, rank++);
        }
      }
      a = i + 1;
    }
  bpe->cat.assign(cat, cat + ncp);
  // GPT-2's bytes_to_unicode: printable Latin-1 bytes map to themselves, the rest to U+0100, U+0101, ...
  uint32_t extra = 0;
  for (int b = 0; b < 256; ++b) {
    const bool keep = (b >= 0x21 && b <= 0x7E) || (b >= 0xA1 && b <= 0xAC) || (b >= 0xAE && b <= 0xFF);
    append_utf8(bpe->byte_str[b], keep ? (uint32_t)b : 256u + extra++);
  }
  return bpe;
}

void bpe_destroy(void* h) { delete static_cast<Bpe*>(h); }

// same calling convention as wp_encode_batch
int64_t bpe_encode_batch(void* h, const char* buf, const int64_t* offs, int64_t n, int32_t* out, int64_t cap,
                         int64_t* out_offs, uint8_t* ok, int threads) {
  const Bpe& bpe = *static_cast<const Bpe*>(h);
  std::vector<std::vector<int32_t>> res((size_t)n);
  parallel_for(n, threads, [&](int64_t i) {
    auto& r = res[(size_t)i];
    ok[i] = bpe_encode_text(bpe, reinterpret_cast<const unsigned char*>(buf) + offs[i], offs[i + 1] - offs[i], r) ? 1 : 0;
    if (!ok[i]) r.clear();
  });
  int64_t total = 0;
  for (int64_t i = 0; i < n; ++i) { out_offs[i] = total; total += (int64_t)res[(size_t)i].size(); }
  out_offs[n] = total;
  if (total > cap) return total;
  for (int64_t i = 0; i < n; ++i)
    if (!res[(size_t)i].empty()) std::memcpy(out + out_offs[i], res[(size_t)i].data(), res[(size_t)i].size() * sizeof(int32_t));
  return total;
}

}  // extern "C"
