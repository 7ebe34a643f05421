// oracle/oracle.cpp -- CPU oracle: a restatement of SereneDB's query-time hot path.
//
// TEST INFRASTRUCTURE, NOT PRODUCT (see oracle.h). Compiled with -ffp-contract=off so that the
// fp32 BM25 expression is evaluated with exactly the operation order of the reference source.
// Every section cites the reference file:line it follows (paths relative to /root/reference,
// "irs/" = libs/iresearch/include/iresearch/).
#include "oracle.h"

#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cassert>
#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <mutex>
#include <memory>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

constexpr uint32_t kBlock = 128;     // irs/utils/type_limits.hpp:49 doc_limits::kBlockSize
constexpr uint32_t kSkipN = 32;      // irs/utils/type_limits.hpp:50 doc_limits::kSkipSize
constexpr uint32_t kMaxLevels = 5;   // irs/utils/type_limits.hpp:51 doc_limits::kMaxSkipLevels

// ------------------------------------------------------------------------------------------
// byte sink / vint  (irs/utils/bytes_utils.hpp:93-175: LEB128, 7 bits per byte, LSB group first)
// ------------------------------------------------------------------------------------------
struct Out {
  std::vector<uint8_t> b;
  size_t pos() const { return b.size(); }
  void byte(uint8_t v) { b.push_back(v); }
  void u16(uint16_t v) { b.push_back(uint8_t(v)); b.push_back(uint8_t(v >> 8)); }
  void u32(uint32_t v) { for (int i = 0; i < 4; ++i) b.push_back(uint8_t(v >> (8 * i))); }
  void data(const void* p, size_t n) {
    const auto* c = static_cast<const uint8_t*>(p);
    b.insert(b.end(), c, c + n);
  }
  void v32(uint32_t v) { while (v >= 0x80) { b.push_back(uint8_t(v) | 0x80); v >>= 7; } b.push_back(uint8_t(v)); }
  void v64(uint64_t v) { while (v >= 0x80) { b.push_back(uint8_t(v) | 0x80); v >>= 7; } b.push_back(uint8_t(v)); }
  void clear() { b.clear(); }
};
inline uint32_t vsize32(uint32_t v) { uint32_t n = 1; while (v >= 0x80) { v >>= 7; ++n; } return n; }
inline uint32_t rv32(const uint8_t*& p) {
  uint32_t v = 0; int s = 0; uint8_t c;
  do { c = *p++; v |= uint32_t(c & 0x7F) << s; s += 7; } while (c & 0x80);
  return v;
}
inline uint64_t rv64(const uint8_t*& p) {
  uint64_t v = 0; int s = 0; uint8_t c;
  do { c = *p++; v |= uint64_t(c & 0x7F) << s; s += 7; } while (c & 0x80);
  return v;
}
inline uint16_t ru16(const uint8_t* p) { return uint16_t(p[0] | (p[1] << 8)); }
inline uint32_t ru32(const uint8_t* p) { uint32_t v; std::memcpy(&v, p, 4); return v; }

// ------------------------------------------------------------------------------------------
// simdcomp bit layout (third_party/simdcomp/src/simdbitpacking.c, simdintegratedbitpacking.c:7-18):
// 128 values = 32 rows of one __m128i; value i -> lane l=i&3, row r=i>>2; each lane is an
// LSB-first stream of 32 b-bit fields held in b consecutive 32-bit words; word w of the four
// lanes forms output vector w. "d1" packs v[i]-v[i-1] (v[-1]=init) and decodes by prefix sum.
// ------------------------------------------------------------------------------------------
void pack128(const uint32_t* in, uint32_t* out, uint32_t bits) {
  std::memset(out, 0, 16 * bits);
  for (uint32_t i = 0; i < 128; ++i) {
    const uint32_t l = i & 3, r = i >> 2, p = r * bits, w = p >> 5, s = p & 31;
    out[4 * w + l] |= in[i] << s;
    if (s + bits > 32) out[4 * (w + 1) + l] |= in[i] >> (32 - s);
  }
}
void unpack128_scalar(const uint32_t* in, uint32_t* out, uint32_t bits) {
  const uint32_t mask = bits >= 32 ? 0xFFFFFFFFu : ((1u << bits) - 1u);
  for (uint32_t i = 0; i < 128; ++i) {
    const uint32_t l = i & 3, r = i >> 2, p = r * bits, w = p >> 5, s = p & 31;
    uint32_t v = in[4 * w + l] >> s;
    if (s + bits > 32) v |= in[4 * (w + 1) + l] << (32 - s);
    out[i] = v & mask;
  }
}
void pack128_d1(uint32_t prev, const uint32_t* in, uint32_t* out, uint32_t bits) {
  uint32_t d[128];
  for (uint32_t i = 0; i < 128; ++i) { d[i] = in[i] - prev; prev = in[i]; }
  pack128(d, out, bits);
}
void unpack128_d1_scalar(uint32_t prev, const uint32_t* in, uint32_t* out, uint32_t bits) {
  unpack128_scalar(in, out, bits);
  for (uint32_t i = 0; i < 128; ++i) { prev += out[i]; out[i] = prev; }
}

// Optional indirection to the reference's own simdcomp (oracle/_ref/libsimdcomp_ref.so).
using unpack_fn = void (*)(const void*, uint32_t*, uint32_t);
using unpackd1_fn = void (*)(uint32_t, const void*, uint32_t*, uint32_t);
unpack_fn g_ref_unpack = nullptr;
unpackd1_fn g_ref_unpackd1 = nullptr;

inline void unpack128(const uint8_t* in, uint32_t* out, uint32_t bits) {
  if (g_ref_unpack) { g_ref_unpack(in, out, bits); return; }
  uint32_t w[128]; std::memcpy(w, in, 16 * bits);
  unpack128_scalar(w, out, bits);
}
inline void unpack128_d1(uint32_t prev, const uint8_t* in, uint32_t* out, uint32_t bits) {
  if (g_ref_unpackd1) { g_ref_unpackd1(prev, in, out, bits); return; }
  uint32_t w[128]; std::memcpy(w, in, 16 * bits);
  unpack128_d1_scalar(prev, w, out, bits);
}

// ------------------------------------------------------------------------------------------
// StreamVByte "1234" (github.com/serenedb/streamvbyte, fork of lemire/streamvbyte; submodule is
// empty in /root/reference => restated from the public format):
// ceil(len/4) key bytes (2 bits per value, LSB-first, code = byte length-1) then the data bytes.
// ------------------------------------------------------------------------------------------
inline uint32_t bytes1234(uint32_t v) { return v < (1u << 8) ? 1 : v < (1u << 16) ? 2 : v < (1u << 24) ? 3 : 4; }
size_t svb_encode(const uint32_t* in, uint32_t len, uint8_t* out, bool delta, uint32_t prev) {
  const uint32_t nkeys = (len + 3) / 4;
  uint8_t* keys = out;
  uint8_t* data = out + nkeys;
  std::memset(keys, 0, nkeys);
  for (uint32_t i = 0; i < len; ++i) {
    uint32_t v = in[i];
    if (delta) { v = in[i] - prev; prev = in[i]; }
    const uint32_t n = bytes1234(v);
    keys[i >> 2] |= uint8_t((n - 1) << (2 * (i & 3)));
    for (uint32_t j = 0; j < n; ++j) *data++ = uint8_t(v >> (8 * j));
  }
  return size_t(data - out);
}
size_t svb_decode(const uint8_t* in, uint32_t* out, uint32_t len, bool delta, uint32_t prev) {
  const uint32_t nkeys = (len + 3) / 4;
  const uint8_t* keys = in;
  const uint8_t* data = in + nkeys;
  for (uint32_t i = 0; i < len; ++i) {
    const uint32_t n = ((keys[i >> 2] >> (2 * (i & 3))) & 3) + 1;
    uint32_t v = 0;
    for (uint32_t j = 0; j < n; ++j) v |= uint32_t(*data++) << (8 * j);
    if (delta) { prev += v; v = prev; }
    out[i] = v;
  }
  return size_t(data - in);
}

// ------------------------------------------------------------------------------------------
// 128-doc block codec. irs/formats/posting/format_block_128.hpp
// ------------------------------------------------------------------------------------------
enum DeltaEncoding : uint8_t {  // :652-713
  de_values = 0, de_delta_all_same_08, de_delta_all_same_16, de_delta_all_same_32, de_for_bitset,
  de_streamvbyte1234, de_for_streamvbyte1234, de_delta_streamvbyte1234, de_delta_bitpack_02 /* .. _31 */
};
enum Encoding : uint8_t {  // :722-770
  e_values = 0, e_all_same_08, e_all_same_16, e_all_same_32, e_streamvbyte1234, e_bitpack_01 /* .. _31 */
};
inline uint32_t bytes0124(uint32_t v) { return v == 0 ? 0 : v < (1u << 8) ? 1 : v < (1u << 16) ? 2 : 4; }  // :797-808
inline uint32_t bit_width32(uint32_t v) { return v ? 32 - __builtin_clz(v) : 0; }

// WriteTailDelta, :57-242. Candidate order and the strict '<' comparisons are the reference's.
void write_doc_block(Out& out, const uint32_t* in, uint32_t len, uint32_t prev) {
  uint8_t best = de_values;
  uint32_t best_size = len * 4;
  bool all_same = true;
  const uint32_t max = in[len - 1];
  const uint32_t for_max = max - prev;
  uint32_t delta_prev = prev;
  uint32_t delta_max = in[0] - delta_prev;
  const uint32_t groups = (len + 3) / 4;
  uint32_t size_svb = 2 + groups, size_dsvb = 2 + groups;
  for (uint32_t i = 0; i < len; ++i) {
    const uint32_t value = in[i];
    const uint32_t dv = value - delta_prev;
    delta_prev = value;
    all_same &= (delta_max == dv);
    delta_max = std::max(delta_max, dv);
    size_svb += bytes1234(value);
    size_dsvb += bytes1234(dv);
  }
  bool decided = false;
  if (all_same) {  // :103-120 (early return: bitset is not even considered)
    switch (bytes0124(delta_max)) {
      case 1: best = de_delta_all_same_08; best_size = 1; break;
      case 2: best = de_delta_all_same_16; best_size = 2; break;
      default: best = de_delta_all_same_32; best_size = 4; break;
    }
    decided = true;
  }
  if (!decided) {
    if (len == kBlock) {  // SupportIfBlock :775-777
      const uint32_t bits = bit_width32(delta_max);
      const uint32_t size = (kBlock * bits + 7) / 8;
      if (size < best_size) { best = uint8_t(de_delta_bitpack_02 + (bits - 2)); best_size = size; }
    }
    if (len != kBlock && size_svb < best_size) { best = de_streamvbyte1234; best_size = size_svb; }      // :133
    if (len != kBlock && size_dsvb < best_size) { best = de_delta_streamvbyte1234; best_size = size_dsvb; }  // :141
    {
      const uint32_t size = 1 + ((for_max + 1 + 63) / 64) * 8;  // :147-148
      if (size - 2 < best_size) { best = de_for_bitset; best_size = size; }
    }
  }
  out.byte(best);
  switch (best) {
    case de_values: out.data(in, best_size); break;
    case de_delta_all_same_08: out.byte(uint8_t(delta_max)); break;
    case de_delta_all_same_16: out.u16(uint16_t(delta_max)); break;
    case de_delta_all_same_32: out.u32(delta_max); break;
    case de_for_bitset: {  // WriteBitset :815-839
      const uint32_t bytes = best_size - 1, words = bytes / 8;
      std::vector<uint64_t> bs(words, 0);
      for (uint32_t i = 0; i < len; ++i) { const uint32_t v = in[i] - prev; bs[v / 64] |= uint64_t(1) << (v % 64); }
      out.byte(uint8_t(words));
      out.data(bs.data(), bytes);
    } break;
    case de_streamvbyte1234: {
      uint8_t buf[4 * kBlock + 64];
      const size_t n = svb_encode(in, len, buf, false, 0);
      out.u16(uint16_t(n)); out.data(buf, n);
    } break;
    case de_delta_streamvbyte1234: {
      uint8_t buf[4 * kBlock + 64];
      const size_t n = svb_encode(in, len, buf, true, prev);
      out.u16(uint16_t(n)); out.data(buf, n);
    } break;
    default: {  // de_delta_bitpack_b: full blocks only, so MakeBlockFromTail (:841-848) is a no-op
      const uint32_t bits = uint32_t(best - de_delta_bitpack_02) + 2;
      uint32_t w[128];
      pack128_d1(prev, in, w, bits);
      out.data(w, best_size);
    }
  }
}

// ReadTailDelta :475-559. Values are returned left-aligned (out[0..len)); the reference right-aligns
// them in a 128-slot buffer (:482) which is a CPU buffer-management detail.
size_t read_doc_block(const uint8_t* in, uint32_t len, uint32_t prev, uint32_t* out) {
  const uint8_t* p = in;
  const uint8_t type = *p++;
  switch (type) {
    case de_values: std::memcpy(out, p, len * 4); p += len * 4; break;
    case de_delta_all_same_08: { const uint32_t d = *p++; for (uint32_t i = 0; i < len; ++i) out[i] = prev + d + d * i; } break;  // FillSameDelta :953-959
    case de_delta_all_same_16: { const uint32_t d = ru16(p); p += 2; for (uint32_t i = 0; i < len; ++i) out[i] = prev + d + d * i; } break;
    case de_delta_all_same_32: { const uint32_t d = ru32(p); p += 4; for (uint32_t i = 0; i < len; ++i) out[i] = prev + d + d * i; } break;
    case de_for_bitset: {  // MaterializeBitset :402-445
      const uint32_t words = *p++;
      uint32_t n = 0;
      for (uint32_t i = 0; i < words; ++i) {
        uint64_t word; std::memcpy(&word, p + 8 * i, 8);
        while (word) { out[n++] = prev + i * 64 + uint32_t(__builtin_ctzll(word)); word &= word - 1; }
      }
      assert(n == len);
      p += 8 * words;
    } break;
    case de_streamvbyte1234: { const uint32_t size = ru16(p); p += 2; svb_decode(p, out, len, false, 0); p += size; } break;
    case de_delta_streamvbyte1234: { const uint32_t size = ru16(p); p += 2; svb_decode(p, out, len, true, prev); p += size; } break;
    default: {
      const uint32_t bits = uint32_t(type - de_delta_bitpack_02) + 2;
      assert(len == kBlock && bits <= 31);
      unpack128_d1(prev, p, out, bits);
      p += 16 * bits;
    }
  }
  return size_t(p - in);
}

// WriteTail :249-379
void write_freq_block(Out& out, const uint32_t* in, uint32_t len) {
  uint8_t best = e_values;
  uint32_t best_size = len * 4;
  bool all_same = true;
  uint32_t max = in[0];
  uint32_t size_svb = 2 + (len + 3) / 4;
  for (uint32_t i = 0; i < len; ++i) {
    all_same &= (max == in[i]);
    max = std::max(max, in[i]);
    size_svb += bytes1234(in[i]);
  }
  if (all_same) {
    switch (bytes0124(max)) {
      case 0: case 1: best = e_all_same_08; best_size = 1; break;
      case 2: best = e_all_same_16; best_size = 2; break;
      default: best = e_all_same_32; best_size = 4; break;
    }
  } else {
    if (len == kBlock) {
      const uint32_t bits = bit_width32(max);
      const uint32_t size = (kBlock * bits + 7) / 8;
      if (size < best_size) { best = uint8_t(e_bitpack_01 + (bits - 1)); best_size = size; }
    }
    if (len != kBlock && size_svb < best_size) { best = e_streamvbyte1234; best_size = size_svb; }
  }
  out.byte(best);
  switch (best) {
    case e_values: out.data(in, best_size); break;
    case e_all_same_08: out.byte(uint8_t(max)); break;
    case e_all_same_16: out.u16(uint16_t(max)); break;
    case e_all_same_32: out.u32(max); break;
    case e_streamvbyte1234: {
      uint8_t buf[4 * kBlock + 64];
      const size_t n = svb_encode(in, len, buf, false, 0);
      out.u16(uint16_t(n)); out.data(buf, n);
    } break;
    default: {
      const uint32_t bits = uint32_t(best - e_bitpack_01) + 1;
      uint32_t w[128];
      pack128(in, w, bits);
      out.data(w, best_size);
    }
  }
}

// ReadTail :567-636
size_t read_freq_block(const uint8_t* in, uint32_t len, uint32_t* out) {
  const uint8_t* p = in;
  const uint8_t type = *p++;
  switch (type) {
    case e_values: std::memcpy(out, p, len * 4); p += len * 4; break;
    case e_all_same_08: { const uint32_t v = *p++; std::fill_n(out, len, v); } break;
    case e_all_same_16: { const uint32_t v = ru16(p); p += 2; std::fill_n(out, len, v); } break;
    case e_all_same_32: { const uint32_t v = ru32(p); p += 4; std::fill_n(out, len, v); } break;
    case e_streamvbyte1234: { const uint32_t size = ru16(p); p += 2; svb_decode(p, out, len, false, 0); p += size; } break;
    default: {
      const uint32_t bits = uint32_t(type - e_bitpack_01) + 1;
      assert(len == kBlock && bits <= 31);
      unpack128(p, out, bits);
      p += 16 * bits;
    }
  }
  return size_t(p - in);
}

// Size of an encoded block without decoding it (SkipTail :643-648 / SizeDelta :851-895 / Size :898-951).
size_t doc_block_size(const uint8_t* in, uint32_t len) {
  const uint8_t type = in[0];
  switch (type) {
    case de_values: return 1 + len * 4;
    case de_delta_all_same_08: return 2;
    case de_delta_all_same_16: return 3;
    case de_delta_all_same_32: return 5;
    case de_for_bitset: return 2 + 8 * size_t(in[1]);
    case de_streamvbyte1234: case de_delta_streamvbyte1234: return 3 + ru16(in + 1);
    default: return 1 + 16 * (size_t(type - de_delta_bitpack_02) + 2);
  }
}
size_t freq_block_size(const uint8_t* in, uint32_t len) {
  const uint8_t type = in[0];
  switch (type) {
    case e_values: return 1 + len * 4;
    case e_all_same_08: return 2;
    case e_all_same_16: return 3;
    case e_all_same_32: return 5;
    case e_streamvbyte1234: return 3 + ru16(in + 1);
    default: return 1 + 16 * (size_t(type - e_bitpack_01) + 1);
  }
}

// ------------------------------------------------------------------------------------------
// Block-max producer for BM25 with known avg_dl (kWandTagAvgDL; bm25.cpp:386-392 default
// "approximate"), irs/formats/posting/wand_writer.hpp.
// ------------------------------------------------------------------------------------------
struct WandEntry { uint32_t freq = 1; uint32_t norm = std::numeric_limits<uint32_t>::max(); };  // :178-182

// CmpBm25 :142-175 -- returns true iff (tf_1,dl_1) is strictly better than (tf_2,dl_2).
inline bool bm25_better(float avg_dl, float b, uint32_t tf_1, uint32_t dl_1, uint32_t tf_2, uint32_t dl_2) {
  const float x = (1.f - b) * avg_dl;
  const float lhs = float(tf_1) * (x + b * float(dl_2));
  const float rhs = float(tf_2) * (x + b * float(dl_1));
  return !(lhs <= rhs);  // ProduceBM25 :302-311: "if (cmp <= 0) return;" (unordered falls through)
}
inline void wand_produce(float avg_dl, float b, uint32_t freq, uint32_t norm, WandEntry& to) {
  if (bm25_better(avg_dl, b, freq, norm, to.freq, to.norm)) { to.freq = freq; to.norm = norm; }
}
inline uint8_t wand_size(const WandEntry& e) {  // :208-218
  uint32_t s = vsize32(e.freq);
  if (e.norm != e.freq) s += vsize32(e.norm - e.freq);
  return uint8_t(s);
}
inline void wand_write(const WandEntry& e, Out& out) {  // :196-206
  out.v32(e.freq);
  if (e.norm != e.freq) out.v32(e.norm - e.freq);
}
inline WandEntry wand_read(const uint8_t*& p, uint32_t size) {  // FreqNormSource::Read :366-381
  WandEntry e;
  const uint8_t* start = p;
  e.freq = rv32(p);
  e.norm = e.freq;
  if (uint32_t(p - start) != size) e.norm += rv32(p);
  return e;
}

// ------------------------------------------------------------------------------------------
// Segment: .doc stream written by the PostingsWriter restatement + norms + table columns.
// ------------------------------------------------------------------------------------------
struct Column {
  int type = 0;  // 0 int64, 1 float64, 2 int32
  uint64_t rows = 0;
  std::vector<uint8_t> data;
  std::vector<uint64_t> validity;  // empty => all valid
  // integer min / max over the valid rows, kept with the column like the reference keeps ColumnBlockMeta::statistics
  // (irs/formats/column/column_reader.hpp:90-96): computed when the column is added, never at query time
  bool has_stats = false; int64_t mn = 0, mx = 0;
  bool valid(uint64_t r) const { return validity.empty() || ((validity[r >> 6] >> (r & 63)) & 1); }
  int64_t i64(uint64_t r) const {
    if (type == 2) { int32_t v; std::memcpy(&v, data.data() + 4 * r, 4); return v; }
    int64_t v; std::memcpy(&v, data.data() + 8 * r, 8); return v;
  }
  double f64(uint64_t r) const { double v; std::memcpy(&v, data.data() + 8 * r, 8); return v; }
};

}  // namespace

struct orc_segment {
  uint32_t N = 0;
  bool has_wand = false;
  float wand_b = 0.75f;
  Out doc;  // the ".doc" stream (without file header/footer)
  std::vector<orc_term_meta> terms;
  std::vector<uint32_t> norms;     // norms[d-1]
  std::vector<uint8_t> norm_bytes;  // fixed-width LE, width = norm_width (norm_column_reader.hpp:99-108)
  uint32_t norm_width = 0;
  uint64_t norm_sum = 0, norm_nonzero = 0;
  std::map<uint64_t, Column> cols;
  std::vector<bool> deleted;       // DocumentMask (index_meta.hpp:39-43) as a bitmap over doc ids; empty = none

  uint32_t norm(uint32_t doc) const { return norms.empty() ? 1u : norms[doc - 1]; }
  float avg_dl() const {  // NormReader::GetAvg, irs/formats/norm_reader_impl.hpp:83-88
    if (norm_nonzero == 0) return 0.f;
    return float(double(norm_sum) / double(norm_nonzero));
  }
};

namespace {

inline uint32_t count_max_levels(uint64_t skip_0, uint64_t skip_n, uint64_t count) {  // skip_list.cpp:38-41
  if (!(skip_0 < count)) return 0;
  uint64_t x = count / skip_0; uint32_t res = 0;
  while (x >= skip_n) { x /= skip_n; ++res; }  // basics/math_utils.hpp:111-118
  return 1 + res;
}

// PostingsWriterImpl::Write + BeginDocument/EndDocument/EndTerm/WriteSkip (writer.hpp:305-331,
// 443-488, 617-641, 699-779), SkipWriter::Skip/FlushLevels (skip_list.hpp:93-118, skip_list.cpp:77-94),
// WandWriterImpl (wand_writer.hpp:42-106).
int64_t segment_add_term(orc_segment& seg, const uint32_t* docs, const uint32_t* freqs, uint32_t n) {
  orc_term_meta meta{};
  Out& out = seg.doc;
  meta.doc_start = out.pos();
  if (n == 0) { seg.terms.push_back(meta); return int64_t(seg.terms.size()) - 1; }

  const uint32_t max_levels = std::min<uint32_t>(kMaxLevels, count_max_levels(kBlock, kSkipN, seg.N));
  Out levels[kMaxLevels];
  uint64_t skip_ptr[kMaxLevels];
  std::fill_n(skip_ptr, kMaxLevels, meta.doc_start);
  WandEntry wand[kMaxLevels + 1];
  const float avg_dl = seg.avg_dl();
  const float b = seg.wand_b;

  uint32_t buf_docs[kBlock], buf_freqs[kBlock];
  uint32_t fill = 0, last = 0, block_last = 0;
  uint32_t docs_count = 0, total_freq = 0;

  auto write_skip = [&](uint32_t level, Out& o) {
    const uint64_t doc_ptr = out.pos();
    o.v32(block_last);
    o.v64(doc_ptr - skip_ptr[level]);
    skip_ptr[level] = doc_ptr;
    if (seg.has_wand) {
      WandEntry& e = wand[level];
      o.byte(wand_size(e));                                   // writer.hpp:742-746
      wand_produce(avg_dl, b, e.freq, e.norm, wand[level + 1]);  // WandWriterImpl::Write :69-75
      wand_write(e, o);
      e = WandEntry{};
    }
  };

  for (uint32_t i = 0; i < n; ++i) {
    if (last != 0 && fill == 0) {  // writer.hpp:736-749 -> SkipWriter::Skip(docs_count, ...)
      uint32_t count = docs_count;
      if (count % kBlock == 0 && max_levels > 0) {
        write_skip(0, levels[0]);
        count /= kBlock;
        uint64_t child = levels[0].pos();
        for (uint32_t lv = 1; count % kSkipN == 0 && lv < max_levels; ++lv, count /= kSkipN) {
          write_skip(lv, levels[lv]);
          const uint64_t next_child = levels[lv].pos();
          levels[lv].v64(child);
          child = next_child;
        }
      }
    }
    assert(docs[i] > last);
    buf_docs[fill] = docs[i]; buf_freqs[fill] = freqs[i]; ++fill; last = docs[i];
    if (fill == kBlock) {  // BeginDocument :621-627
      write_doc_block(out, buf_docs, kBlock, block_last);
      write_freq_block(out, buf_freqs, kBlock);
    }
    if (seg.has_wand) wand_produce(avg_dl, b, freqs[i], seg.norm(docs[i]), wand[0]);  // Update() :63-66
    ++docs_count; total_freq += freqs[i];
    if (fill == kBlock) { block_last = last; fill = 0; }  // EndDocument :435-441
  }

  meta.docs_count = docs_count;
  meta.freq = total_freq;
  const bool has_skip_list = kBlock < docs_count;
  auto write_max_score = [&](uint32_t level) {  // EndTerm :449-455, SizeRoot :94-102
    if (!seg.has_wand) return;
    for (uint32_t l = 0; l < level; ++l) wand_produce(avg_dl, b, wand[l].freq, wand[l].norm, wand[l + 1]);
    out.byte(wand_size(wand[level]));
    wand_write(wand[level], out);
  };
  if (docs_count == 1) {
    meta.e_skip_start = docs[0] - 1;  // e_single_doc = doc - doc_limits::min()
  } else {
    if (!has_skip_list) write_max_score(0);
    if ((docs_count & (kBlock - 1)) != 0) {  // FlushTailDoc :551-559
      write_doc_block(out, buf_docs, fill, block_last);
      write_freq_block(out, buf_freqs, fill);
    }
  }
  if (has_skip_list) {
    meta.e_skip_start = out.pos() - meta.doc_start;
    uint32_t num_levels = 0;  // CountLevels, skip_list.cpp:62-75
    for (uint32_t l = max_levels; l-- > 0;) if (levels[l].pos()) { num_levels = l + 1; break; }
    write_max_score(num_levels);
    out.v32(num_levels);  // FlushLevels, skip_list.cpp:77-94
    for (uint32_t l = num_levels; l-- > 0;) { out.v64(levels[l].pos()); out.data(levels[l].b.data(), levels[l].pos()); }
  }
  seg.terms.push_back(meta);
  return int64_t(seg.terms.size()) - 1;
}

// Where a term's first block starts (after the leading wand entry of short lists; SURVEY A.1,
// iterator_score.hpp:1037-1041).
const uint8_t* term_blocks_begin(const orc_segment& seg, const orc_term_meta& m) {
  const uint8_t* p = seg.doc.b.data() + m.doc_start;
  if (seg.has_wand && m.docs_count > 1 && m.docs_count < kBlock) { const uint8_t size = *p++; p += size; }
  return p;
}

uint32_t segment_decode_term(const orc_segment& seg, uint32_t term, uint32_t* docs, uint32_t* freqs) {
  const orc_term_meta& m = seg.terms[term];
  if (m.docs_count == 0) return 0;
  if (m.docs_count == 1) { docs[0] = 1 + uint32_t(m.e_skip_start); freqs[0] = m.freq; return 1; }  // iterator_score.hpp:1015-1030
  const uint8_t* p = term_blocks_begin(seg, m);
  uint32_t left = m.docs_count, prev = 0, o = 0;
  while (left) {
    const uint32_t len = std::min(left, kBlock);
    p += read_doc_block(p, len, prev, docs + o);
    p += read_freq_block(p, len, freqs + o);
    prev = docs[o + len - 1];
    o += len; left -= len;
  }
  return m.docs_count;
}

struct SkipInfo {
  std::vector<uint32_t> last_doc;  // level-0 entries: last doc of block j
  std::vector<uint64_t> doc_ptr;   // absolute .doc offset of block j+1
  std::vector<WandEntry> wand;     // block-max pair of block j
  WandEntry root;                  // whole-list maximum
  uint32_t num_levels = 0;
};

// SkipReaderBase::Prepare (skip_list.hpp:190-236) + WandReadSkip::Read (iterator_score.hpp:240-250)
SkipInfo segment_skip(const orc_segment& seg, uint32_t term) {
  SkipInfo s;
  const orc_term_meta& m = seg.terms[term];
  const uint8_t* base = seg.doc.b.data();
  if (m.docs_count <= 1) return s;
  if (m.docs_count <= kBlock) {
    if (seg.has_wand) {
      const uint8_t* p = base + m.doc_start;
      if (m.docs_count == kBlock) { p += doc_block_size(p, kBlock); p += freq_block_size(p, kBlock); }
      const uint8_t size = *p++;
      s.root = wand_read(p, size);
    }
    return s;
  }
  const uint8_t* p = base + m.doc_start + m.e_skip_start;
  if (seg.has_wand) { const uint8_t size = *p++; s.root = wand_read(p, size); }
  s.num_levels = rv32(p);
  const uint8_t* lvl0 = nullptr; uint64_t lvl0_len = 0;
  for (uint32_t l = s.num_levels; l-- > 0;) {
    const uint64_t len = rv64(p);
    if (l == 0) { lvl0 = p; lvl0_len = len; }
    p += len;
  }
  const uint8_t* q = lvl0; const uint8_t* end = lvl0 + lvl0_len;
  uint64_t ptr = m.doc_start;
  while (q < end) {
    const uint32_t d = rv32(q);
    ptr += rv64(q);
    s.last_doc.push_back(d); s.doc_ptr.push_back(ptr);
    if (seg.has_wand) { const uint8_t size = *q++; s.wand.push_back(wand_read(q, size)); }
  }
  return s;
}

}  // namespace

// ==========================================================================================
// BM25  (irs/search/bm25.cpp)
// ==========================================================================================
namespace {

// BM25::collect :279-310. idf in double then narrowed; avg_dl divides two floats;
// norm_const = k - k*b (not k*(1-b)); norm_length = (k*b)/avg_dl.
void bm25_collect(uint64_t docs_with_field, uint64_t total_term_freq, uint64_t docs_with_term, float k,
                  float b, orc_bm25_stats* st) {
  st->idf = 0.f; st->norm_const = 0.f; st->norm_length = 0.f;
  st->idf += float(std::log1p((double(docs_with_field - docs_with_term) + 0.5) / (double(docs_with_term) + 0.5)));
  const float kb = k * b;
  st->norm_const = k - kb;
  if (total_term_freq && docs_with_field) {
    const float avg_dl = float(total_term_freq) / float(docs_with_field);
    st->norm_length = kb / avg_dl;
  } else {
    st->norm_length = kb;
  }
}
// TFIDF is selected with the sentinel k1 = -1 (b != 0: normalised by sqrt(doc length), b == 0: not): its scorer has no k / b.
inline float bm25_num(float k, float boost, float idf) { return k == -1.f ? boost * idf : boost * (k + 1) * idf; }  // bm25.cpp:224; tfidf.cpp:101 (idf{boost * idf.value})
// The scoring form BM25::PrepareScorer picks (bm25.cpp:312-365): k == 0 -> Bm1Score, b == 0 -> Bm15Score, else Bm25Score.
enum ScoreForm { kFormBm25 = 0, kFormBm15 = 1, kFormBm1 = 2, kFormTfidf = 3, kFormTfidfNorm = 4 };
inline int score_form(float k, float b) {
  if (k == -1.f) return b == 0.f ? kFormTfidf : kFormTfidfNorm;
  return k == 0.f ? kFormBm1 : b == 0.f ? kFormBm15 : kFormBm25;
}
// Bm25<MergeType,false> :90-107; Bm15<MergeType,false> :70-87 (c1 = norm_const = k, norms unused);
// Bm1Score without a filter boost zero-fills (:118-126).
// g_contract: the reference is built with clang (-ffp-contract=on is clang's default for C++) for haswell, which has FMA
// (cmake/OptimizeForArchitecture.cmake:47,71-72), so its binary most plausibly evaluates bm25.cpp:105
// `c1 = norm_const + norm_length * norm` as ONE fused multiply-add; nothing else in :105-106 has the a*b+c shape (the
// product c0*c1 is divided before it is subtracted). The oracle is compiled -ffp-contract=off and restates the
// unfused source order; orc_set_contract(1) switches c1 to the fused form so that tests can bound the difference
// between the two possible reference binaries (tests/test_oracle_goldens.py: <= 2 ulp of the score, top-k equal up
// to near-ties), both far inside north_star's 1e-5 relative bar for fp32 scores.
int g_contract = 0;
inline float bm25_one(uint32_t freq, uint32_t norm, float c0, float norm_const, float norm_length, int form = kFormBm25) {
  if (form == kFormBm1) return 0.f;
  if (form == kFormBm15) return c0 - c0 / (1.f + float(freq) / norm_const);
  // TfIdf<MergeType, HasNorm, false> (search/tfidf.cpp:59-80): sqrt(freq) * idf, divided by sqrt(norm) when normalised
  if (form == kFormTfidf) return std::sqrt(float(freq)) * c0;
  if (form == kFormTfidfNorm) return std::sqrt(float(freq)) * c0 / std::sqrt(float(norm));
  const float c1 = g_contract ? std::fmaf(norm_length, float(norm), norm_const) : norm_const + norm_length * float(norm);
  return c0 - c0 * c1 / (c1 + float(freq));
}

// ------------------------------------------------------------------------------------------
// Canonical total order on hits: score desc, then segment asc, then doc asc. The reference sorts by
// score only (doc_collector.hpp:132-134) and leaves ties unspecified; the canonical order is one
// valid outcome of it and makes parity checks deterministic.
// ------------------------------------------------------------------------------------------
inline bool hit_before(const orc_hit& l, const orc_hit& r) {
  if (l.score != r.score) return l.score > r.score;
  if (l.seg != r.seg) return l.seg < r.seg;
  return l.doc < r.doc;
}

// Exact top-k under the canonical order with the reference's buffer discipline (capacity 2k,
// select when full): NthPartitionScoreCollector restated over a total order.
struct CanonCollector {
  uint32_t k;
  float thr_in;
  std::vector<orc_hit> buf;
  orc_hit kth{};  // current k-th best (valid when have_kth)
  bool have_kth = false;
  uint64_t total = 0;
  explicit CanonCollector(uint32_t k_, float thr) : k(k_), thr_in(thr) { buf.reserve(2 * size_t(k_) + 1); }
  float threshold_score() const { return have_kth ? kth.score : thr_in; }
  inline void offer(float score, uint32_t doc, uint32_t seg) {
    ++total;
    if (!(score > thr_in)) return;
    const orc_hit h{score, doc, seg};
    if (have_kth && !hit_before(h, kth)) return;
    buf.push_back(h);
    if (buf.size() >= 2 * size_t(k)) compact();
  }
  void compact() {
    if (buf.size() <= k) return;
    std::nth_element(buf.begin(), buf.begin() + (k - 1), buf.end(), hit_before);
    buf.resize(k);
    kth = buf[k - 1];
    have_kth = true;
  }
  void finish(orc_hit* out, uint32_t* n_out) {
    compact();
    std::sort(buf.begin(), buf.end(), hit_before);
    *n_out = uint32_t(buf.size());
    std::copy(buf.begin(), buf.end(), out);
  }
};

// ------------------------------------------------------------------------------------------
// predicates (SQL three-valued logic: NULL never passes; duckdb_search_full_scan.cpp:1785-1786)
// ------------------------------------------------------------------------------------------
inline bool pred_row(const Column& c, const orc_pred& p, uint64_t r) {
  const bool valid = c.valid(r);
  if (p.op == ORC_OP_IS_NULL) return !valid;
  if (p.op == ORC_OP_IS_NOT_NULL) return valid;
  if (!valid) return false;
  if (c.type == 1) {
    const double v = c.f64(r);
    switch (p.op) {
      case ORC_OP_LT: return v < p.lo_f;
      case ORC_OP_LE: return v <= p.lo_f;
      case ORC_OP_GT: return v > p.lo_f;
      case ORC_OP_GE: return v >= p.lo_f;
      case ORC_OP_EQ: return v == p.lo_f;
      case ORC_OP_NE: return v != p.lo_f;
      case ORC_OP_BETWEEN: return v >= p.lo_f && v <= p.hi_f;
    }
  } else {
    const int64_t v = c.i64(r);
    switch (p.op) {
      case ORC_OP_LT: return v < p.lo_i;
      case ORC_OP_LE: return v <= p.lo_i;
      case ORC_OP_GT: return v > p.lo_i;
      case ORC_OP_GE: return v >= p.lo_i;
      case ORC_OP_EQ: return v == p.lo_i;
      case ORC_OP_NE: return v != p.lo_i;
      case ORC_OP_BETWEEN: return v >= p.lo_i && v <= p.hi_i;
    }
  }
  return false;
}
inline const Column* find_col(const orc_segment& s, uint64_t field) {
  auto it = s.cols.find(field);
  return it == s.cols.end() ? nullptr : &it->second;
}

// ------------------------------------------------------------------------------------------
// Posting cursor: sequential block-at-a-time decode (PostingIteratorImpl::ReadBlock/ReadTail,
// irs/formats/posting/iterator_doc.hpp:636-835) with an optional block-offset table taken from the
// level-0 skip entries so whole blocks can be skipped without decoding.
// ------------------------------------------------------------------------------------------
struct Cursor {
  const orc_segment* seg = nullptr;
  orc_term_meta m{};
  const uint8_t* p = nullptr;  // next block to decode
  uint32_t left = 0;           // postings not yet decoded
  uint32_t prev = 0;           // last doc of the previous block
  uint32_t docs[kBlock], freqs[kBlock];
  uint32_t len = 0, pos = 0;   // current decoded block
  uint32_t blk = 0;            // index of the NEXT block to decode
  SkipInfo skip;               // filled when pruning is wanted
  float c0 = 0, norm_const = 0, norm_length = 0;
  int form = kFormBm25;
  uint64_t scored = 0;

  void open(const orc_segment& s, uint32_t term, bool want_skip) {
    seg = &s; m = s.terms[term]; left = m.docs_count; prev = 0; len = pos = 0; blk = 0;
    if (m.docs_count == 1) {
      docs[0] = 1 + uint32_t(m.e_skip_start); freqs[0] = m.freq; len = 1; pos = 0; left = 0; p = nullptr;
    } else if (m.docs_count > 1) {
      p = term_blocks_begin(s, m);
    }
    if (want_skip) skip = segment_skip(s, term);
  }
  bool next_block() {
    if (!left) { len = pos = 0; return false; }
    len = std::min(left, kBlock);
    p += read_doc_block(p, len, prev, docs);
    p += read_freq_block(p, len, freqs);
    prev = docs[len - 1];
    left -= len; pos = 0; ++blk;
    return true;
  }
  // current doc or UINT32_MAX at end
  uint32_t doc() { if (pos == len && !next_block()) return 0xFFFFFFFFu; return docs[pos]; }
  float score_at(uint32_t i) {
    ++scored;
    return bm25_one(freqs[i], seg->norm(docs[i]), c0, norm_const, norm_length, form);
  }
  // Upper bound of the NEXT (undecoded) block and its last doc; blocks without a level-0 entry
  // (the final one) fall back to the list maximum (WandReadSkip::GetMaxScore, iterator_score.hpp:289-296).
  float bound_of(const WandEntry& e) const { return bm25_one(e.freq, e.norm, c0, norm_const, norm_length, form); }
  float next_block_bound() const { return blk < skip.wand.size() ? bound_of(skip.wand[blk]) : bound_of(skip.root); }
  uint32_t next_block_last() const { return blk < skip.last_doc.size() ? skip.last_doc[blk] : 0xFFFFFFFFu; }
  // Skip the next block entirely (only valid when it has a level-0 entry).
  void skip_next_block() {
    assert(blk < skip.last_doc.size());
    prev = skip.last_doc[blk];
    p = seg->doc.b.data() + skip.doc_ptr[blk];
    left -= kBlock; ++blk; len = pos = 0;
  }
};

struct QTerm { orc_bm25_term t; uint32_t docs_count; size_t order; int form = kFormBm25; };

// Terms of one segment sorted by ascending docs_count (MakeConjunction, conjunction.hpp:520-523).
std::vector<QTerm> order_terms(const orc_segment& s, const orc_bm25_term* terms, size_t n, int form) {
  std::vector<QTerm> q;
  for (size_t i = 0; i < n; ++i) q.push_back({terms[i], s.terms[terms[i].term].docs_count, i, form});
  std::stable_sort(q.begin(), q.end(), [](const QTerm& a, const QTerm& b) { return a.docs_count < b.docs_count; });
  return q;
}

// MaskDocIterator (segment_reader_impl.cpp:95-157) sits below the column filter wrap: a deleted doc is never
// seen by the collector, the filter or the match count.
inline bool filter_doc(const orc_segment& s, const orc_pred* filt, uint32_t doc) {
  if (!s.deleted.empty() && s.deleted[doc]) return false;
  if (!filt) return true;
  const Column* c = find_col(s, filt->field);
  if (!c) return false;
  return pred_row(*c, *filt, uint64_t(doc) - 1);  // row = doc - 1 (irs/index/column_extract.hpp:46-48)
}

// mode 0: dense accumulators over the whole segment.
void topk_dense(const orc_segment& s, uint32_t seg_idx, int kind, const std::vector<QTerm>& q, float k1,
                const orc_pred* filt, CanonCollector& col, uint64_t* scored) {
  std::vector<float> acc(size_t(s.N) + 1, 0.f);
  std::vector<uint8_t> cnt(size_t(s.N) + 1, 0);
  std::vector<uint32_t> docs, freqs;
  for (const QTerm& qt : q) {
    const orc_term_meta& m = s.terms[qt.t.term];
    docs.resize(m.docs_count); freqs.resize(m.docs_count);
    segment_decode_term(s, qt.t.term, docs.data(), freqs.data());
    const float c0 = bm25_num(k1, qt.t.boost, qt.t.idf);
    for (uint32_t i = 0; i < m.docs_count; ++i) {
      acc[docs[i]] = acc[docs[i]] + bm25_one(freqs[i], s.norm(docs[i]), c0, qt.t.norm_const, qt.t.norm_length, qt.form);
      ++cnt[docs[i]];
    }
    *scored += m.docs_count;
  }
  const uint8_t need = kind == ORC_QUERY_AND ? uint8_t(q.size()) : 1;
  for (uint32_t d = 1; d <= s.N; ++d) {
    if (cnt[d] < need || cnt[d] == 0) continue;
    if (!filter_doc(s, filt, d)) continue;
    col.offer(acc[d], d, seg_idx);
  }
}

// mode 1 / 2: 4096-doc windows (MaxScoreIterator::ScoreAndCollectWindow, max_score_iterator.hpp:311-356,
// window = 64x64 docs :40-42). mode 2 additionally skips (a) for a single term: any block whose
// block-max <= threshold (SingleWandIterator, iterator_score.hpp:218-233) and (b) for a disjunction:
// any window whose summed block-max bound <= threshold (UpdateWindowScores :437).
void topk_windows(const orc_segment& s, uint32_t seg_idx, int kind, const std::vector<QTerm>& q, float k1,
                  const orc_pred* filt, bool prune, CanonCollector& col, uint64_t* scored) {
  constexpr uint32_t W = 4096;
  const size_t T = q.size();
  std::vector<std::unique_ptr<Cursor>> cur;
  for (const QTerm& qt : q) {
    cur.emplace_back(new Cursor);
    Cursor& c = *cur.back();
    c.open(s, qt.t.term, prune);
    c.c0 = bm25_num(k1, qt.t.boost, qt.t.idf);
    c.form = qt.form;
    c.norm_const = qt.t.norm_const; c.norm_length = qt.t.norm_length;
  }
  float acc[W]; uint8_t cnt[W];
  std::memset(acc, 0, sizeof acc); std::memset(cnt, 0, sizeof cnt);
  const uint8_t need = kind == ORC_QUERY_AND ? uint8_t(T) : 1;
  const bool can_prune = prune && kind == ORC_QUERY_OR && !filt && s.has_wand;
  for (uint64_t lo = 1; lo <= s.N; lo += W) {
    const uint64_t hi = std::min<uint64_t>(lo + W, uint64_t(s.N) + 1);  // [lo, hi)
    if (can_prune && col.have_kth) {
      // Skip rule is strict in the canonical order: a later doc with an equal score loses the tie.
      const float thr = col.threshold_score();
      if (T == 1) {
        // handled block by block inside the consumption loop below
      } else {
        // Window bound: for each term the max bound over undecoded blocks that can reach into the
        // window; a partially consumed decoded block contributes the list maximum (safe).
        float bound = 0.f; bool all_skippable = true;
        for (auto& cp : cur) {
          Cursor& c = *cp;
          if (c.pos < c.len) { if (c.docs[c.pos] < hi) { bound += c.bound_of(c.skip.root); all_skippable = false; } continue; }
          if (!c.left) continue;
          float tb = 0.f; uint32_t b = c.blk; uint32_t first = c.prev + 1; bool reach = false;
          while (first < hi) {
            reach = true;
            const bool has = b < c.skip.wand.size();
            tb = std::max(tb, has ? c.bound_of(c.skip.wand[b]) : c.bound_of(c.skip.root));
            if (!has) break;
            first = c.skip.last_doc[b] + 1; ++b;
          }
          if (reach) bound += tb;
        }
        if (bound <= thr && all_skippable) {
          // advance every cursor past blocks that end inside the window; blocks straddling hi stay.
          for (auto& cp : cur) {
            Cursor& c = *cp;
            while (c.pos == c.len && c.left >= kBlock && c.blk < c.skip.last_doc.size() && c.next_block_last() < hi) c.skip_next_block();
          }
          // straddling blocks still have docs < hi: they must be consumed (decoded, docs < hi dropped
          // un-scored is NOT allowed for counting) -- fall through to normal processing, which only
          // touches what is left of the window.
        }
      }
    }
    bool any = false;
    for (auto& cp : cur) {
      Cursor& c = *cp;
      for (;;) {
        if (can_prune && T == 1 && col.have_kth && c.pos == c.len) {
          // SingleWandIterator: drop whole blocks whose block-max cannot beat the threshold
          // (iterator_score.hpp:218-233). Ascending doc order makes '<=' canonical-safe.
          const float thr = col.threshold_score();
          while (c.left >= kBlock && c.blk < c.skip.last_doc.size() && c.next_block_bound() <= thr) c.skip_next_block();
        }
        const uint32_t d = c.doc();
        if (d >= hi) break;
        if (d >= lo) {
          const uint32_t i = uint32_t(d - lo);
          acc[i] = acc[i] + c.score_at(c.pos);
          ++cnt[i]; any = true;
        }
        ++c.pos;
      }
    }
    if (!any) continue;
    for (uint32_t i = 0; i < uint32_t(hi - lo); ++i) {
      if (!cnt[i]) continue;
      const uint8_t cn = cnt[i]; const float sc = acc[i];
      cnt[i] = 0; acc[i] = 0.f;
      if (cn < need) continue;
      const uint32_t d = uint32_t(lo + i);
      if (!filter_doc(s, filt, d)) continue;
      col.offer(sc, d, seg_idx);
    }
  }
  for (auto& cp : cur) *scored += cp->scored;
}

}  // namespace

// ==========================================================================================
// C API
// ==========================================================================================
extern "C" {

void orc_pack128(const uint32_t* in, uint32_t* out, uint32_t bits) { pack128(in, out, bits); }
void orc_unpack128(const uint32_t* in, uint32_t* out, uint32_t bits) { unpack128_scalar(in, out, bits); }
void orc_pack128_d1(uint32_t prev, const uint32_t* in, uint32_t* out, uint32_t bits) { pack128_d1(prev, in, out, bits); }
void orc_unpack128_d1(uint32_t prev, const uint32_t* in, uint32_t* out, uint32_t bits) { unpack128_d1_scalar(prev, in, out, bits); }

int orc_use_simdcomp_ref(const char* so_path) {
  if (!so_path) { g_ref_unpack = nullptr; g_ref_unpackd1 = nullptr; return 0; }
  void* h = dlopen(so_path, RTLD_NOW | RTLD_LOCAL);
  if (!h) return -1;
  auto u = reinterpret_cast<unpack_fn>(dlsym(h, "simdunpack"));
  auto d = reinterpret_cast<unpackd1_fn>(dlsym(h, "simdunpackd1"));
  if (!u || !d) return -2;
  g_ref_unpack = u; g_ref_unpackd1 = d;
  return 0;
}

size_t orc_svb_encode(const uint32_t* in, uint32_t len, uint8_t* out) { return svb_encode(in, len, out, false, 0); }
size_t orc_svb_decode(const uint8_t* in, uint32_t* out, uint32_t len) { return svb_decode(in, out, len, false, 0); }
size_t orc_svb_delta_encode(const uint32_t* in, uint32_t len, uint8_t* out, uint32_t prev) { return svb_encode(in, len, out, true, prev); }
size_t orc_svb_delta_decode(const uint8_t* in, uint32_t* out, uint32_t len, uint32_t prev) { return svb_decode(in, out, len, true, prev); }

size_t orc_encode_doc_block(const uint32_t* docs, uint32_t len, uint32_t prev, uint8_t* out) {
  Out o; write_doc_block(o, docs, len, prev); std::memcpy(out, o.b.data(), o.pos()); return o.pos();
}
size_t orc_decode_doc_block(const uint8_t* in, uint32_t len, uint32_t prev, uint32_t* docs_out) { return read_doc_block(in, len, prev, docs_out); }
size_t orc_encode_freq_block(const uint32_t* freqs, uint32_t len, uint8_t* out) {
  Out o; write_freq_block(o, freqs, len); std::memcpy(out, o.b.data(), o.pos()); return o.pos();
}
size_t orc_decode_freq_block(const uint8_t* in, uint32_t len, uint32_t* freqs_out) { return read_freq_block(in, len, freqs_out); }

void orc_bm25_collect(uint64_t dwf, uint64_t ttf, uint64_t dwt, float k, float b, orc_bm25_stats* out) { bm25_collect(dwf, ttf, dwt, k, b, out); }
float orc_bm25_num(float k, float boost, float idf) { return bm25_num(k, boost, idf); }
void orc_bm25_score(const uint32_t* freq, const uint32_t* norm, uint32_t n, float num, float norm_const,
                    float norm_length, float* out) {
  for (uint32_t i = 0; i < n; ++i) out[i] = bm25_one(freq[i], norm ? norm[i] : 1u, num, norm_const, norm_length);
}

// NthPartitionScoreCollector, irs/index/iterators.hpp:103-250: accept score > threshold; buffer 2k;
// when full nth_element(begin, begin+k, end, score desc), threshold = hits[k].score, continue at k.
uint64_t orc_collect_nth(const float* scores, const uint32_t* docs, uint64_t n, uint32_t k, float threshold_in,
                         orc_hit* hits, uint32_t* accepted, float* threshold_out) {
  float thr = threshold_in;
  uint64_t count = 0;
  orc_hit* it = hits; orc_hit* const begin = hits; orc_hit* const pivot = hits + k; orc_hit* const end = hits + 2 * size_t(k);
  for (uint64_t i = 0; i < n; ++i) {
    ++count;
    if (!(scores[i] > thr)) continue;
    *it++ = orc_hit{scores[i], docs[i], 0};
    if (it != end) continue;
    it = pivot;
    std::nth_element(begin, pivot, end, [](const orc_hit& l, const orc_hit& r) { return l.score > r.score; });
    thr = pivot->score;
  }
  *accepted = uint32_t(it - begin);
  std::sort(begin, it, [](const orc_hit& l, const orc_hit& r) { return l.score > r.score; });  // doc_collector.hpp:132-134
  if (threshold_out) *threshold_out = thr;
  return count;
}

orc_segment* orc_segment_new(uint32_t docs_count, int has_wand, float wand_b) {
  auto* s = new orc_segment;
  s->N = docs_count; s->has_wand = has_wand != 0; s->wand_b = wand_b;
  return s;
}
void orc_segment_free(orc_segment* s) { delete s; }
void orc_segment_set_norms(orc_segment* s, const uint32_t* norms) {
  s->norms.assign(norms, norms + s->N);
  uint32_t mx = 0; s->norm_sum = 0; s->norm_nonzero = 0;
  for (uint32_t v : s->norms) { mx = std::max(mx, v); s->norm_sum += v; s->norm_nonzero += v != 0; }
  s->norm_width = mx < 256 ? 1 : mx < 65536 ? 2 : 4;
  s->norm_bytes.resize(size_t(s->N) * s->norm_width);
  for (uint32_t i = 0; i < s->N; ++i) std::memcpy(s->norm_bytes.data() + size_t(i) * s->norm_width, &s->norms[i], s->norm_width);
}
int64_t orc_segment_add_term(orc_segment* s, const uint32_t* docs, const uint32_t* freqs, uint32_t n) { return segment_add_term(*s, docs, freqs, n); }
const uint8_t* orc_segment_doc_bytes(const orc_segment* s, uint64_t* size) { *size = s->doc.pos(); return s->doc.b.data(); }
uint32_t orc_segment_num_terms(const orc_segment* s) { return uint32_t(s->terms.size()); }
void orc_segment_term_meta(const orc_segment* s, uint32_t term, orc_term_meta* out) { *out = s->terms[term]; }
uint32_t orc_segment_docs(const orc_segment* s) { return s->N; }
uint64_t orc_segment_norm_sum(const orc_segment* s) { return s->norm_sum; }
const uint8_t* orc_segment_norm_bytes(const orc_segment* s, uint32_t* w) { *w = s->norm_width; return s->norm_bytes.data(); }
uint32_t orc_segment_decode_term(const orc_segment* s, uint32_t term, uint32_t* docs, uint32_t* freqs) { return segment_decode_term(*s, term, docs, freqs); }
uint32_t orc_segment_skip_level0(const orc_segment* s, uint32_t term, uint32_t* last_doc, uint64_t* doc_ptr,
                                 uint32_t* wand_freq, uint32_t* wand_norm, uint32_t* root_freq,
                                 uint32_t* root_norm, uint32_t* num_levels) {
  const SkipInfo si = segment_skip(*s, term);
  for (size_t i = 0; i < si.last_doc.size(); ++i) {
    if (last_doc) last_doc[i] = si.last_doc[i];
    if (doc_ptr) doc_ptr[i] = si.doc_ptr[i];
    if (wand_freq && i < si.wand.size()) wand_freq[i] = si.wand[i].freq;
    if (wand_norm && i < si.wand.size()) wand_norm[i] = si.wand[i].norm;
  }
  if (root_freq) *root_freq = si.root.freq;
  if (root_norm) *root_norm = si.root.norm;
  if (num_levels) *num_levels = si.num_levels;
  return uint32_t(si.last_doc.size());
}
int orc_segment_add_column(orc_segment* s, uint64_t field, int type, const void* values, const uint64_t* validity, uint64_t rows) {
  Column c; c.type = type; c.rows = rows;
  const size_t w = type == 2 ? 4 : 8;
  c.data.assign(static_cast<const uint8_t*>(values), static_cast<const uint8_t*>(values) + rows * w);
  if (validity) c.validity.assign(validity, validity + (rows + 63) / 64);
  if (type != 1 && rows) {
    int64_t mn = std::numeric_limits<int64_t>::max(), mx = std::numeric_limits<int64_t>::min();
    bool any = false;
    for (uint64_t r = 0; r < rows; ++r) if (c.valid(r)) { const int64_t v = c.i64(r); mn = std::min(mn, v); mx = std::max(mx, v); any = true; }
    if (any) { c.has_stats = true; c.mn = mn; c.mx = mx; }
  }
  s->cols[field] = std::move(c);
  return 0;
}

int orc_segment_set_docs_mask(orc_segment* s, const uint32_t* deleted_docs, size_t n) {
  s->deleted.clear();
  if (!n) return 0;
  s->deleted.assign(size_t(s->N) + 1, false);
  for (size_t i = 0; i < n; ++i) {
    if (deleted_docs[i] == 0 || deleted_docs[i] > s->N) return -1;
    s->deleted[deleted_docs[i]] = true;
  }
  return 0;
}

int orc_bm25_topk(orc_segment* const* segs, size_t n_segs, int kind, const orc_bm25_term* terms, size_t n_terms,
                  float k1, float b, const orc_pred* filt, uint32_t k, float threshold_in, int mode, orc_hit* out,
                  uint32_t* n_out, uint64_t* total_matches, uint64_t* postings_scored) {
  if (!k || !n_terms) { *n_out = 0; if (total_matches) *total_matches = 0; return 0; }
  const int form = score_form(k1, b);
  if (form != kFormBm25 && mode == 2) mode = 1;   // the index's block-max pairs were chosen for the BM25 form (wand_type(), bm25.cpp:407-418)
  CanonCollector col(k, threshold_in);
  uint64_t scored = 0;
  for (size_t si = 0; si < n_segs; ++si) {
    const orc_segment& s = *segs[si];
    const auto q = order_terms(s, terms, n_terms, form);
    if (mode == 0) topk_dense(s, uint32_t(si), kind, q, k1, filt, col, &scored);
    else topk_windows(s, uint32_t(si), kind, q, k1, filt, mode == 2, col, &scored);
  }
  col.finish(out, n_out);
  if (total_matches) *total_matches = col.total;
  if (postings_scored) *postings_scored = scored;
  return 0;
}

}  // extern "C"

// ==========================================================================================
// Columnar: filter bitmap, COUNT/SUM, GROUP BY  (server/connector/full_scanner.cpp:81-147 scan+filter;
// the aggregate above it is DuckDB's and NOT in the tree: SUM(BIGINT) is exact 128-bit
// two's-complement, SUM/AVG(DOUBLE) is a double sum followed by one division, COUNT is u64, NULL
// inputs are skipped -- a DEFINITION (SURVEY §8c "parity unpinned" for GROUP BY), pinned only for
// COUNT/SUM by search_table_scan_10k.test.)
// ==========================================================================================
namespace {

template <class F>
void parallel_chunks(uint64_t n, int threads, uint64_t grain, F&& f) {
  if (threads <= 1 || n <= grain) { f(0, 0, n); return; }
  std::atomic<uint64_t> next{0};
  std::vector<std::thread> th;
  for (int t = 0; t < threads; ++t) {
    th.emplace_back([&, t] {
      for (;;) {  // row-group-unit claiming, like next_unit in duckdb_search_full_scan.cpp:2414-2419
        const uint64_t b = next.fetch_add(grain);
        if (b >= n) break;
        f(t, b, std::min(n, b + grain));
      }
    });
  }
  for (auto& x : th) x.join();
}

struct BoundPred { const Column* c; orc_pred p; };
bool bind_preds(const orc_segment& s, const orc_pred* preds, size_t n, std::vector<BoundPred>& out) {
  out.clear();
  for (size_t i = 0; i < n; ++i) {
    const Column* c = find_col(s, preds[i].field);
    if (!c) return false;
    out.push_back({c, preds[i]});
  }
  return true;
}
inline bool row_passes(const std::vector<BoundPred>& bp, uint64_t r) {
  for (const auto& b : bp) if (!pred_row(*b.c, b.p, r)) return false;
  return true;
}
inline uint64_t seg_rows(const orc_segment& s, const std::vector<BoundPred>& bp, const Column* extra) {
  if (!bp.empty()) return bp[0].c->rows;
  if (extra) return extra->rows;
  return s.N;
}

struct Agg { uint64_t count = 0; __int128 sum_i = 0; double sum_f = 0; uint64_t cnt_f = 0; };

}  // namespace

extern "C" {

int orc_filter_bitmap(const orc_segment* s, const orc_pred* preds, size_t n_preds, uint64_t* mask_out) {
  std::vector<BoundPred> bp;
  if (!bind_preds(*s, preds, n_preds, bp)) return -1;
  const uint64_t rows = seg_rows(*s, bp, nullptr);
  std::memset(mask_out, 0, ((rows + 63) / 64) * 8);
  for (uint64_t r = 0; r < rows; ++r) if (row_passes(bp, r)) mask_out[r >> 6] |= uint64_t(1) << (r & 63);
  return 0;
}

int orc_filter_count_sum(orc_segment* const* segs, size_t n_segs, const orc_pred* preds, size_t n_preds,
                         uint64_t sum_field, int threads, uint64_t* count, int64_t sum_i128[2], double* sum_f64) {
  threads = std::max(threads, 1);
  std::vector<Agg> part(static_cast<size_t>(threads), Agg{});
  for (size_t si = 0; si < n_segs; ++si) {
    const orc_segment& s = *segs[si];
    std::vector<BoundPred> bp;
    if (!bind_preds(s, preds, n_preds, bp)) return -1;
    const Column* sc = find_col(s, sum_field);
    const uint64_t rows = seg_rows(s, bp, sc);
    parallel_chunks(rows, threads, uint64_t(1) << 20, [&](int t, uint64_t b, uint64_t e) {
      Agg a;
      for (uint64_t r = b; r < e; ++r) {
        if (!row_passes(bp, r)) continue;
        ++a.count;
        if (sc && sc->valid(r)) {
          if (sc->type == 1) a.sum_f += sc->f64(r); else a.sum_i += sc->i64(r);
        }
      }
      Agg& d = part[size_t(t)];
      d.count += a.count; d.sum_i += a.sum_i; d.sum_f += a.sum_f;
    });
  }
  Agg tot;
  for (const Agg& a : part) { tot.count += a.count; tot.sum_i += a.sum_i; tot.sum_f += a.sum_f; }
  *count = tot.count;
  if (sum_i128) { sum_i128[0] = int64_t(uint64_t(tot.sum_i)); sum_i128[1] = int64_t(tot.sum_i >> 64); }
  if (sum_f64) *sum_f64 = tot.sum_f;
  return 0;
}

}  // extern "C"

// Vectorised in the way the reference's engine executes this plan (DuckDB: 2048-row vectors, one selection vector
// per chunk, predicates evaluated column-at-a-time with the type switch outside the row loop, thread-local aggregate
// states merged at the end): restated, since the DuckDB fork is not vendored.
namespace {
constexpr uint64_t kVec = 2048;            // STANDARD_VECTOR_SIZE
constexpr uint64_t kRowGroup = 60 * kVec;  // DuckDB row group = 122880 rows: the unit a worker claims

template <class V, class P>
inline uint32_t select_typed(const V* v, uint64_t base, const uint16_t* in, uint32_t n_in, bool dense_in, uint16_t* out, P pass) {
  uint32_t n = 0;
  if (dense_in) { for (uint32_t i = 0; i < n_in; ++i) { out[n] = uint16_t(i); n += pass(v[base + i]) ? 1u : 0u; } }
  else { for (uint32_t i = 0; i < n_in; ++i) { const uint16_t r = in[i]; out[n] = r; n += pass(v[base + r]) ? 1u : 0u; } }
  return n;
}
template <class V>
inline uint32_t select_op(const V* v, uint64_t base, const uint16_t* in, uint32_t n_in, bool dense_in, uint16_t* out, int op, V lo, V hi) {
  switch (op) {
    case ORC_OP_LT: return select_typed(v, base, in, n_in, dense_in, out, [=](V x) { return x < lo; });
    case ORC_OP_LE: return select_typed(v, base, in, n_in, dense_in, out, [=](V x) { return x <= lo; });
    case ORC_OP_GT: return select_typed(v, base, in, n_in, dense_in, out, [=](V x) { return x > lo; });
    case ORC_OP_GE: return select_typed(v, base, in, n_in, dense_in, out, [=](V x) { return x >= lo; });
    case ORC_OP_EQ: return select_typed(v, base, in, n_in, dense_in, out, [=](V x) { return x == lo; });
    case ORC_OP_NE: return select_typed(v, base, in, n_in, dense_in, out, [=](V x) { return x != lo; });
    default: return select_typed(v, base, in, n_in, dense_in, out, [=](V x) { return x >= lo && x <= hi; });
  }
}
// One predicate over one vector: narrows the selection. Nullable columns and the NULL tests take the row-wise path.
inline uint32_t select_pred(const BoundPred& b, uint64_t base, const uint16_t* in, uint32_t n_in, bool dense_in, uint16_t* out) {
  const Column& c = *b.c;
  if (!c.validity.empty() || b.p.op == ORC_OP_IS_NULL || b.p.op == ORC_OP_IS_NOT_NULL) {
    uint32_t n = 0;
    for (uint32_t i = 0; i < n_in; ++i) { const uint16_t r = dense_in ? uint16_t(i) : in[i]; if (pred_row(c, b.p, base + r)) out[n++] = r; }
    return n;
  }
  if (c.type == 1) return select_op<double>(reinterpret_cast<const double*>(c.data.data()), base, in, n_in, dense_in, out, b.p.op, b.p.lo_f, b.p.hi_f);
  if (c.type == 2) {
    const int64_t lo = std::max<int64_t>(std::min<int64_t>(b.p.lo_i, INT32_MAX), INT32_MIN), hi = std::max<int64_t>(std::min<int64_t>(b.p.hi_i, INT32_MAX), INT32_MIN);
    if (lo != b.p.lo_i || hi != b.p.hi_i) {   // bound outside int32: keep the exact 64-bit comparison
      uint32_t n = 0;
      for (uint32_t i = 0; i < n_in; ++i) { const uint16_t r = dense_in ? uint16_t(i) : in[i]; if (pred_row(c, b.p, base + r)) out[n++] = r; }
      return n;
    }
    return select_op<int32_t>(reinterpret_cast<const int32_t*>(c.data.data()), base, in, n_in, dense_in, out, b.p.op, int32_t(lo), int32_t(hi));
  }
  return select_op<int64_t>(reinterpret_cast<const int64_t*>(c.data.data()), base, in, n_in, dense_in, out, b.p.op, b.p.lo_i, b.p.hi_i);
}
struct DenseAgg { __int128 sum_i = 0; double sum_f = 0; uint32_t count = 0; uint32_t cnt_f = 0; };   // 32 B: one group = half a cache line
}  // namespace

extern "C" {

int orc_filter_groupby(orc_segment* const* segs, size_t n_segs, const orc_pred* preds, size_t n_preds,
                       uint64_t key_field, uint64_t sum_int_field, uint64_t avg_f64_field, int threads,
                       orc_group_row* out, uint64_t cap, uint64_t* n_out) {
  threads = std::max(threads, 1);
  // Key range from the column statistics (ColumnBlockMeta, irs/formats/column/column_reader.hpp:90-96); a small
  // range selects the dense ("perfect hash") aggregate.
  int64_t kmin = std::numeric_limits<int64_t>::max(), kmax = std::numeric_limits<int64_t>::min();
  for (size_t si = 0; si < n_segs; ++si) {
    const Column* kc = find_col(*segs[si], key_field);
    if (!kc || !kc->validity.empty() || kc->type == 1) return -1;  // GROUP BY key must be a NOT NULL integer here
    if (kc->has_stats) { kmin = std::min(kmin, kc->mn); kmax = std::max(kmax, kc->mx); }
  }
  if (kmin > kmax) { *n_out = 0; return 0; }
  { uint64_t all_rows = 0; for (size_t si = 0; si < n_segs; ++si) all_rows += find_col(*segs[si], key_field)->rows; if (all_rows >> 32) return -3; }   // 32-bit group counts
  const bool dense = (unsigned __int128)((__int128)kmax - kmin) < ((unsigned __int128)1 << 24);
  const uint64_t span = dense ? uint64_t(kmax - kmin) + 1 : 0;
  // Thread-local aggregate states come from a pool that outlives the call (an engine keeps its buffers; first-touch
  // page faults of ~3 MB x threads would otherwise dominate a 100 M-row scan). A table is zeroed by the worker that
  // uses it, the first time it claims a row group in this call.
  static std::vector<std::vector<DenseAgg>> pool;
  static std::mutex pool_mu;
  std::lock_guard<std::mutex> pool_lock(pool_mu);
  if (dense && pool.size() < size_t(threads)) pool.resize(size_t(threads));
  std::vector<std::vector<DenseAgg>>& dpart = pool;
  std::vector<uint8_t> used(size_t(threads), 0);
  std::vector<std::unordered_map<int64_t, DenseAgg>> hpart(dense ? 0 : size_t(threads));
  for (size_t si = 0; si < n_segs; ++si) {
    const orc_segment& s = *segs[si];
    std::vector<BoundPred> bp;
    if (!bind_preds(s, preds, n_preds, bp)) return -1;
    const Column* kc = find_col(s, key_field);
    const Column* ic = find_col(s, sum_int_field);
    const Column* fc = find_col(s, avg_f64_field);
    parallel_chunks(kc->rows, threads, kRowGroup, [&](int t, uint64_t b, uint64_t e) {
      if (dense && !used[size_t(t)]) {
        auto& tab = dpart[size_t(t)];
        if (tab.size() < span) tab.resize(span);
        std::fill(tab.begin(), tab.begin() + span, DenseAgg{});
        used[size_t(t)] = 1;
      }
      uint16_t sel_a[kVec], sel_b[kVec];
      for (uint64_t base = b; base < e; base += kVec) {
        uint32_t n = uint32_t(std::min<uint64_t>(kVec, e - base));
        const uint16_t* sel = nullptr;   // nullptr = every row of the vector
        uint16_t* bufs[2] = {sel_a, sel_b};
        int which = 0;
        for (const BoundPred& p : bp) {
          n = select_pred(p, base, sel, n, sel == nullptr, bufs[which]);
          sel = bufs[which]; which ^= 1;
          if (!n) break;
        }
        for (uint32_t i = 0; i < n; ++i) {
          const uint64_t r = base + (sel ? sel[i] : i);
          const int64_t key = kc->i64(r);
          DenseAgg& a = dense ? dpart[size_t(t)][uint64_t(key - kmin)] : hpart[size_t(t)][key];
          ++a.count;
          if (ic && ic->valid(r)) a.sum_i += ic->i64(r);
          if (fc && fc->valid(r)) { a.sum_f += fc->f64(r); ++a.cnt_f; }
        }
      }
    });
  }
  auto emit_row = [](orc_group_row& g, int64_t key, const DenseAgg& a) {
    g.key = key; g.count = a.count;
    g.sum_i128[0] = int64_t(uint64_t(a.sum_i)); g.sum_i128[1] = int64_t(a.sum_i >> 64);
    g.sum_f64 = a.sum_f; g.cnt_f64 = a.cnt_f;
  };
  if (dense) {
    // merge by key range, one range per thread (thread-local states are combined in thread order, so the floating-
    // point sum of a group does not depend on the merge's own parallelism); output in key order
    std::vector<DenseAgg> merged(span);
    const uint64_t parts = uint64_t(std::min<uint64_t>(uint64_t(threads), std::max<uint64_t>(1, span / 1024)));
    std::vector<uint64_t> part_groups(parts + 1, 0);
    parallel_chunks(parts, threads, 1, [&](int, uint64_t pb, uint64_t pe) {
      for (uint64_t pi = pb; pi < pe; ++pi) {
        const uint64_t lo = span * pi / parts, hi = span * (pi + 1) / parts;
        uint64_t groups = 0;
        for (uint64_t i = lo; i < hi; ++i) {
          DenseAgg tot;
          for (size_t ti = 0; ti < size_t(threads); ++ti) if (used[ti]) { const DenseAgg& a = dpart[ti][i]; tot.count += a.count; tot.sum_i += a.sum_i; tot.sum_f += a.sum_f; tot.cnt_f += a.cnt_f; }
          merged[i] = tot;
          groups += tot.count ? 1 : 0;
        }
        part_groups[pi + 1] = groups;
      }
    });
    for (uint64_t pi = 0; pi < parts; ++pi) part_groups[pi + 1] += part_groups[pi];
    *n_out = part_groups[parts];
    if (*n_out > cap) return -2;
    parallel_chunks(parts, threads, 1, [&](int, uint64_t pb, uint64_t pe) {
      for (uint64_t pi = pb; pi < pe; ++pi) {
        uint64_t o = part_groups[pi];
        for (uint64_t i = span * pi / parts; i < span * (pi + 1) / parts; ++i)
          if (merged[i].count) emit_row(out[o++], kmin + int64_t(i), merged[i]);
      }
    });
    return 0;
  }
  std::unordered_map<int64_t, DenseAgg> all;
  for (auto& p : hpart) for (auto& kv : p) { DenseAgg& d = all[kv.first]; d.count += kv.second.count; d.sum_i += kv.second.sum_i; d.sum_f += kv.second.sum_f; d.cnt_f += kv.second.cnt_f; }
  std::vector<int64_t> keys; keys.reserve(all.size());
  for (auto& kv : all) keys.push_back(kv.first);
  std::sort(keys.begin(), keys.end());
  *n_out = keys.size();
  if (keys.size() > cap) return -2;
  for (size_t i = 0; i < keys.size(); ++i) emit_row(out[i], keys[i], all[keys[i]]);
  return 0;
}

// ==========================================================================================
// Deterministic synthetic inputs. splitmix64 finaliser over seed ^ (stream<<48) ^ index.
// ==========================================================================================
void orc_set_contract(int on) { g_contract = on ? 1 : 0; }

// TFIDF::collect (search/tfidf.cpp:149-150): idf = (float) log1p((docs_with_field + 1.0) / (docs_with_term + 1.0))
float orc_tfidf_idf(uint64_t docs_with_field, uint64_t docs_with_term) {
  return static_cast<float>(std::log1p((double(docs_with_field) + 1.0) / (double(docs_with_term) + 1.0)));
}

// test hook: SkipWriter::Prepare's level count (skip_list.cpp:38-41,50-51) as the writer above uses it
uint32_t orc_count_max_levels(uint64_t skip_0, uint64_t skip_n, uint64_t count) { return count_max_levels(skip_0, skip_n, count); }

uint64_t orc_synth_hash(uint64_t stream, uint64_t index) {
  uint64_t z = (UINT64_C(0x5EDB2026) ^ (stream << 48) ^ index) + UINT64_C(0x9E3779B97F4A7C15);
  z = (z ^ (z >> 30)) * UINT64_C(0xBF58476D1CE4E5B9);
  z = (z ^ (z >> 27)) * UINT64_C(0x94D049BB133111EB);
  return z ^ (z >> 31);
}

void orc_synth_column(uint64_t stream, int kind, uint64_t row0, uint64_t rows, void* out) {
  auto* oi = static_cast<int64_t*>(out);
  auto* of = static_cast<double*>(out);
  for (uint64_t i = 0; i < rows; ++i) {
    const uint64_t h = orc_synth_hash(stream, row0 + i);
    switch (kind) {
      case 0: oi[i] = int64_t(h % 100000); break;
      case 1: oi[i] = int64_t(h % 1000000); break;
      case 2: of[i] = double(h >> 11) * 0x1.0p-53; break;
      case 3: oi[i] = int64_t(h % 2001) - 1000; break;
      case 4: of[i] = double(h >> 11) * 0x1.0p-53 * 1000.0; break;
      case 7: oi[i] = int64_t((row0 + i) / 100); break;   // clustered (an insertion timestamp)
      default: oi[i] = int64_t(h); break;
    }
  }
}

void orc_synth_doc_lengths(uint64_t doc0, uint32_t n, uint32_t* out) {
  for (uint32_t i = 0; i < n; ++i) out[i] = 16 + uint32_t(orc_synth_hash(1, doc0 + 1 + i) % 240);
}

uint32_t orc_synth_term(uint32_t t, uint64_t doc0, uint32_t n, const uint32_t* dl, uint32_t* docs, uint32_t* freqs) {
  // terms 1000000 .. 1000004: BASELINE configs[3]'s conjunction terms, p = 0.50, 0.40, 0.30, 0.25, 0.20
  static const double kCfg4P[5] = {0.50, 0.40, 0.30, 0.25, 0.20};
  const double p = (t >= 1000000u && t < 1000005u) ? kCfg4P[t - 1000000u] : std::min(0.5, 0.6 / double(t + 1));
  const uint64_t thr = uint64_t(std::ldexp(p, 64));
  uint32_t c = 0;
  for (uint32_t i = 0; i < n; ++i) {
    const uint64_t g = doc0 + 1 + i;
    if (orc_synth_hash(100 + t, g) >= thr) continue;
    const uint64_t h2 = orc_synth_hash(1000 + t, g);
    uint32_t f = 1 + (h2 ? uint32_t(__builtin_ctzll(h2)) : 64);
    if (f > dl[i]) f = dl[i];
    docs[c] = i + 1; freqs[c] = f; ++c;
  }
  return c;
}

}  // extern "C"

// ==========================================================================================
// CPU-baseline helpers: many queries on all host cores (one query per thread at a time, the way
// DuckDB workers each drive their own iterator: duckdb_search_full_scan.cpp:1925-1944), and a
// multi-threaded builder for the synthetic shard.
// ==========================================================================================
extern "C" {

int orc_bm25_topk_batch(orc_segment* const* segs, size_t n_segs, int kind, const orc_bm25_term* terms,
                        const uint32_t* term_off, size_t n_queries, float k1, float b, const orc_pred* filt, uint32_t k,
                        float threshold_in, int mode, int threads, orc_hit* out, uint32_t* n_out,
                        uint64_t* total_matches, uint64_t* postings_scored) {
  threads = std::max(threads, 1);
  std::atomic<size_t> next{0};
  std::atomic<uint64_t> scored_all{0};
  auto work = [&]() {
    for (;;) {
      const size_t q = next.fetch_add(1);
      if (q >= n_queries) break;
      uint64_t scored = 0, tot = 0;
      orc_bm25_topk(segs, n_segs, kind, terms + term_off[q], term_off[q + 1] - term_off[q], k1, b, filt, k, threshold_in,
                    mode, out + q * size_t(k), n_out + q, &tot, &scored);
      if (total_matches) total_matches[q] = tot;
      scored_all += scored;
    }
  };
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t) pool.emplace_back(work);
  for (auto& t : pool) t.join();
  if (postings_scored) *postings_scored = scored_all.load();
  return 0;
}

// Synthetic shard built with `threads` workers: docs (doc0, doc0+n], terms [t0, t0+nt).
// Terms are generated in parallel and appended in term order, so the stream equals the sequential one.
orc_segment* orc_synth_segment(uint64_t doc0, uint32_t n_docs, uint32_t t0, uint32_t nt, int threads,
                               uint32_t* docs_count_out, uint64_t* sum_dl_out) {
  threads = std::max(threads, 1);
  auto* seg = orc_segment_new(n_docs, 1, 0.75f);
  std::vector<uint32_t> dl(n_docs);
  orc_synth_doc_lengths(doc0, n_docs, dl.data());
  orc_segment_set_norms(seg, dl.data());
  std::vector<std::vector<uint32_t>> docs(nt), freqs(nt);
  std::atomic<uint32_t> next{0};
  auto work = [&]() {
    for (;;) {
      const uint32_t i = next.fetch_add(1);
      if (i >= nt) break;
      docs[i].resize(n_docs); freqs[i].resize(n_docs);
      const uint32_t c = orc_synth_term(t0 + i, doc0, n_docs, dl.data(), docs[i].data(), freqs[i].data());
      docs[i].resize(c); freqs[i].resize(c);
      docs[i].shrink_to_fit(); freqs[i].shrink_to_fit();
    }
  };
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t) pool.emplace_back(work);
  for (auto& t : pool) t.join();
  for (uint32_t i = 0; i < nt; ++i) {
    segment_add_term(*seg, docs[i].data(), freqs[i].data(), uint32_t(docs[i].size()));
    if (docs_count_out) docs_count_out[i] = uint32_t(docs[i].size());
    std::vector<uint32_t>().swap(docs[i]); std::vector<uint32_t>().swap(freqs[i]);
  }
  if (sum_dl_out) *sum_dl_out = seg->norm_sum;
  return seg;
}

}  // extern "C"
