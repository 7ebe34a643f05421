// posting_format.hpp -- host-side knowledge of the "1_5simd" posting format (product code).
//
// Mirrors irs/formats/posting/format_block_128.hpp (block codec), writer.hpp / skip_list.hpp
// (stream layout) and wand_writer.hpp (block-max entries) of the reference; paths relative to
// /root/reference/libs/iresearch/include/iresearch. Used by the index-build side (PostingWriter,
// synthetic corpus) and by staging, which turns a ".doc" stream into the block table + aligned
// payload arena the kernels read. Nothing here runs at query time.
#pragma once

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace sdbg {

constexpr uint32_t kBlockSize = 128;  // doc_limits::kBlockSize, utils/type_limits.hpp:49
constexpr uint32_t kSkipN = 32;       // doc_limits::kSkipSize
constexpr uint32_t kMaxSkipLevels = 5;

// DeltaEncoding / Encoding header bytes, format_block_128.hpp:652-713, :722-770.
enum : uint8_t {
  kDeValues = 0, kDeSame08 = 1, kDeSame16 = 2, kDeSame32 = 3, kDeBitset = 4, kDeSvb = 5,
  kDeForSvb = 6 /* reserved, never written */, kDeDeltaSvb = 7, kDeBitpack02 = 8 /* +b-2, b<=31 */
};
enum : uint8_t { kEValues = 0, kESame08 = 1, kESame16 = 2, kESame32 = 3, kESvb = 4, kEBitpack01 = 5 /* +b-1 */ };

// ---- little-endian / vint helpers (utils/bytes_utils.hpp:93-175) ----
struct ByteWriter {
  std::vector<uint8_t> buf;
  size_t size() const { return buf.size(); }
  void put(uint8_t b) { buf.push_back(b); }
  void put16(uint32_t v) { put(uint8_t(v)); put(uint8_t(v >> 8)); }
  void put32(uint32_t v) { put16(v); put16(v >> 16); }
  void put_bytes(const void* p, size_t n) { auto c = static_cast<const uint8_t*>(p); buf.insert(buf.end(), c, c + n); }
  template <class U> void put_vint(U v) { while (v > 0x7F) { put(uint8_t(v & 0x7F) | 0x80); v >>= 7; } put(uint8_t(v)); }
};
template <class U> inline U get_vint(const uint8_t*& p, const uint8_t* end) {
  U v = 0; unsigned shift = 0;
  while (p < end) { const uint8_t c = *p++; v |= U(c & 0x7F) << shift; if (!(c & 0x80)) break; shift += 7; }
  return v;
}
inline uint32_t vint_len(uint32_t v) { uint32_t n = 1; while (v > 0x7F) { v >>= 7; ++n; } return n; }
inline uint32_t load16(const uint8_t* p) { return uint32_t(p[0]) | (uint32_t(p[1]) << 8); }
inline uint32_t load32(const uint8_t* p) { uint32_t v; std::memcpy(&v, p, 4); return v; }

// ---- payload geometry: how many payload bytes follow a header byte ----
// (SizeDelta / Size, format_block_128.hpp:851-951). Returns SIZE_MAX on a corrupt header.
inline size_t doc_payload_bytes(const uint8_t* hdr, const uint8_t* end, uint32_t len) {
  if (hdr >= end) return SIZE_MAX;
  const uint8_t e = hdr[0];
  switch (e) {
    case kDeValues: return size_t(len) * 4;
    case kDeSame08: return 1;
    case kDeSame16: return 2;
    case kDeSame32: return 4;
    case kDeBitset: return hdr + 1 < end ? 1 + size_t(hdr[1]) * 8 : SIZE_MAX;
    case kDeSvb: case kDeDeltaSvb: return hdr + 2 < end ? 2 + size_t(load16(hdr + 1)) : SIZE_MAX;
    case kDeForSvb: return SIZE_MAX;
    default: {
      const uint32_t bits = uint32_t(e - kDeBitpack02) + 2;
      return (bits <= 31 && len == kBlockSize) ? size_t(bits) * 16 : SIZE_MAX;
    }
  }
}
inline size_t freq_payload_bytes(const uint8_t* hdr, const uint8_t* end, uint32_t len) {
  if (hdr >= end) return SIZE_MAX;
  const uint8_t e = hdr[0];
  switch (e) {
    case kEValues: return size_t(len) * 4;
    case kESame08: return 1;
    case kESame16: return 2;
    case kESame32: return 4;
    case kESvb: return hdr + 2 < end ? 2 + size_t(load16(hdr + 1)) : SIZE_MAX;
    default: {
      const uint32_t bits = uint32_t(e - kEBitpack01) + 1;
      return (bits <= 31 && len == kBlockSize) ? size_t(bits) * 16 : SIZE_MAX;
    }
  }
}

// ---- scalar block codec (host) ----
void encode_doc_block(ByteWriter& out, const uint32_t* docs, uint32_t len, uint32_t prev);
void encode_freq_block(ByteWriter& out, const uint32_t* freqs, uint32_t len);
// Decodes into out[0..len); returns bytes consumed (header included) or 0 on error.
size_t decode_doc_block(const uint8_t* p, const uint8_t* end, uint32_t len, uint32_t prev, uint32_t* out);
size_t decode_freq_block(const uint8_t* p, const uint8_t* end, uint32_t len, uint32_t* out);

// ---- block-max pair (wand_writer.hpp:178-218, 366-381) ----
struct MaxPair { uint32_t freq = 1; uint32_t norm = 0xFFFFFFFFu; };

struct TermMeta {  // == sdbg_term_meta
  uint32_t docs_count = 0;
  uint32_t freq = 0;
  uint64_t doc_start = 0;
  uint64_t e_skip_start = 0;
};

// PostingsWriterImpl mirror: one ".doc" stream per segment (writer.hpp:443-488, 617-641, 699-779).
class PostingWriter {
 public:
  // Copies `norms` (per-doc field lengths, may be null) and derives the segment's average length.
  PostingWriter(uint32_t segment_docs, bool has_wand, float wand_b, const uint32_t* norms);
  // Borrows `norms` (must outlive the writer) with a precomputed average length.
  PostingWriter(uint32_t segment_docs, bool has_wand, float wand_b, const uint32_t* borrowed_norms, float avg_dl);
  void add_term(const uint32_t* docs, const uint32_t* freqs, uint32_t n);
  const std::vector<uint8_t>& bytes() const { return out_.buf; }
  const std::vector<TermMeta>& terms() const { return terms_; }
  // Appends another writer's stream (used by the multi-threaded corpus builder).
  void append(const PostingWriter& other);

 private:
  uint32_t norm_of(uint32_t doc) const { return norms_ptr_ ? norms_ptr_[doc - 1] : 1u; }
  void fold(const MaxPair& from, MaxPair& to) const;
  void feed(uint32_t freq, uint32_t norm, MaxPair& to) const;
  uint32_t segment_docs_;
  bool has_wand_;
  float b_;
  float avg_dl_ = 0.f;
  std::vector<uint32_t> norms_;
  const uint32_t* norms_ptr_ = nullptr;
  ByteWriter out_;
  std::vector<TermMeta> terms_;
};

// ---- staged form: what the kernels read ----
// One 16-byte descriptor per posting block (see DESIGN.md "HBM layout").
struct BlockDesc {
  uint32_t off16;      // doc payload offset in the arena, in 16-byte units
  uint32_t last_doc;   // last doc id of the block
  uint32_t prev_last;  // last doc id of the previous block of the same term (0 for the first)
  uint32_t packed;     // doc_enc | freq_enc<<6 | (len-1)<<12 | freq_off16_delta<<19 | bitset_words<<25
};
static_assert(sizeof(BlockDesc) == 16, "descriptor must be one 16-byte vector load");
inline uint32_t pack_desc(uint32_t doc_enc, uint32_t freq_enc, uint32_t len, uint32_t fdelta, uint32_t words) {
  return doc_enc | (freq_enc << 6) | ((len - 1) << 12) | (fdelta << 19) | (words << 25);
}

struct StagedPostings {
  std::vector<uint8_t> arena;          // 16-byte aligned payloads: [doc payload][pad][freq payload][pad] ...
  std::vector<BlockDesc> blocks;       // all terms back to back
  std::vector<uint32_t> term_blk_begin;  // n_terms + 1
  std::vector<uint32_t> term_docs;     // docs_count per term
  std::vector<MaxPair> blk_max;        // per block; blocks without a skip entry carry the term's root pair
  std::vector<uint32_t> blk_anchor;    // per block 4 x u32: doc ids of postings 31, 63, 95 (0xFFFFFFFF past the block's
                                       // length) and 0 -- lets one GPU lane find a doc inside a bit-packed block by
                                       // summing at most 32 gaps (probes of non-essential / conjunction lists)
  std::vector<MaxPair> term_max;       // per term root pair ({0,0} when !has_wand)
  std::vector<uint64_t> term_bytes;    // per term: encoded block bytes in the .doc stream (headers + payloads)
  std::vector<uint8_t> term_probe;     // per term: 1 when (nearly) every full block is a bitset with random-access freqs
                                       // (a doc can be looked up without decoding the block: driver-mode probes)
  uint64_t n_postings = 0;
  bool has_wand = false;
};

// Parses a ".doc" stream into the staged form. Returns "" or an error text.
std::string stage_postings(const uint8_t* doc, size_t n, const TermMeta* terms, size_t n_terms, bool has_wand,
                           StagedPostings* out);

}  // namespace sdbg
