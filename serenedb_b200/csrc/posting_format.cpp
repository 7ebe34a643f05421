// posting_format.cpp -- host-side "1_5simd" codec, PostingWriter mirror and staging parser.
// See posting_format.hpp for the reference files each part follows.
#include "posting_format.hpp"

#include <algorithm>
#include <cmath>

namespace sdbg {
namespace {

inline uint32_t nbytes_1234(uint32_t v) { return 1 + (v > 0xFF) + (v > 0xFFFF) + (v > 0xFFFFFF); }
inline uint32_t width_of(uint32_t v) { return v ? 32u - uint32_t(__builtin_clz(v)) : 0u; }

// simdcomp "vertical" layout: value i lives in 32-bit lane (i & 3), at bit (i >> 2) * b of that
// lane's stream; stream word w of lane l is stored at out[4 * w + l].
void vertical_pack(const uint32_t* v, uint32_t b, uint32_t* out) {
  std::fill(out, out + 4 * b, 0u);
  for (uint32_t lane = 0; lane < 4; ++lane) {
    uint64_t acc = 0; uint32_t have = 0, w = 0;
    for (uint32_t row = 0; row < 32; ++row) {
      acc |= uint64_t(v[4 * row + lane]) << have;
      have += b;
      if (have >= 32) { out[4 * w++ + lane] = uint32_t(acc); acc >>= 32; have -= 32; }
    }
  }
}
void vertical_unpack(const uint8_t* in, uint32_t b, uint32_t* v) {
  const uint64_t mask = (uint64_t(1) << b) - 1;
  for (uint32_t lane = 0; lane < 4; ++lane) {
    uint64_t acc = 0; uint32_t have = 0, w = 0;
    for (uint32_t row = 0; row < 32; ++row) {
      if (have < b) { acc |= uint64_t(load32(in + 4 * (4 * w++ + lane))) << have; have += 32; }
      v[4 * row + lane] = uint32_t(acc & mask);
      acc >>= b; have -= b;
    }
  }
}

// StreamVByte 1234: control bytes first, 2 bits per value, then the value bytes.
size_t svb_put(ByteWriter& out, const uint32_t* v, uint32_t n) {
  const size_t start = out.size();
  const uint32_t nctl = (n + 3) / 4;
  out.buf.resize(start + nctl, 0);
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t nb = nbytes_1234(v[i]);
    out.buf[start + i / 4] |= uint8_t((nb - 1) << (2 * (i % 4)));
    for (uint32_t k = 0; k < nb; ++k) out.put(uint8_t(v[i] >> (8 * k)));
  }
  return out.size() - start;
}
bool svb_get(const uint8_t* p, const uint8_t* end, uint32_t n, uint32_t* v) {
  const uint32_t nctl = (n + 3) / 4;
  if (p + nctl > end) return false;
  const uint8_t* d = p + nctl;
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t nb = ((p[i / 4] >> (2 * (i % 4))) & 3) + 1;
    if (d + nb > end) return false;
    uint32_t x = 0;
    for (uint32_t k = 0; k < nb; ++k) x |= uint32_t(d[k]) << (8 * k);
    d += nb; v[i] = x;
  }
  return true;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// Encoders: candidates are evaluated in the reference's order with its strict '<' tests
// (WriteTailDelta format_block_128.hpp:57-154, WriteTail :249-308).
// ------------------------------------------------------------------------------------------
void encode_doc_block(ByteWriter& out, const uint32_t* docs, uint32_t len, uint32_t prev) {
  uint32_t gaps[kBlockSize];
  uint32_t max_gap = 0, svb_raw = 2 + (len + 3) / 4, svb_gap = svb_raw;
  bool uniform = true;
  for (uint32_t i = 0, p = prev; i < len; ++i) {
    gaps[i] = docs[i] - p; p = docs[i];
    uniform &= gaps[i] == gaps[0];
    max_gap = std::max(max_gap, gaps[i]);
    svb_raw += nbytes_1234(docs[i]);
    svb_gap += nbytes_1234(gaps[i]);
  }
  if (uniform) {  // :103-120 -- wins outright
    if (max_gap <= 0xFF) { out.put(kDeSame08); out.put(uint8_t(max_gap)); }
    else if (max_gap <= 0xFFFF) { out.put(kDeSame16); out.put16(max_gap); }
    else { out.put(kDeSame32); out.put32(max_gap); }
    return;
  }
  uint8_t choice = kDeValues;
  uint32_t cost = len * 4;
  if (len == kBlockSize) {
    const uint32_t b = width_of(max_gap);
    if (16 * b < cost) { choice = uint8_t(kDeBitpack02 + b - 2); cost = 16 * b; }
  } else {
    if (svb_raw < cost) { choice = kDeSvb; cost = svb_raw; }
    if (svb_gap < cost) { choice = kDeDeltaSvb; cost = svb_gap; }
  }
  const uint32_t span = docs[len - 1] - prev;  // bit index of the last doc relative to prev
  const uint32_t words = (span + 1 + 63) / 64;
  if (1 + 8 * words - 2 < cost) { choice = kDeBitset; cost = 1 + 8 * words; }

  out.put(choice);
  if (choice == kDeValues) {
    out.put_bytes(docs, size_t(len) * 4);
  } else if (choice == kDeBitset) {
    std::vector<uint64_t> bits(words, 0);
    for (uint32_t i = 0; i < len; ++i) { const uint32_t o = docs[i] - prev; bits[o >> 6] |= uint64_t(1) << (o & 63); }
    out.put(uint8_t(words));
    out.put_bytes(bits.data(), size_t(words) * 8);
  } else if (choice == kDeSvb || choice == kDeDeltaSvb) {
    const size_t at = out.size();
    out.put16(0);
    const size_t n = svb_put(out, choice == kDeSvb ? docs : gaps, len);
    out.buf[at] = uint8_t(n); out.buf[at + 1] = uint8_t(n >> 8);
  } else {
    const uint32_t b = uint32_t(choice - kDeBitpack02) + 2;
    uint32_t packed[4 * 32];
    vertical_pack(gaps, b, packed);
    out.put_bytes(packed, size_t(b) * 16);
  }
}

void encode_freq_block(ByteWriter& out, const uint32_t* f, uint32_t len) {
  uint32_t mx = 0, svb = 2 + (len + 3) / 4;
  bool uniform = true;
  for (uint32_t i = 0; i < len; ++i) { uniform &= f[i] == f[0]; mx = std::max(mx, f[i]); svb += nbytes_1234(f[i]); }
  if (uniform) {
    if (mx <= 0xFF) { out.put(kESame08); out.put(uint8_t(mx)); }
    else if (mx <= 0xFFFF) { out.put(kESame16); out.put16(mx); }
    else { out.put(kESame32); out.put32(mx); }
    return;
  }
  uint8_t choice = kEValues;
  uint32_t cost = len * 4;
  if (len == kBlockSize) {
    const uint32_t b = width_of(mx);
    if (16 * b < cost) { choice = uint8_t(kEBitpack01 + b - 1); cost = 16 * b; }
  } else if (svb < cost) {
    choice = kESvb;
  }
  out.put(choice);
  if (choice == kEValues) {
    out.put_bytes(f, size_t(len) * 4);
  } else if (choice == kESvb) {
    const size_t at = out.size();
    out.put16(0);
    const size_t n = svb_put(out, f, len);
    out.buf[at] = uint8_t(n); out.buf[at + 1] = uint8_t(n >> 8);
  } else {
    const uint32_t b = uint32_t(choice - kEBitpack01) + 1;
    uint32_t packed[4 * 32];
    vertical_pack(f, b, packed);
    out.put_bytes(packed, size_t(b) * 16);
  }
}

size_t decode_doc_block(const uint8_t* p, const uint8_t* end, uint32_t len, uint32_t prev, uint32_t* out) {
  const size_t payload = doc_payload_bytes(p, end, len);
  if (payload == SIZE_MAX || p + 1 + payload > end) return 0;
  const uint8_t e = p[0];
  const uint8_t* d = p + 1;
  switch (e) {
    case kDeValues: std::memcpy(out, d, size_t(len) * 4); break;
    case kDeSame08: case kDeSame16: case kDeSame32: {
      const uint32_t g = e == kDeSame08 ? d[0] : e == kDeSame16 ? load16(d) : load32(d);
      for (uint32_t i = 0; i < len; ++i) out[i] = prev + g * (i + 1);
    } break;
    case kDeBitset: {
      uint32_t n = 0;
      for (uint32_t w = 0; w < d[0]; ++w) {
        uint64_t x; std::memcpy(&x, d + 1 + 8 * w, 8);
        for (; x; x &= x - 1) { if (n == len) return 0; out[n++] = prev + 64 * w + uint32_t(__builtin_ctzll(x)); }
      }
      if (n != len) return 0;
    } break;
    case kDeSvb: case kDeDeltaSvb: {
      if (!svb_get(d + 2, d + payload, len, out)) return 0;
      if (e == kDeDeltaSvb) for (uint32_t i = 0, a = prev; i < len; ++i) { a += out[i]; out[i] = a; }
    } break;
    default: {
      vertical_unpack(d, uint32_t(e - kDeBitpack02) + 2, out);
      for (uint32_t i = 0, a = prev; i < len; ++i) { a += out[i]; out[i] = a; }
    }
  }
  return 1 + payload;
}

size_t decode_freq_block(const uint8_t* p, const uint8_t* end, uint32_t len, uint32_t* out) {
  const size_t payload = freq_payload_bytes(p, end, len);
  if (payload == SIZE_MAX || p + 1 + payload > end) return 0;
  const uint8_t e = p[0];
  const uint8_t* d = p + 1;
  switch (e) {
    case kEValues: std::memcpy(out, d, size_t(len) * 4); break;
    case kESame08: std::fill(out, out + len, uint32_t(d[0])); break;
    case kESame16: std::fill(out, out + len, load16(d)); break;
    case kESame32: std::fill(out, out + len, load32(d)); break;
    case kESvb: if (!svb_get(d + 2, d + payload, len, out)) return 0; break;
    default: vertical_unpack(d, uint32_t(e - kEBitpack01) + 1, out);
  }
  return 1 + payload;
}

// ------------------------------------------------------------------------------------------
// PostingWriter
// ------------------------------------------------------------------------------------------
PostingWriter::PostingWriter(uint32_t segment_docs, bool has_wand, float wand_b, const uint32_t* norms)
    : segment_docs_(segment_docs), has_wand_(has_wand), b_(wand_b) {
  if (norms) {
    norms_.assign(norms, norms + segment_docs);
    uint64_t sum = 0, nz = 0;
    for (uint32_t v : norms_) { sum += v; nz += v != 0; }
    avg_dl_ = nz ? float(double(sum) / double(nz)) : 0.f;  // NormReader::GetAvg, formats/norm_reader_impl.hpp:83-88
    norms_ptr_ = norms_.data();
  }
}

PostingWriter::PostingWriter(uint32_t segment_docs, bool has_wand, float wand_b, const uint32_t* borrowed_norms, float avg_dl)
    : segment_docs_(segment_docs), has_wand_(has_wand), b_(wand_b), avg_dl_(avg_dl), norms_ptr_(borrowed_norms) {}

// FreqNormProducer<kWandTagAvgDL>::ProduceBM25 / CmpBm25 (wand_writer.hpp:142-175, 302-311):
// keep the (freq, norm) pair with the larger tf / (k(1-b+b*dl/avgdl) + tf); replace only when
// strictly better.
void PostingWriter::feed(uint32_t freq, uint32_t norm, MaxPair& to) const {
  const float x = (1.f - b_) * avg_dl_;
  const float mine = float(freq) * (x + b_ * float(to.norm));
  const float theirs = float(to.freq) * (x + b_ * float(norm));
  if (mine <= theirs) return;
  to.freq = freq; to.norm = norm;
}
void PostingWriter::fold(const MaxPair& from, MaxPair& to) const { feed(from.freq, from.norm, to); }

namespace {
void put_pair(ByteWriter& w, const MaxPair& m) {  // Write/Size, wand_writer.hpp:196-218
  const uint32_t extra = m.norm != m.freq ? vint_len(m.norm - m.freq) : 0;
  w.put(uint8_t(vint_len(m.freq) + extra));
  w.put_vint(m.freq);
  if (extra) w.put_vint(m.norm - m.freq);
}
uint32_t skip_levels_for(uint64_t segment_docs) {  // CountMaxLevels, skip_list.cpp:38-41
  if (segment_docs <= kBlockSize) return 0;
  uint32_t levels = 1;
  for (uint64_t x = segment_docs / kBlockSize; x >= kSkipN; x /= kSkipN) ++levels;
  return std::min(levels, kMaxSkipLevels);
}
}  // namespace

void PostingWriter::add_term(const uint32_t* docs, const uint32_t* freqs, uint32_t n) {
  TermMeta meta;
  meta.doc_start = out_.size();
  meta.docs_count = n;
  if (n == 0) { terms_.push_back(meta); return; }
  uint64_t tf = 0;
  for (uint32_t i = 0; i < n; ++i) tf += freqs[i];
  meta.freq = uint32_t(tf);
  if (n == 1) {  // inline in the term meta; nothing goes to ".doc" (writer.hpp:457-458)
    meta.e_skip_start = docs[0] - 1;
    terms_.push_back(meta);
    return;
  }

  const uint32_t max_levels = skip_levels_for(segment_docs_);
  ByteWriter level[kMaxSkipLevels];
  uint64_t level_ptr[kMaxSkipLevels];
  std::fill(level_ptr, level_ptr + kMaxSkipLevels, meta.doc_start);
  MaxPair best[kMaxSkipLevels + 1];  // per-level running block-max; [top] collects the whole list

  // An entry for block j is emitted when the first doc of block j+1 arrives (writer.hpp:736-749):
  // level 0 every 128 docs, level i every 128*32^i docs (skip_list.hpp:93-118).
  auto emit_entries = [&](uint32_t blocks_done, uint32_t block_last) {
    uint64_t child = 0;
    uint32_t c = blocks_done;
    for (uint32_t lv = 0; lv < max_levels; ++lv) {
      if (lv > 0) { if (c % kSkipN) break; c /= kSkipN; }
      ByteWriter& s = level[lv];
      s.put_vint(block_last);                               // WriteSkip, writer.hpp:305-316
      s.put_vint(uint64_t(out_.size()) - level_ptr[lv]);
      level_ptr[lv] = out_.size();
      if (has_wand_) { put_pair(s, best[lv]); fold(best[lv], best[lv + 1]); best[lv] = MaxPair{}; }
      const uint64_t here = s.size();
      if (lv > 0) s.put_vint(child);
      child = here;
    }
  };

  const uint32_t full = n / kBlockSize, tail = n % kBlockSize;
  if (n < kBlockSize && has_wand_) {  // short list: list maximum first (EndTerm, writer.hpp:460-462)
    for (uint32_t i = 0; i < n; ++i) feed(freqs[i], norm_of(docs[i]), best[0]);
    put_pair(out_, best[0]);
  }
  uint32_t prev_last = 0;
  for (uint32_t blk = 0; blk < full; ++blk) {
    const uint32_t* d = docs + size_t(blk) * kBlockSize;
    const uint32_t* f = freqs + size_t(blk) * kBlockSize;
    if (blk > 0) emit_entries(blk, prev_last);
    encode_doc_block(out_, d, kBlockSize, prev_last);
    encode_freq_block(out_, f, kBlockSize);
    if (has_wand_) for (uint32_t i = 0; i < kBlockSize; ++i) feed(f[i], norm_of(d[i]), best[0]);
    prev_last = d[kBlockSize - 1];
  }
  if (n == kBlockSize && has_wand_) put_pair(out_, best[0]);  // exactly one block: maximum after it
  if (tail && n > kBlockSize) {
    const uint32_t* d = docs + size_t(full) * kBlockSize;
    const uint32_t* f = freqs + size_t(full) * kBlockSize;
    emit_entries(full, prev_last);
    encode_doc_block(out_, d, tail, prev_last);
    encode_freq_block(out_, f, tail);
    if (has_wand_) for (uint32_t i = 0; i < tail; ++i) feed(f[i], norm_of(d[i]), best[0]);
  } else if (tail) {
    encode_doc_block(out_, docs, tail, 0);
    encode_freq_block(out_, freqs, tail);
  }
  if (n > kBlockSize) {  // skip data (EndTerm :471-476, FlushLevels skip_list.cpp:77-94)
    meta.e_skip_start = out_.size() - meta.doc_start;
    uint32_t used = 0;
    for (uint32_t lv = 0; lv < max_levels; ++lv) if (level[lv].size()) used = lv + 1;
    if (has_wand_) {
      for (uint32_t lv = 0; lv < used; ++lv) fold(best[lv], best[lv + 1]);  // SizeRoot, wand_writer.hpp:94-102
      put_pair(out_, best[used]);
    }
    out_.put_vint(used);
    for (uint32_t lv = used; lv-- > 0;) { out_.put_vint(uint64_t(level[lv].size())); out_.put_bytes(level[lv].buf.data(), level[lv].size()); }
  }
  terms_.push_back(meta);
}

void PostingWriter::append(const PostingWriter& other) {
  const uint64_t base = out_.size();
  out_.put_bytes(other.out_.buf.data(), other.out_.size());
  for (TermMeta m : other.terms_) { m.doc_start += base; terms_.push_back(m); }
}

// ------------------------------------------------------------------------------------------
// Staging: ".doc" stream -> block descriptors + 16-byte aligned payload arena.
// ------------------------------------------------------------------------------------------
namespace {

struct Arena {
  std::vector<uint8_t>& a;
  uint32_t push(const uint8_t* p, size_t n) {  // returns offset in 16-byte units
    const size_t at = a.size();
    a.insert(a.end(), p, p + n);
    a.resize((a.size() + 15) & ~size_t(15), 0);
    return uint32_t(at / 16);
  }
};

bool read_pair(const uint8_t*& p, const uint8_t* end, MaxPair* m) {  // FreqNormSource::Read, wand_writer.hpp:366-381
  if (p >= end) return false;
  const uint32_t size = *p++;
  if (p + size > end) return false;
  const uint8_t* q = p;
  m->freq = get_vint<uint32_t>(q, p + size);
  m->norm = m->freq;
  if (q < p + size) m->norm += get_vint<uint32_t>(q, p + size);
  p += size;
  return true;
}

}  // namespace

std::string stage_postings(const uint8_t* doc, size_t n, const TermMeta* terms, size_t n_terms, bool has_wand,
                           StagedPostings* sp) {
  sp->arena.clear(); sp->blocks.clear(); sp->blk_max.clear(); sp->blk_anchor.clear(); sp->term_max.clear();
  sp->term_blk_begin.assign(1, 0); sp->term_docs.clear(); sp->term_bytes.clear(); sp->term_probe.clear();
  sp->n_postings = 0; sp->has_wand = has_wand;
  Arena arena{sp->arena};
  const uint8_t* const end = doc + n;
  std::vector<uint32_t> scratch(kBlockSize);

  for (size_t t = 0; t < n_terms; ++t) {
    const TermMeta& m = terms[t];
    const uint32_t cnt = m.docs_count;
    sp->term_docs.push_back(cnt);
    sp->n_postings += cnt;
    MaxPair root{0, 0};
    uint64_t enc_bytes = 0;
    uint32_t direct_blocks = 0;   // blocks a probe can answer without a decode
    const size_t first_block = sp->blocks.size();
    if (cnt == 1) {
      // Single-doc terms live in the term meta (iterator_score.hpp:1015-1030); give them one raw block.
      const uint32_t d = uint32_t(m.e_skip_start) + 1, f = m.freq;
      BlockDesc b;
      b.off16 = arena.push(reinterpret_cast<const uint8_t*>(&d), 4);
      arena.push(reinterpret_cast<const uint8_t*>(&f), 4);
      b.last_doc = d; b.prev_last = 0;
      b.packed = pack_desc(kDeValues, kEValues, 1, 1, 0);
      sp->blocks.push_back(b);
      sp->blk_max.push_back(MaxPair{0, 0});  // unknown: resolved to "no bound" by the caller
      for (uint32_t a = 0; a < 4; ++a) sp->blk_anchor.push_back(a == 3 ? 0u : 0xFFFFFFFFu);
    } else if (cnt > 1) {
      if (m.doc_start > n) return "term " + std::to_string(t) + ": doc_start beyond stream";
      const uint8_t* p = doc + m.doc_start;
      if (has_wand && cnt < kBlockSize && !read_pair(p, end, &root)) return "truncated block-max entry";
      const uint32_t nblk = (cnt + kBlockSize - 1) / kBlockSize;
      uint32_t prev_last = 0;
      std::vector<uint64_t> blk_start(nblk + 1);
      for (uint32_t j = 0; j < nblk; ++j) {
        const uint32_t len = std::min(kBlockSize, cnt - j * kBlockSize);
        blk_start[j] = uint64_t(p - doc);
        const size_t dsz = doc_payload_bytes(p, end, len);
        if (dsz == SIZE_MAX || p + 1 + dsz > end) return "term " + std::to_string(t) + ": bad doc block header";
        const uint8_t denc = p[0];
        const uint8_t* dpay = p + 1;
        size_t dcopy = dsz;
        uint32_t words = 0;
        if (denc == kDeBitset) { words = dpay[0]; dpay += 1; dcopy -= 1; if (words > 64) return "bitset too wide"; }
        else if (denc == kDeSvb || denc == kDeDeltaSvb) { dpay += 2; dcopy -= 2; }
        const uint8_t* fp = p + 1 + dsz;
        const size_t fsz = freq_payload_bytes(fp, end, len);
        if (fsz == SIZE_MAX || fp + 1 + fsz > end) return "term " + std::to_string(t) + ": bad freq block header";
        const uint8_t fenc = fp[0];
        const uint8_t* fpay = fp + 1;
        size_t fcopy = fsz;
        if (fenc == kESvb) { fpay += 2; fcopy -= 2; }
        BlockDesc b;
        b.off16 = arena.push(dpay, dcopy);
        const uint32_t foff = arena.push(fpay, fcopy);
        b.prev_last = prev_last;
        b.last_doc = 0;  // filled below
        b.packed = pack_desc(denc, fenc, len, foff - b.off16, words);
        // Last doc of the block: needed as the delta base of the next block and for window lookup.
        // all-same / raw / bitset give it in O(1); bit-packed and svb blocks are decoded once here.
        if (decode_doc_block(p, end, len, prev_last, scratch.data()) == 0) return "term " + std::to_string(t) + ": undecodable block";
        b.last_doc = scratch[len - 1];
        if (b.last_doc <= prev_last) return "term " + std::to_string(t) + ": doc ids not ascending";
        prev_last = b.last_doc;
        sp->blocks.push_back(b);
        sp->blk_max.push_back(MaxPair{0, 0});
        for (uint32_t a = 0; a < 3; ++a) sp->blk_anchor.push_back(32u * a + 31u < len ? scratch[32u * a + 31u] : 0xFFFFFFFFu);
        sp->blk_anchor.push_back(0u);
        if (denc == kDeBitset && fenc != kESvb) ++direct_blocks;
        p = fp + 1 + fsz;
        enc_bytes += 2 + dsz + fsz;
      }
      blk_start[nblk] = uint64_t(p - doc);
      if (cnt == kBlockSize && has_wand && !read_pair(p, end, &root)) return "truncated block-max entry";
      if (cnt > kBlockSize) {
        // Skip data: [u8 size][root pair] [vint levels] then levels top-down, each [vlong len][bytes]
        // (skip_list.cpp:77-94). Only level 0 is needed: one entry per block that has a successor.
        const uint8_t* s = doc + m.doc_start + m.e_skip_start;
        if (s != p) return "term " + std::to_string(t) + ": skip data not where the blocks end";
        if (has_wand && !read_pair(s, end, &root)) return "truncated root block-max";
        const uint32_t levels = get_vint<uint32_t>(s, end);
        if (levels == 0 || levels > kMaxSkipLevels) return "bad skip level count";
        const uint8_t* l0 = nullptr; uint64_t l0_len = 0;
        for (uint32_t lv = levels; lv-- > 0;) {
          const uint64_t len = get_vint<uint64_t>(s, end);
          if (s + len > end) return "truncated skip level";
          if (lv == 0) { l0 = s; l0_len = len; }
          s += len;
        }
        const uint8_t* q = l0; const uint8_t* qe = l0 + l0_len;
        uint64_t ptr = m.doc_start;
        for (uint32_t j = 0; j + 1 < nblk; ++j) {
          if (q >= qe) return "term " + std::to_string(t) + ": level-0 skip entries missing";
          const uint32_t last = get_vint<uint32_t>(q, qe);
          ptr += get_vint<uint64_t>(q, qe);
          if (last != sp->blocks[first_block + j].last_doc || ptr != blk_start[j + 1]) return "term " + std::to_string(t) + ": skip entry disagrees with blocks";
          if (has_wand && !read_pair(q, qe, &sp->blk_max[first_block + j])) return "truncated level-0 block-max";
        }
      }
      if (has_wand) for (size_t j = first_block; j < sp->blocks.size(); ++j) if (sp->blk_max[j].freq == 0) sp->blk_max[j] = root;
    }
    sp->term_max.push_back(root);
    sp->term_bytes.push_back(enc_bytes);
    {
      const size_t nb = sp->blocks.size() - first_block;
      sp->term_probe.push_back(nb >= 8 && direct_blocks * 16 >= nb * 15 ? 1 : 0);   // >= 15/16 of the blocks
    }
    sp->term_blk_begin.push_back(uint32_t(sp->blocks.size()));
  }
  sp->arena.resize(sp->arena.size() + 1024, 0);  // slack so 16-byte over-reads of the last block stay in bounds
  return "";
}

}  // namespace sdbg
