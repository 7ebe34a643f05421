// bm25_kernels.cuh -- sm_100a kernels for the BM25 posting scan + top-k.
//
// Reference behaviour being reproduced (paths relative to /root/reference/libs/iresearch/include/iresearch):
//   decode      formats/posting/format_block_128.hpp:475-636 (ReadTailDelta / ReadTail)
//   norms       search/column_collector.hpp:52-74, formats/column/norm_column_reader.hpp:99-108
//   score       search/bm25.cpp:90-107  (c1 = nc + nl*norm; r = c0 - c0*c1/(c1+freq), no FMA contraction)
//   windows     search/max_score_iterator.hpp:311-356 (score window + bitmap, Sum merge)
//   collector   index/iterators.hpp:103-250 (buffer of 2k, select when full, threshold = k-th)
//   conjunction search/conjunction.hpp:248-505 (all terms must match), score = sum of sub-scores
//   hybrid      index/table_filter_iterator.cpp:381-475 (column predicate on matched docs)
//
// Mapping: one CTA owns a chain of doc-id windows of one query; within a window every warp decodes
// one 128-posting block at a time (lane l holds postings 4l..4l+3 -- the four 32-bit lanes of the
// simdcomp layout are one 16-byte vector per row, so a lane reads at most two uint4 per block),
// scores it and adds into a shared-memory score window; matched slots are then compared with the
// query's running threshold and appended to a per-CTA candidate buffer that is compacted by a
// CTA-wide bitonic sort when full (the GPU analogue of nth_element at 2k).
#pragma once

#include "device_common.cuh"

namespace sdbg {

// ---- device views ----
struct PostingsDev {
  const uint4* arena;     // block payloads, 16-byte units
  const uint4* blocks;    // BlockDesc {off16, last_doc, prev_last, packed}
  const uint2* blk_max;   // {freq, norm} per block (block-max pairs), may be null
  const uint8_t* norms;   // fixed-width field lengths, row = doc-1; null => norm = 1
  uint32_t norm_width;    // 1, 2 or 4
  uint32_t n_docs;
  uint32_t ordinal_base;  // first global ordinal of this segment (keys carry base + doc)
};

struct FilterDev {  // one pushed column predicate for the hybrid path
  const void* values;        // null => no filter
  const uint64_t* validity;  // null => NOT NULL column
  int32_t type;              // 0 i64, 1 f64, 2 i32
  int32_t op;
  int64_t lo_i, hi_i;
  double lo_f, hi_f;
};

struct QTermDev {  // one term of one query, 32 bytes
  uint32_t blk_begin;  // first BlockDesc of the term
  uint32_t nblk;
  float c0;            // boost*(k1+1)*idf   (bm25.cpp:224)
  float norm_const;
  float norm_length;
  uint32_t docs_count;
  uint32_t pad0, pad1;
};

constexpr uint32_t kMaxQueryTerms = 16;
constexpr uint32_t kTopkThreads = 256;
constexpr uint32_t kTopkWarps = kTopkThreads / 32;

__device__ __forceinline__ uint32_t desc_doc_enc(uint32_t p) { return p & 63u; }
__device__ __forceinline__ uint32_t desc_freq_enc(uint32_t p) { return (p >> 6) & 63u; }
__device__ __forceinline__ uint32_t desc_len(uint32_t p) { return ((p >> 12) & 127u) + 1u; }
__device__ __forceinline__ uint32_t desc_fdelta(uint32_t p) { return (p >> 19) & 63u; }
__device__ __forceinline__ uint32_t desc_words(uint32_t p) { return p >> 25; }

// ---- block decode: lane holds values 4*lane .. 4*lane+3 ----
// Bit-packed payload (simdunpack layout): row = lane, the four 32-bit lanes of that row are the
// four components of one uint4; row r occupies bits [r*b, r*b+b) of every lane stream.
__device__ __forceinline__ void unpack4(const uint4* p, uint32_t b, uint32_t lane, uint32_t v[4]) {
  const uint32_t bit = lane * b;
  const uint32_t w = bit >> 5, sh = bit & 31u;
  const uint4 lo = ld_ro_v4(p + w);
  const uint4 hi = ld_ro_v4(p + min(w + 1u, b - 1u));
  const uint32_t mask = (1u << b) - 1u;  // b <= 31
  v[0] = __funnelshift_r(lo.x, hi.x, sh) & mask;
  v[1] = __funnelshift_r(lo.y, hi.y, sh) & mask;
  v[2] = __funnelshift_r(lo.z, hi.z, sh) & mask;
  v[3] = __funnelshift_r(lo.w, hi.w, sh) & mask;
}

// StreamVByte 1234 (tails only): control byte `lane` describes this lane's four values.
__device__ __forceinline__ void svb4(const uint4* p, uint32_t len, uint32_t lane, uint32_t v[4]) {
  const uint8_t* bytes = reinterpret_cast<const uint8_t*>(p);
  const uint32_t nctl = (len + 3u) >> 2;
  const uint32_t ctl = lane < nctl ? uint32_t(__ldg(bytes + lane)) : 0u;
  uint32_t n[4], mine = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    n[j] = (4u * lane + j < len) ? ((ctl >> (2 * j)) & 3u) + 1u : 0u;
    mine += n[j];
  }
  uint32_t pos = nctl + warp_incl_scan(mine, lane) - mine;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t x = 0;
    for (uint32_t k = 0; k < n[j]; ++k) x |= uint32_t(__ldg(bytes + pos + k)) << (8 * k);
    pos += n[j];
    v[j] = x;
  }
}

__device__ __forceinline__ uint32_t same_value(const uint4* p, uint32_t width_code /*1,2,3*/) {
  const uint32_t raw = __ldg(reinterpret_cast<const uint32_t*>(p));
  return width_code == 1 ? (raw & 0xFFu) : width_code == 2 ? (raw & 0xFFFFu) : raw;
}

// Turns four per-lane gaps into absolute ids: running sum across the warp in value order.
__device__ __forceinline__ void prefix_from_gaps(uint32_t prev, uint32_t lane, uint32_t v[4]) {
  v[1] += v[0]; v[2] += v[1]; v[3] += v[2];
  const uint32_t base = prev + warp_incl_scan(v[3], lane) - v[3];
  v[0] += base; v[1] += base; v[2] += base; v[3] += base;
}

// Doc ids of one block. `stage` = 128 u32 of per-warp shared scratch (bitset rank scatter).
__device__ __forceinline__ void decode_docs(const uint4* arena, const uint4& d, uint32_t lane,
                                            uint32_t* stage, uint32_t doc[4]) {
  const uint4* p = arena + d.x;
  const uint32_t enc = desc_doc_enc(d.w), len = desc_len(d.w), prev = d.z;
  if (enc >= 8u) {  // de_delta_bitpack_b, b = enc - 6
    unpack4(p, enc - 6u, lane, doc);
    prefix_from_gaps(prev, lane, doc);
  } else if (enc == 4u) {  // de_for_bitset: bit j set => id prev + j
    const uint32_t words = desc_words(d.w);
    unsigned long long w0 = 0, w1 = 0;
    if (2u * lane < words) {
      const uint4 x = ld_ro_v4(p + lane);
      w0 = (static_cast<unsigned long long>(x.y) << 32) | x.x;
      if (2u * lane + 1u < words) w1 = (static_cast<unsigned long long>(x.w) << 32) | x.z;
    }
    const uint32_t c = __popcll(w0) + __popcll(w1);
    uint32_t r = warp_incl_scan(c, lane) - c;
    const uint32_t id0 = prev + 128u * lane;
    for (; w0; w0 &= w0 - 1) stage[r++] = id0 + uint32_t(__ffsll(static_cast<long long>(w0)) - 1);
    for (; w1; w1 &= w1 - 1) stage[r++] = id0 + 64u + uint32_t(__ffsll(static_cast<long long>(w1)) - 1);
    __syncwarp();
    const uint4 o = reinterpret_cast<const uint4*>(stage)[lane];
    doc[0] = o.x; doc[1] = o.y; doc[2] = o.z; doc[3] = o.w;
    __syncwarp();
  } else if (enc >= 1u && enc <= 3u) {  // de_delta_all_same_{08,16,32}
    const uint32_t g = same_value(p, enc);
#pragma unroll
    for (int j = 0; j < 4; ++j) doc[j] = prev + g * (4u * lane + j + 1u);
  } else if (enc == 0u) {  // de_values
    uint4 x = make_uint4(0, 0, 0, 0);
    if (4u * lane < len) x = ld_ro_v4(p + lane);
    doc[0] = x.x; doc[1] = x.y; doc[2] = x.z; doc[3] = x.w;
  } else {  // 5 de_streamvbyte1234, 7 de_delta_streamvbyte1234 (tails)
    svb4(p, len, lane, doc);
    if (enc == 7u) prefix_from_gaps(prev, lane, doc);
  }
}

__device__ __forceinline__ void decode_freqs(const uint4* arena, const uint4& d, uint32_t lane, uint32_t f[4]) {
  const uint4* p = arena + d.x + desc_fdelta(d.w);
  const uint32_t enc = desc_freq_enc(d.w), len = desc_len(d.w);
  if (enc >= 5u) {  // e_bitpack_b, b = enc - 4
    unpack4(p, enc - 4u, lane, f);
  } else if (enc >= 1u && enc <= 3u) {
    const uint32_t g = same_value(p, enc);
    f[0] = f[1] = f[2] = f[3] = g;
  } else if (enc == 0u) {
    uint4 x = make_uint4(0, 0, 0, 0);
    if (4u * lane < len) x = ld_ro_v4(p + lane);
    f[0] = x.x; f[1] = x.y; f[2] = x.z; f[3] = x.w;
  } else {
    svb4(p, len, lane, f);
  }
}

__device__ __forceinline__ uint32_t load_norm(const uint8_t* norms, uint32_t width, uint32_t doc) {
  if (norms == nullptr) return 1u;  // bm25.cpp:353-360
  const size_t row = size_t(doc) - 1u;
  if (width == 1u) return __ldg(norms + row);
  if (width == 2u) return __ldg(reinterpret_cast<const uint16_t*>(norms) + row);
  return __ldg(reinterpret_cast<const uint32_t*>(norms) + row);
}

// bm25.cpp:105-106 with the reference's operation order; intrinsics forbid FMA contraction so the
// result is bit-identical to the g++ -ffp-contract=off oracle.
__device__ __forceinline__ float bm25(uint32_t freq, uint32_t norm, float c0, float nc, float nl) {
  const float c1 = __fadd_rn(nc, __fmul_rn(nl, static_cast<float>(norm)));
  return __fsub_rn(c0, __fdiv_rn(__fmul_rn(c0, c1), __fadd_rn(c1, static_cast<float>(freq))));
}

__device__ __forceinline__ bool filter_pass(const FilterDev& f, uint32_t doc) {
  if (f.values == nullptr) return true;
  const size_t r = size_t(doc) - 1u;  // row = doc - 1 (index/column_extract.hpp:46-48)
  const bool valid = f.validity == nullptr || ((f.validity[r >> 6] >> (r & 63)) & 1ull);
  if (f.op == 7) return !valid;
  if (f.op == 8) return valid;
  if (!valid) return false;
  if (f.type == 1) {
    const double v = __ldg(static_cast<const double*>(f.values) + r);
    switch (f.op) {
      case 0: return v < f.lo_f; case 1: return v <= f.lo_f; case 2: return v > f.lo_f;
      case 3: return v >= f.lo_f; case 4: return v == f.lo_f; case 5: return v != f.lo_f;
      default: return v >= f.lo_f && v <= f.hi_f;
    }
  }
  const long long v = f.type == 2 ? static_cast<long long>(__ldg(static_cast<const int*>(f.values) + r))
                                  : __ldg(static_cast<const long long*>(f.values) + r);
  switch (f.op) {
    case 0: return v < f.lo_i; case 1: return v <= f.lo_i; case 2: return v > f.lo_i;
    case 3: return v >= f.lo_i; case 4: return v == f.lo_i; case 5: return v != f.lo_i;
    default: return v >= f.lo_i && v <= f.hi_i;
  }
}

// ------------------------------------------------------------------------------------------
// Probe kernel: decode + score one whole posting list (exhaustive). One warp per block.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTopkThreads)
decode_score_kernel(PostingsDev seg, uint32_t blk_begin, uint32_t nblk, float c0, float nc, float nl,
                    uint32_t* __restrict__ docs, uint32_t* __restrict__ freqs, float* __restrict__ scores) {
  __shared__ __align__(16) uint32_t stage[kTopkWarps][128];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  for (uint32_t b = blockIdx.x * kTopkWarps + warp; b < nblk; b += gridDim.x * kTopkWarps) {
    const uint4 d = ld_ro_v4(seg.blocks + blk_begin + b);
    uint32_t doc[4], f[4];
    decode_docs(seg.arena, d, lane, stage[warp], doc);
    decode_freqs(seg.arena, d, lane, f);
    const uint32_t len = desc_len(d.w);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t i = 4u * lane + j;
      if (i < len) {
        const size_t o = size_t(b) * 128u + i;
        docs[o] = doc[j]; freqs[o] = f[j];
        scores[o] = bm25(f[j], load_norm(seg.norms, seg.norm_width, doc[j]), c0, nc, nl);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Fused scan + score + top-k kernel.
// grid = (chains G, queries Q); CTA (g, q) handles windows g, g+G, ... of query q.
// Shared memory (dynamic): acc[W] f32 | cnt[W] u8 (AND only) | mask[W/32] u32 | cand[cap] u64.
// ------------------------------------------------------------------------------------------
struct TopkParams {
  PostingsDev seg;
  FilterDev filt;
  const QTermDev* qterms;      // flattened, per query sorted by ascending docs_count
  const uint32_t* qterm_off;   // n_queries + 1
  unsigned long long* theta;   // per query running threshold key (shared by all chains / segments)
  unsigned long long* total;   // per query matched-doc count
  unsigned long long* cand;    // [Q][lists][cap] candidate keys, sorted descending on exit
  uint32_t* cand_n;            // [Q][lists]
  uint32_t lists;              // candidate lists per query = segments * chains
  uint32_t list_base;          // this segment's first list
  uint32_t W;                  // window size in docs, multiple of 64
  uint32_t n_windows;
  uint32_t k;
  uint32_t cap;                // candidate buffer capacity, power of two, > k
  int32_t conjunction;         // 0 OR, 1 AND
};

// acc index swizzle: at emission a thread owns one mask word => 32 consecutive slots; XOR-ing the
// low five bits with the word index spreads a warp's reads over all banks.
__device__ __forceinline__ uint32_t swz(uint32_t i) { return i ^ ((i >> 5) & 31u); }

__global__ void __launch_bounds__(kTopkThreads)
bm25_topk_kernel(const TopkParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* acc = reinterpret_cast<float*>(smem_raw);
  uint8_t* cnt = reinterpret_cast<uint8_t*>(acc + P.W);
  uint32_t* mask = reinterpret_cast<uint32_t*>(cnt + (P.conjunction ? P.W : 0u));
  unsigned long long* cand = reinterpret_cast<unsigned long long*>(mask + P.W / 32u);

  __shared__ __align__(16) uint32_t stage[kTopkWarps][128];
  __shared__ uint32_t s_first[2][kMaxQueryTerms];   // first overlapping block per term (double buffered)
  __shared__ uint32_t s_prefix[2][kMaxQueryTerms + 1];
  __shared__ uint32_t s_cursor[kMaxQueryTerms];
  __shared__ uint32_t s_ncand, s_matched;
  __shared__ unsigned long long s_theta;

  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t q = blockIdx.y, g = blockIdx.x, G = gridDim.x;
  const uint32_t t0 = P.qterm_off[q];
  const uint32_t T = min(P.qterm_off[q + 1] - t0, kMaxQueryTerms);
  const QTermDev* qt = P.qterms + t0;
  const uint32_t words = P.W / 32u;

  for (uint32_t i = tid; i < P.W; i += blockDim.x) acc[i] = 0.f;
  if (P.conjunction) for (uint32_t i = tid; i < P.W / 4u; i += blockDim.x) reinterpret_cast<uint32_t*>(cnt)[i] = 0u;
  for (uint32_t i = tid; i < words; i += blockDim.x) mask[i] = 0u;
  for (uint32_t i = tid; i < P.cap; i += blockDim.x) cand[i] = 0ull;
  if (tid < kMaxQueryTerms) s_cursor[tid] = 0u;
  if (tid == 0) { s_ncand = 0u; s_matched = 0u; s_theta = 0ull; }
  __syncthreads();

  // Block range of term `lane` for window w: blocks whose [prev_last+1, last_doc] meets [lo, hi).
  // Galloping search from the term's cursor (the windows of a chain only move forward).
  auto plan_window = [&](uint32_t w, uint32_t buf) {
    if (warp != 0) return;
    uint32_t n = 0;
    if (lane < T && w < P.n_windows) {
      const uint32_t lo = 1u + w * P.W;
      const unsigned long long hi_last64 = static_cast<unsigned long long>(lo) + P.W - 1ull;
      const uint32_t hi_last = hi_last64 > 0xFFFFFFFFull ? 0xFFFFFFFFu : uint32_t(hi_last64);  // last doc of the window
      const uint4* B = P.seg.blocks + qt[lane].blk_begin;
      const uint32_t nblk = qt[lane].nblk;
      uint32_t a = s_cursor[lane], step = 1u;  // first block with last_doc >= lo
      while (a + step <= nblk && __ldg(&B[a + step - 1u].y) < lo) { a += step; step <<= 1; }
      uint32_t l = a, r = min(a + step - 1u, nblk);
      while (l < r) { const uint32_t m = (l + r) >> 1; if (__ldg(&B[m].y) < lo) l = m + 1u; else r = m; }
      const uint32_t first = l;
      s_cursor[lane] = first;
      uint32_t e = first; step = 1u;           // first block that starts after the window
      while (e + step <= nblk && __ldg(&B[e + step - 1u].z) < hi_last) { e += step; step <<= 1; }
      l = e; r = min(e + step - 1u, nblk);
      while (l < r) { const uint32_t m = (l + r) >> 1; if (__ldg(&B[m].z) < hi_last) l = m + 1u; else r = m; }
      n = l - first;
      s_first[buf][lane] = first;
    }
    const uint32_t incl = warp_incl_scan(n, lane);
    if (lane < kMaxQueryTerms) s_prefix[buf][lane + 1] = incl;
    if (lane == 0) s_prefix[buf][0] = 0u;
  };

  // Sort the candidate buffer, keep the best k, raise the thresholds. All threads call it.
  auto compact = [&]() {
    block_sort_desc(cand, P.cap);
    if (tid == 0) {
      const uint32_t have = min(s_ncand, P.cap);
      if (have > P.k) {
        const unsigned long long kth = cand[P.k - 1u];
        if (kth > s_theta) s_theta = kth;
        atomicMax(P.theta + q, kth);
        s_ncand = P.k;
      } else {
        s_ncand = have;
      }
    }
    __syncthreads();
    for (uint32_t i = s_ncand + tid; i < P.cap; i += blockDim.x) cand[i] = 0ull;
    __syncthreads();
  };

  plan_window(g, 0);
  __syncthreads();

  uint32_t buf = 0;
  for (uint32_t w = g; w < P.n_windows; w += G, buf ^= 1u) {
    const uint32_t lo = 1u + w * P.W;
    const uint32_t span = min(P.W, P.seg.n_docs - (lo - 1u));  // docs lo .. lo+span-1
    if (tid == 0) {  // pick up thresholds published by other chains / earlier segments
      const unsigned long long gt = *reinterpret_cast<volatile unsigned long long*>(P.theta + q);
      if (gt > s_theta) s_theta = gt;
    }
    plan_window(w + G, buf ^ 1u);  // warp 0 plans the next window, then joins the work below

    // ---- accumulate: one 128-posting block per warp iteration ----
    // fp32 addition commutes, so two terms may add into a slot in any order; with three or more
    // the sum order matters, so terms are processed in phases of ascending cost (the order
    // ConjunctionScore uses, conjunction.hpp:185-195) to keep scores bit-reproducible.
    const uint32_t phases = T > 2u ? T : 1u;
    for (uint32_t ph = 0; ph < phases; ++ph) {
    const uint32_t it_begin = phases > 1u ? s_prefix[buf][ph] : 0u;
    const uint32_t it_end = phases > 1u ? s_prefix[buf][ph + 1u] : s_prefix[buf][T];
    for (uint32_t it = it_begin + warp; it < it_end; it += kTopkWarps) {
      uint32_t t = ph;
      if (phases == 1u) { t = 0; while (t + 1u < T && it >= s_prefix[buf][t + 1u]) ++t; }
      const uint32_t b = s_first[buf][t] + (it - s_prefix[buf][t]);
      const uint4 d = ld_ro_v4(P.seg.blocks + qt[t].blk_begin + b);
      uint32_t doc[4], f[4];
      decode_docs(P.seg.arena, d, lane, stage[warp], doc);
      decode_freqs(P.seg.arena, d, lane, f);
      const uint32_t len = desc_len(d.w);
      const float c0 = qt[t].c0, nc = qt[t].norm_const, nl = qt[t].norm_length;
      uint32_t nrm[4]; bool in[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t off = doc[j] - lo;
        in[j] = (4u * lane + j < len) && off < span;
        nrm[j] = in[j] ? load_norm(P.seg.norms, P.seg.norm_width, doc[j]) : 1u;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (in[j]) {
          const uint32_t off = doc[j] - lo;
          atomicAdd(&acc[swz(off)], bm25(f[j], nrm[j], c0, nc, nl));
          atomicOr(&mask[off >> 5], 1u << (off & 31u));
          if (P.conjunction) atomicAdd(reinterpret_cast<uint32_t*>(cnt) + (off >> 2), 1u << (8u * (off & 3u)));
        }
      }
    }
    if (ph + 1u < phases) __syncthreads();
    }
    __syncthreads();

    // ---- emit: matched slots -> filter -> threshold -> candidate buffer (optimistic append;
    //      slots that do not fit stay set and are retried after a compaction) ----
    for (;;) {
      const unsigned long long theta = s_theta;
      uint32_t matched = 0, pending = 0;
      for (uint32_t wi = tid; wi < words; wi += blockDim.x) {
        uint32_t m = mask[wi];
        if (!m) continue;
        uint32_t keep = 0;
        while (m) {
          const uint32_t bit = __ffs(m) - 1u;
          m &= m - 1u;
          const uint32_t off = wi * 32u + bit;
          const uint32_t si = swz(off);
          const uint32_t doc = lo + off;
          const bool ok = (!P.conjunction || cnt[off] == T) && filter_pass(P.filt, doc);
          if (ok) {
            const unsigned long long key = make_key(acc[si], P.seg.ordinal_base + doc);
            if (key > theta) {
              const uint32_t pos = atomicAdd(&s_ncand, 1u);
              if (pos >= P.cap) { keep |= 1u << bit; continue; }
              cand[pos] = key;
            }
            ++matched;
          }
          acc[si] = 0.f;
          if (P.conjunction) cnt[off] = 0;
        }
        mask[wi] = keep;
        pending |= keep;
      }
      matched = warp_sum(matched);
      if (lane == 0 && matched) atomicAdd(&s_matched, matched);
      if (!__syncthreads_or(pending != 0u)) break;
      compact();  // buffer overflowed: select, raise the threshold, retry what is left
    }
  }

  // ---- chain epilogue: sorted candidates + counts to global ----
  __syncthreads();
  compact();
  const uint32_t n_out = min(s_ncand, P.k);
  const size_t list = size_t(q) * P.lists + P.list_base + g;
  unsigned long long* out = P.cand + list * P.cap;
  for (uint32_t i = tid; i < P.cap; i += blockDim.x) out[i] = i < n_out ? cand[i] : 0ull;
  if (tid == 0) {
    P.cand_n[list] = n_out;
    if (s_matched) atomicAdd(P.total + q, static_cast<unsigned long long>(s_matched));
  }
}

// ------------------------------------------------------------------------------------------
// Merge the G sorted candidate lists of each query into its final top-k (one CTA per query).
// keys_out[q][k] sorted descending, zero-padded. Streaming: buffer = [best k so far | next chunk].
// ------------------------------------------------------------------------------------------
struct MergeParams {
  const unsigned long long* cand;  // [Q][G][stride]
  const uint32_t* cand_n;          // [Q][G] (null => every list holds `stride` entries, zeros = empty)
  uint32_t G, stride, k, cap;      // cap = power of two >= k + chunk
  unsigned long long* keys_out;    // [Q][k]
  uint32_t* n_out;                 // [Q]
};

__global__ void __launch_bounds__(kTopkThreads)
topk_merge_kernel(const MergeParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* buf = reinterpret_cast<unsigned long long*>(smem_raw);
  __shared__ uint32_t s_n;
  const uint32_t q = blockIdx.x, tid = threadIdx.x;
  for (uint32_t i = tid; i < P.cap; i += blockDim.x) buf[i] = 0ull;
  if (tid == 0) s_n = 0u;
  __syncthreads();
  const uint32_t room = P.cap - P.k;  // new entries per round
  for (uint32_t g = 0; g < P.G; ++g) {
    const unsigned long long* src = P.cand + (size_t(q) * P.G + g) * P.stride;
    const uint32_t n = P.cand_n ? min(P.cand_n[size_t(q) * P.G + g], P.stride) : P.stride;
    for (uint32_t base = 0; base < n; base += room) {
      const uint32_t take = min(room, n - base);
      const uint32_t have = s_n;
      __syncthreads();
      // Lists are sorted descending: once the head of a chunk is below the current k-th, stop.
      if (have == P.k && src[base] <= buf[P.k - 1u]) break;
      for (uint32_t i = tid; i < take; i += blockDim.x) buf[have + i] = src[base + i];
      __syncthreads();
      block_sort_desc(buf, P.cap);
      if (tid == 0) {
        uint32_t cntv = min(have + take, P.k);
        s_n = cntv;
      }
      __syncthreads();
      for (uint32_t i = s_n + tid; i < P.cap; i += blockDim.x) buf[i] = 0ull;
      __syncthreads();
    }
  }
  // zeros are "empty": count the real ones
  uint32_t real = 0;
  for (uint32_t i = tid; i < P.k; i += blockDim.x) {
    const unsigned long long v = i < s_n ? buf[i] : 0ull;
    P.keys_out[size_t(q) * P.k + i] = v;
    real += v != 0ull;
  }
  real = warp_sum(real);
  __shared__ uint32_t s_real;
  if (tid == 0) s_real = 0u;
  __syncthreads();
  if ((tid & 31u) == 0 && real) atomicAdd(&s_real, real);
  __syncthreads();
  if (tid == 0) P.n_out[q] = s_real;
}

}  // namespace sdbg
