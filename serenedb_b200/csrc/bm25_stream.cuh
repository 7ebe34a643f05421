// bm25_stream.cuh -- warp-autonomous BM25 scan + score + top-k for disjunctions of 1..4 terms (sm_100a).
//
// Reference behaviour being reproduced (paths relative to /root/reference/libs/iresearch/include/iresearch):
//   block walk    formats/posting/iterator_doc.hpp:309-430 (Collect / ScoreBlock / ProcessBatch: one 128-posting
//                 block at a time: decode -> norms -> score -> collector)
//   disjunction   search/max_score_iterator.hpp:311-356 (ScoreAndCollectWindow: Sum merge of the lists' scores)
//   sum order     search/conjunction.hpp:185-195 (sub-scores added in ascending-cost order)
//   collector     index/iterators.hpp:103-250 (buffer, select at capacity, threshold = k-th)
//
// Why a second kernel: bm25_topk_kernel (bm25_kernels.cuh) moves every window through CTA-wide phases (decode ->
// barrier -> fold by binary search over up to 4096 entries -> barrier -> emit -> barrier); ncu showed it bound by
// instruction issue with the block barrier as the largest stall. Here nothing in the scan is CTA-wide:
//
//   * a CTA is eight independent WARPS, each owning a contiguous doc sub-range of the work item's chain; the only
//     CTA-wide events are the candidate-buffer compactions (a rendezvous every ~1000 accepted candidates);
//   * every term has one LIVE block per warp (128 sorted doc ids + scores in shared memory). A step replaces the
//     live block(s) that ended at the previous frontier and then finalises the docs up to phi = min over terms of the
//     live block's last doc -- every list has been decoded at least that far. Each step retires at least one block;
//     each posting is decoded, scored and finalised exactly once;
//   * the pending entries of a live block are an index range [a0, a1) of its sorted ids, handled 32 at a time, one
//     entry per lane: an entry of term t looks for its doc in the live block of term t+1, t+2, ... (7-step binary
//     search over 128 ids); on a hit it adds its score INTO that slot (acc + s_u: the reference's ascending-cost sum
//     order, bit-reproducible for any number of terms) and dies; an entry nobody absorbs is final and is tested
//     against the threshold;
//   * block payloads arrive through the TMA engine: per warp and term two 512-byte slots, filled two blocks ahead by
//     cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes (SASS UBLKCP) and awaited on an mbarrier
//     (SYNCS), so the decode reads shared memory and does not wait for L2 / HBM;
//   * BM25 per posting is one shared-memory load: per CTA a table score[term][freq <= 8][norm byte] is built once with
//     exactly the arithmetic of bm25() (so a table hit is bit-identical to computing it); rarer (freq, norm) pairs
//     are computed.
// The term count is a run-time value and every per-term loop is rolled: the whole scan is ~2 k instructions, so the
// eight warps of a CTA, which are all at different places of it, stay inside the instruction cache (the first
// version unrolled everything per term: 17 k instructions for two terms and `no_instruction` as its top stall).
#pragma once

#include "bm25_kernels.cuh"

namespace sdbg {

constexpr uint32_t kStreamMaxTerms = 4;
constexpr uint32_t kSlotUnits = 32;    // prefetch slot = 32 x 16 B
constexpr uint32_t kLutFreqs = 8;      // table rows: freq 1..8 (a geometric(1/2) freq exceeds 8 once in 256 postings)
constexpr uint32_t kNoDoc = 0xFFFFFFFFu;
// per warp and term: docs[128] u32 | scores[128] f32 | slots[2][32] uint4 | descriptor window[32] uint4
constexpr uint32_t kStreamTermBytes = 512u + 512u + 1024u + 512u;

struct StreamCtl {   // CTA-wide control block (shared memory)
  unsigned long long theta;
  uint32_t ncand, matched, full, active;
  uint32_t hist[258];
};

// A block is prefetched into its slot unless it is a StreamVByte tail (decoded from the arena by the scalar-ish svb
// path) or larger than a slot. Its size is exact: payloads are contiguous in the arena in block order, so
// units = off16 of the next block - off16 of this one (the block table ends with a sentinel).
__device__ __forceinline__ bool block_is_svb(uint32_t packed) {
  const uint32_t de = desc_doc_enc(packed);
  return de == 5u || de == 7u || desc_freq_enc(packed) == 4u;
}

__device__ __forceinline__ void unpack4s(const uint4* p, uint32_t b, uint32_t lane, uint32_t v[4]) {
  const uint32_t bit = lane * b;
  const uint32_t w = bit >> 5, sh = bit & 31u;
  const uint4 lo = p[w];
  const uint4 hi = p[min(w + 1u, b - 1u)];
  const uint32_t mask = (1u << b) - 1u;  // b <= 31
  v[0] = __funnelshift_r(lo.x, hi.x, sh) & mask;
  v[1] = __funnelshift_r(lo.y, hi.y, sh) & mask;
  v[2] = __funnelshift_r(lo.z, hi.z, sh) & mask;
  v[3] = __funnelshift_r(lo.w, hi.w, sh) & mask;
}

// Doc ids and frequencies of one prefetched block; `pd` / `pf` point at the doc / freq payload in SHARED memory.
// `stage` = 128 u32 of per-warp shared scratch (bitset rank scatter). Lane l gets postings 4l .. 4l+3.
__device__ __forceinline__ void decode_block_smem(const uint4* pd, const uint4* pf, const uint4& d, uint32_t lane,
                                                  uint32_t* stage, uint32_t doc[4], uint32_t f[4]) {
  const uint32_t enc = desc_doc_enc(d.w), len = desc_len(d.w), prev = d.z;
  if (enc >= 8u) {                                  // de_delta_bitpack_b, b = enc - 6
    unpack4s(pd, enc - 6u, lane, doc);
    prefix_from_gaps(prev, lane, doc);
  } else if (enc == 4u) {                           // de_for_bitset (position-parallel expansion, see decode_docs)
    const uint32_t words = desc_words(d.w);
    uint4 x = make_uint4(0, 0, 0, 0);
    if (2u * lane < words) x = pd[lane];
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t base = 0;
    const uint32_t chunks = 2u * words;
    for (uint32_t i = 0; i < chunks; i += 4u) {
      const uint32_t src = i >> 2;
      const uint32_t c0 = __shfl_sync(kFull, x.x, src), c1 = __shfl_sync(kFull, x.y, src);
      const uint32_t c2 = __shfl_sync(kFull, x.z, src), c3 = __shfl_sync(kFull, x.w, src);
      const uint32_t id = prev + 32u * i + lane;
      if ((c0 >> lane) & 1u) stage[base + __popc(c0 & lt)] = id;
      base += __popc(c0);
      if ((c1 >> lane) & 1u) stage[base + __popc(c1 & lt)] = id + 32u;
      base += __popc(c1);
      if ((c2 >> lane) & 1u) stage[base + __popc(c2 & lt)] = id + 64u;
      base += __popc(c2);
      if ((c3 >> lane) & 1u) stage[base + __popc(c3 & lt)] = id + 96u;
      base += __popc(c3);
    }
    __syncwarp();
    const uint4 o = reinterpret_cast<const uint4*>(stage)[lane];
    doc[0] = o.x; doc[1] = o.y; doc[2] = o.z; doc[3] = o.w;
    __syncwarp();
  } else if (enc >= 1u && enc <= 3u) {              // de_delta_all_same_{08,16,32}
    const uint32_t raw = *reinterpret_cast<const uint32_t*>(pd);
    const uint32_t g = enc == 1u ? (raw & 0xFFu) : enc == 2u ? (raw & 0xFFFFu) : raw;
#pragma unroll
    for (int j = 0; j < 4; ++j) doc[j] = prev + g * (4u * lane + j + 1u);
  } else {                                          // de_values
    uint4 x = make_uint4(0, 0, 0, 0);
    if (4u * lane < len) x = pd[lane];
    doc[0] = x.x; doc[1] = x.y; doc[2] = x.z; doc[3] = x.w;
  }
  const uint32_t fenc = desc_freq_enc(d.w);
  if (fenc >= 5u) {
    unpack4s(pf, fenc - 4u, lane, f);
  } else if (fenc >= 1u && fenc <= 3u) {
    const uint32_t raw = *reinterpret_cast<const uint32_t*>(pf);
    f[0] = f[1] = f[2] = f[3] = fenc == 1u ? (raw & 0xFFu) : fenc == 2u ? (raw & 0xFFFFu) : raw;
  } else {
    uint4 x = make_uint4(0, 0, 0, 0);
    if (4u * lane < len) x = pf[lane];
    f[0] = x.x; f[1] = x.y; f[2] = x.z; f[3] = x.w;
  }
}

// Blocks that are not prefetched (StreamVByte tails, oversized raw blocks): decoded straight from the arena.
// Results go through shared memory (out_docs / out_freqs, 128 u32 each) so that the caller's registers stay registers.
__device__ __noinline__ void decode_block_global(const uint4* arena, uint4 d, uint32_t lane, uint32_t* out_docs,
                                                 uint32_t* out_freqs) {
  uint32_t doc[4], f[4];
  decode_docs(arena, d, lane, out_docs, doc);
  decode_freqs(arena, d, lane, f);
  __syncwarp();
  reinterpret_cast<uint4*>(out_docs)[lane] = make_uint4(doc[0], doc[1], doc[2], doc[3]);
  reinterpret_cast<uint4*>(out_freqs)[lane] = make_uint4(f[0], f[1], f[2], f[3]);
  __syncwarp();
}

// First block in B[0, n) whose last doc is >= x (n when none): 32-ary search, one descriptor per lane and round.
__device__ __forceinline__ uint32_t warp_first_block(const uint4* B, uint32_t n, uint32_t x, uint32_t lane) {
  uint32_t lo = 0;
  while (n > 32u) {
    const uint32_t step = (n + 31u) >> 5;
    const uint32_t idx = min((lane + 1u) * step, n) - 1u;      // last block of this lane's slice
    const bool ge = __ldg(&B[lo + idx].y) >= x;
    const uint32_t m = __ballot_sync(kFull, ge);
    if (m == 0u) return lo + n;
    const uint32_t fs = uint32_t(__ffs(m) - 1);
    const uint32_t nn = min((fs + 1u) * step, n) - fs * step;
    lo += fs * step; n = nn;
  }
  const bool ge = lane < n && __ldg(&B[lo + lane].y) >= x;
  const uint32_t m = __ballot_sync(kFull, ge);
  return m ? lo + uint32_t(__ffs(m) - 1) : lo + n;
}

// Candidate buffer full: exact radix select keeps the best k and raises the thresholds. Called by every thread of the
// CTA between two barriers of the rendezvous.
__device__ __noinline__ void stream_compact(StreamCtl* ctl, unsigned long long* cand, uint32_t cap, uint32_t k,
                                            unsigned long long* theta_global) {
  if (min(ctl->ncand, cap) > k) {       // uniform (shared value, read after a barrier)
    const unsigned long long kth = block_select_topk(cand, cap, k, ctl->hist);
    if (threadIdx.x == 0) {
      if (kth > ctl->theta) ctl->theta = kth;
      atomicMax(theta_global, kth);
      ctl->ncand = k;
    }
  } else if (threadIdx.x == 0) {
    ctl->ncand = min(ctl->ncand, cap);
  }
  if (threadIdx.x == 0) ctl->full = 0u;
  __syncthreads();
}

// Every warp passes through here when the buffer overflowed (ctl->full): active warps look at the flag once per step
// and inside an overflowing append; finished warps wait here until the whole CTA is done. Returns true when every
// warp of the CTA has finished its scan.
__device__ __noinline__ bool stream_rendezvous(StreamCtl* ctl, unsigned long long* cand, uint32_t cap, uint32_t k,
                                               unsigned long long* theta_global) {
  __syncthreads();
  if (ctl->full) stream_compact(ctl, cand, cap, k, theta_global);   // uniform after the barrier
  const bool done = *reinterpret_cast<volatile uint32_t*>(&ctl->active) == 0u;
  __syncthreads();
  return done;
}

// Dynamic shared memory: cand[cap] u64 | lut[T][kLutFreqs][256] f32 (kLut) | per warp: T x kStreamTermBytes.
// Terms are in ascending-cost order (the host sorts them); T-1 is the "top" term.
template <uint32_t T, bool kLut>
__global__ void __launch_bounds__(kTopkThreads, 3)
bm25_stream_kernel(const TopkParams P) {
  static_assert(T >= 1 && T <= kStreamMaxTerms, "1..4 terms");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* cand = reinterpret_cast<unsigned long long*>(smem_raw);
  float* lut = reinterpret_cast<float*>(cand + P.cap);
  unsigned char* warp_area = reinterpret_cast<unsigned char*>(lut + (kLut ? T * kLutFreqs * 256u : 0u));

  __shared__ __align__(16) StreamCtl ctl;
  __shared__ uint64_t s_bar[kTopkWarps][kStreamMaxTerms][2];
  __shared__ QTermDev s_qt[kStreamMaxTerms];

  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  unsigned char* mine = warp_area + warp * (T * kStreamTermBytes);
  auto live_docs = [&](uint32_t t) { return reinterpret_cast<uint32_t*>(mine + t * kStreamTermBytes); };
  auto live_scores = [&](uint32_t t) { return reinterpret_cast<float*>(mine + t * kStreamTermBytes + 512u); };
  auto slot_of = [&](uint32_t t, uint32_t s) { return reinterpret_cast<uint4*>(mine + t * kStreamTermBytes + 1024u + s * 512u); };
  auto desc_win = [&](uint32_t t) { return reinterpret_cast<uint4*>(mine + t * kStreamTermBytes + 2048u); };

  const uint4 work = P.work[blockIdx.x];
  const uint32_t q = work.x, g = work.y, chunk = work.z;
  const uint32_t t0 = P.qterm_off[q];
  const unsigned long long first64 = 1ull + static_cast<unsigned long long>(g) * chunk;
  const bool chain_empty = first64 > P.seg.n_docs;
  const uint32_t chain_lo = chain_empty ? 1u : uint32_t(first64);
  const uint32_t chain_hi = chain_empty ? 0u : uint32_t(min(static_cast<unsigned long long>(P.seg.n_docs), first64 + chunk - 1ull));
  const uint32_t clen = chain_empty ? 0u : chain_hi - chain_lo + 1u;
  const uint32_t sub = (clen + kTopkWarps - 1u) / kTopkWarps;
  const bool warp_empty = clen == 0u || warp * sub >= clen;
  const uint32_t lo_w = warp_empty ? 1u : chain_lo + warp * sub;
  const uint32_t hi_w = warp_empty ? 0u : min(chain_hi, lo_w + sub - 1u);

  for (uint32_t i = tid; i < P.cap; i += blockDim.x) cand[i] = 0ull;
  if (tid < T) s_qt[tid] = P.qterms[t0 + tid];
  if (tid == 0) { ctl.ncand = 0u; ctl.matched = 0u; ctl.full = 0u; ctl.active = kTopkWarps; ctl.theta = 0ull; }
  if (lane == 0) {
#pragma unroll
    for (uint32_t t = 0; t < T; ++t) { mbar_init(&s_bar[warp][t][0], 1u); mbar_init(&s_bar[warp][t][1], 1u); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if constexpr (kLut) {
    // thread = norm byte; same arithmetic as the per-posting evaluation, so a table hit is bit-identical
#pragma unroll 1
    for (uint32_t i = 0; i < T * kLutFreqs; ++i) {
      const uint32_t t = i / kLutFreqs, f = i % kLutFreqs;
      lut[i * 256u + tid] = bm25(f + 1u, tid, s_qt[t].c0, s_qt[t].norm_const, s_qt[t].norm_length);
    }
    __syncthreads();
  }
  unsigned long long* const theta_global = P.theta + q;
  const uint8_t* const norms_m1 = P.seg.norms ? P.seg.norms - 1 : nullptr;   // row = doc - 1 (1-byte norms: kLut)

  if (!warp_empty) {
    // ---- per-term stream state: registers (every loop over t is unrolled) ----
    uint32_t cur[T] = {};     // next block to load (index within the term)
    uint32_t wb[T] = {};      // first block of the descriptor window
    uint32_t start[T] = {};   // first block of this warp (slot / parity bookkeeping)
    uint32_t fr[T] = {};      // last doc of the live block (kNoDoc: list exhausted for this warp)
    uint32_t a0[T] = {};      // first pending entry of the live block (lower terms)
    uint32_t matched = 0;     // per lane; summed at the end
    unsigned long long theta = 0ull;
    uint32_t theta_hi = 0u;

    // Appends the lanes' keys (want) to the candidate buffer; on overflow: compaction, then the rest is retried.
    auto append = [&](bool want, unsigned long long key) {
      for (;;) {
        const uint32_t wbal = __ballot_sync(kFull, want);
        if (!wbal) break;
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&ctl.ncand, uint32_t(__popc(wbal)));
        base = __shfl_sync(kFull, base, 0);
        const uint32_t pos = base + __popc(wbal & ((1u << lane) - 1u));
        if (want && pos < P.cap) { cand[pos] = key; want = false; }
        if (!__any_sync(kFull, want)) break;
        if (lane == 0) *reinterpret_cast<volatile uint32_t*>(&ctl.full) = 1u;
        stream_rendezvous(&ctl, cand, P.cap, P.k, theta_global);
        want = want && key > *reinterpret_cast<volatile unsigned long long*>(&ctl.theta);
      }
    };

    // Issues the bulk copy of block b of term t into its slot (b - start) & 1; every block gets exactly one arrival
    // on its slot's barrier, in block order (a block that is not prefetched arrives with 0 bytes). The window holds
    // the descriptors of b and b + 1 (refilled before it runs out), b + 1 possibly being the table's sentinel.
    auto prefetch = [&](const uint32_t t, uint32_t b) {
      if (b >= s_qt[t].nblk) return;
      if (lane == 0) {
        const uint4* w = desc_win(t);
        const uint4 d = w[b - wb[t]];
        uint32_t units = w[b - wb[t] + 1u].x - d.x;
        if (units > kSlotUnits || block_is_svb(d.w)) units = 0u;
        const uint32_t r = b - start[t];
        uint64_t* bar = &s_bar[warp][t][r & 1u];
        mbar_arrive_expect_tx(bar, units * 16u);
        if (units) bulk_g2s(slot_of(t, r & 1u), P.seg.arena + d.x, units * 16u, bar);
      }
    };
    // Window = descriptors [wb, wb + 32) of the term (zeros past the sentinel).
    auto load_window = [&](const uint32_t t, uint32_t first) {
      __syncwarp();
      wb[t] = first;
      desc_win(t)[lane] = (first + lane <= s_qt[t].nblk) ? __ldg(P.seg.blocks + s_qt[t].blk_begin + first + lane) : make_uint4(0, 0, 0, 0);
      __syncwarp();
    };

    // Makes block cur[t] the live block of term t: wait for its payload, decode, gather norms, score, publish.
    auto advance = [&](const uint32_t t, uint32_t plo) {
      uint32_t* ld = live_docs(t);
      float* ls = live_scores(t);
      bool have = cur[t] < s_qt[t].nblk;
      uint4 d = make_uint4(0, 0, 0, 0);
      if (have) {
        if (cur[t] - wb[t] >= 28u) load_window(t, cur[t]);         // keeps cur .. cur + 3 inside the window
        d = desc_win(t)[cur[t] - wb[t]];
        have = d.z < hi_w;                                         // first doc of the block (prev_last + 1) inside the sub-range
      }
      if (!have) {
        reinterpret_cast<uint4*>(ld)[lane] = make_uint4(kNoDoc, kNoDoc, kNoDoc, kNoDoc);
        fr[t] = kNoDoc; a0[t] = 0u;
        return;
      }
      const uint32_t r = cur[t] - start[t];
      mbar_wait(&s_bar[warp][t][r & 1u], (r >> 1) & 1u);
      uint32_t doc[4], f[4];
      const uint32_t units = desc_win(t)[cur[t] - wb[t] + 1u].x - d.x;
      if (units <= kSlotUnits && !block_is_svb(d.w)) {
        const uint4* p = slot_of(t, r & 1u);
        decode_block_smem(p, p + desc_fdelta(d.w), d, lane, ld, doc, f);
      } else {
        decode_block_global(P.seg.arena, d, lane, ld, reinterpret_cast<uint32_t*>(ls));
        const uint4 x = reinterpret_cast<const uint4*>(ld)[lane], y = reinterpret_cast<const uint4*>(ls)[lane];
        doc[0] = x.x; doc[1] = x.y; doc[2] = x.z; doc[3] = x.w;
        f[0] = y.x; f[1] = y.y; f[2] = y.z; f[3] = y.w;
      }
      __syncwarp();                                                // every lane is done with the slot
      prefetch(t, cur[t] + 2u);
      const uint32_t len = desc_len(d.w);
      uint32_t nrm[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool valid = 4u * lane + j < len;
        if (!valid) { doc[j] = kNoDoc; f[j] = 1u; }
        if constexpr (kLut) nrm[j] = (valid && norms_m1) ? __ldg(norms_m1 + doc[j]) : 1u;
        else nrm[j] = valid ? load_norm(P.seg.norms, P.seg.norm_width, doc[j]) : 1u;
      }
      float s[4];
      if constexpr (kLut) {
        bool slow = false;
        const float* lt = lut + t * kLutFreqs * 256u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          slow |= f[j] > kLutFreqs;
          s[j] = lt[min(f[j] - 1u, kLutFreqs - 1u) * 256u + nrm[j]];
        }
        if (__any_sync(kFull, slow)) {
          const float c0 = s_qt[t].c0, nc = s_qt[t].norm_const, nl = s_qt[t].norm_length;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (f[j] > kLutFreqs) s[j] = bm25(f[j], nrm[j], c0, nc, nl);
        }
      } else {
        const float c0 = s_qt[t].c0, nc = s_qt[t].norm_const, nl = s_qt[t].norm_length;
#pragma unroll
        for (int j = 0; j < 4; ++j) s[j] = bm25(f[j], nrm[j], c0, nc, nl);
      }
      reinterpret_cast<uint4*>(ld)[lane] = make_uint4(doc[0], doc[1], doc[2], doc[3]);
      reinterpret_cast<float4*>(ls)[lane] = make_float4(s[0], s[1], s[2], s[3]);
      a0[t] = 0u;
      if (cur[t] == start[t]) {
        // the first block of a warp may hold docs below its sub-range: they are not pending
        a0[t] = __popc(__ballot_sync(kFull, doc[0] <= plo)) + __popc(__ballot_sync(kFull, doc[1] <= plo)) +
                __popc(__ballot_sync(kFull, doc[2] <= plo)) + __popc(__ballot_sync(kFull, doc[3] <= plo));
      }
      fr[t] = d.y;
      ++cur[t];
    };

    // Final entries (nobody absorbs them any more): count, threshold test, append.
    auto emit = [&](bool alive, uint32_t dv, float sv) {
      matched += alive ? 1u : 0u;
      bool want = alive && __float_as_uint(sv) >= theta_hi;
      if (__any_sync(kFull, want)) {
        unsigned long long key = 0ull;
        if (want) { key = make_key(sv, P.seg.ordinal_base + dv); want = key > theta; }
        append(want, key);
      }
    };

#pragma unroll
    for (uint32_t t = 0; t < T; ++t) {
      const uint32_t st = warp_first_block(P.seg.blocks + s_qt[t].blk_begin, s_qt[t].nblk, lo_w, lane);
      start[t] = st; cur[t] = st;
      load_window(t, st);
      fr[t] = lo_w - 1u;
    }
#pragma unroll
    for (uint32_t t = 0; t < T; ++t) { prefetch(t, start[t]); prefetch(t, start[t] + 1u); }

    uint32_t plo = lo_w - 1u;   // docs <= plo are final
    for (uint32_t step = 0;; ++step) {
      // ---- replace the live blocks that ended at plo (first step: every term) ----
#pragma unroll
      for (uint32_t t = 0; t < T; ++t)
        if (fr[t] == plo) advance(t, plo);
      __syncwarp();

      uint32_t phi = hi_w;
#pragma unroll
      for (uint32_t t = 0; t < T; ++t) phi = min(phi, fr[t]);
      theta = *reinterpret_cast<volatile unsigned long long*>(&ctl.theta);
      if ((step & 15u) == 0u) {   // thresholds published by other chains / earlier segments of this query
        const unsigned long long gt = *reinterpret_cast<volatile unsigned long long*>(theta_global);
        if (gt > theta) { theta = gt; if (lane == 0) atomicMax(&ctl.theta, gt); }
      }
      theta_hi = uint32_t(theta >> 32);

      // ---- lower terms: pending entries [a0, a1) with doc <= phi, 32 at a time, one entry per lane ----
#pragma unroll
      for (uint32_t t = 0; t + 1u < T; ++t) {
        const uint32_t* ld = live_docs(t);
        const float* ls = live_scores(t);
        const uint4 dd = reinterpret_cast<const uint4*>(ld)[lane];
        const uint32_t a1 = __popc(__ballot_sync(kFull, dd.x <= phi)) + __popc(__ballot_sync(kFull, dd.y <= phi)) +
                            __popc(__ballot_sync(kFull, dd.z <= phi)) + __popc(__ballot_sync(kFull, dd.w <= phi));
        for (uint32_t e0 = a0[t]; e0 < a1; e0 += 32u) {
          const uint32_t e = e0 + lane;
          bool alive = e < a1;
          const uint32_t dv = alive ? ld[e] : kNoDoc;
          const float sv = alive ? ls[e] : 0.f;
#pragma unroll
          for (uint32_t u = t + 1u; u < T; ++u) {
            // absorbed by a later term's live block? (a pending doc can only sit in live blocks: everything a list
            // holds before its live block is <= plo)
            if (fr[u] == kNoDoc) continue;                         // uniform: nothing live in term u
            const uint32_t* a = live_docs(u);
            uint32_t pos = 0;
#pragma unroll
            for (uint32_t stp = 64u; stp; stp >>= 1) pos += (a[pos + stp - 1u] < dv) ? stp : 0u;
            if (alive && a[pos] == dv) {
              float* as = live_scores(u);
              as[pos] = __fadd_rn(sv, as[pos]);                    // unique writer: docs are unique within term t
              alive = false;
            }
          }
          emit(alive, dv, sv);                                     // whatever is still alive is final
        }
        a0[t] = a1;
        __syncwarp();   // folds into later terms are visible before those terms are read
      }
      // ---- top term: its block is finalised as a whole when it retires (every lower term has been folded in up to
      // its last doc by then); lane l holds entries 4l .. 4l+3 ----
      if (fr[T - 1u] <= phi || phi >= hi_w) {
        const uint4 dd = reinterpret_cast<const uint4*>(live_docs(T - 1u))[lane];
        const float4 ss = reinterpret_cast<const float4*>(live_scores(T - 1u))[lane];
        const uint32_t dv[4] = {dd.x, dd.y, dd.z, dd.w};
        const float sv[4] = {ss.x, ss.y, ss.z, ss.w};
        const uint32_t first = 4u * lane;
        bool want_any = false;
        bool alive[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          alive[j] = first + j >= a0[T - 1u] && dv[j] <= phi;      // pads are kNoDoc > phi
          matched += alive[j] ? 1u : 0u;
          want_any |= alive[j] && __float_as_uint(sv[j]) >= theta_hi;
        }
        if (__any_sync(kFull, want_any)) {
#pragma unroll 1
          for (int j = 0; j < 4; ++j) {
            const uint32_t d1 = j == 0 ? dv[0] : j == 1 ? dv[1] : j == 2 ? dv[2] : dv[3];
            const float s1 = j == 0 ? sv[0] : j == 1 ? sv[1] : j == 2 ? sv[2] : sv[3];
            const bool al = j == 0 ? alive[0] : j == 1 ? alive[1] : j == 2 ? alive[2] : alive[3];
            bool want = al && __float_as_uint(s1) >= theta_hi;
            unsigned long long key = 0ull;
            if (want) { key = make_key(s1, P.seg.ordinal_base + d1); want = key > theta; }
            append(want, key);
          }
        }
        // on the last step of a warp (phi == hi_w) the block may still hold docs beyond the sub-range: they belong
        // to the next warp; a0 keeps what has been emitted if the same block is looked at again
        a0[T - 1u] = 128u;
      }
      plo = phi;
      if (phi >= hi_w) break;
      if (*reinterpret_cast<volatile uint32_t*>(&ctl.full)) stream_rendezvous(&ctl, cand, P.cap, P.k, theta_global);
    }
    // drain bulk copies that were issued but never consumed (they must not outlive the CTA's shared memory)
#pragma unroll
    for (uint32_t t = 0; t < T; ++t) {
      for (uint32_t b = cur[t]; b < min(s_qt[t].nblk, cur[t] + 2u); ++b) {   // issued: every block below cur + 2
        const uint32_t r = b - start[t];
        mbar_wait(&s_bar[warp][t][r & 1u], (r >> 1) & 1u);
      }
    }
    matched = warp_sum(matched);
    if (lane == 0 && matched) atomicAdd(&ctl.matched, matched);
  }
  __syncwarp();
  if (lane == 0) atomicSub(&ctl.active, 1u);
  // finished warps keep serving compactions until every warp of the CTA is done
  while (!stream_rendezvous(&ctl, cand, P.cap, P.k, theta_global)) {}

  // ---- chain epilogue: best k, sorted descending ----
  stream_compact(&ctl, cand, P.cap, P.k, theta_global);
  const uint32_t n_out = min(ctl.ncand, P.k);
  uint32_t sort_n = 256u;
  while (sort_n < n_out) sort_n <<= 1;
  block_sort_desc(cand, sort_n);
  const size_t list = work.w;
  unsigned long long* out = P.cand + list * P.cap;
  for (uint32_t i = tid; i < n_out; i += blockDim.x) out[i] = cand[i];
  if (tid == 0) {
    P.cand_n[list] = n_out;
    if (ctl.matched) atomicAdd(P.total + q, static_cast<unsigned long long>(ctl.matched));
  }
}

}  // namespace sdbg
