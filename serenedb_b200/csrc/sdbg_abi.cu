// sdbg_abi.cu -- the C ABI of libsdbg.so (include/sdbg.h): contexts, staging into HBM, query
// preparation and kernel launches. Host code only orchestrates; all per-posting / per-row work is
// in bm25_kernels.cuh and column_kernels.cuh. There is no CPU fallback anywhere in this file.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <dlfcn.h>
#include <nccl.h>   // types only: the library is resolved at run time (sdbg_dist_init)

#include <algorithm>
#include <array>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <limits>
#include <thread>
#include <vector>

#include "../../include/sdbg.h"
#include "bm25_kernels.cuh"
#include "bm25_stream.cuh"
#include "bm25_merge.cuh"
#include "column_kernels.cuh"
#include "posting_format.hpp"

using namespace sdbg;

// ------------------------------------------------------------------------------------------
// context / segment objects
// ------------------------------------------------------------------------------------------
namespace {

struct DevBuf {  // grow-only device scratch
  void* p = nullptr;
  size_t cap = 0;
};

struct ColumnObj {
  void* d_values = nullptr;
  uint64_t* d_validity = nullptr;
  int type = 0;
  uint64_t rows = 0;
  bool owned = true;
  bool has_minmax = false;
  int64_t mn = 0, mx = 0;
  long long* d_zone = nullptr;  // zonemap: {min, max} per 2048-row block in predicate key space (NOT NULL columns, built on first use)
  bool has_absmax = false;      // double columns: bits of the largest |value| (>= 0x7FF0... when NaN / inf occur)
  uint64_t absmax_bits = 0;
};

}  // namespace

struct sdbg_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;          // second lane for the top-k launch pair (driver-mode / plain chains)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string err;
  uint64_t launches = 0;
  DevBuf scratch[16];
  void* h_pinned = nullptr;  // pinned host staging for small transfers
  size_t h_pinned_cap = 0;
  void* flush = nullptr;
  size_t flush_bytes = 0;
  cudaEvent_t ev_copy[16] = {};   // one per host conversion thread (sdbg_bm25_topk_batch)
  bool scan_attr_set = false;
  bool topk_attr_set = false;
  void* nccl_comm = nullptr;   // ncclComm_t once sdbg_dist_init ran
  unsigned long long* h_oor = nullptr;   // pinned: out-of-range key count of the last deferred GROUP BY partial
  void* h_result = nullptr;     // mapped pinned memory the point-query kernels write their result into (no D2H copy)
  void* d_result = nullptr;     // its device address
  unsigned long long result_seq = 0;   // completion word value of the last point query
  size_t h_result_cap = 0;
  bool counter_zeroed = false;  // scratch[10] starts at zero; every kernel that uses it leaves it at zero
  bool oor_pending = false;
  uint64_t zone_blocks_total = 0;   // last GROUP BY scan: 2048-row blocks seen / proven dead by their zonemaps
  unsigned long long* d_zone_skipped = nullptr;
  int dist_rank = 0, dist_world = 1;
  bool merge_attr_set = false;
  // optional per-kernel timing: CUDA events recorded on `stream` around the hot kernels
  int wand = 1;       // block-max pruning level: 0 off (exact total_matches), 1 planner-level block/window skips, 2 + exact-partial-score skips of the largest term
  bool profiling = false;
  struct ProfSpan { int id; cudaEvent_t a, b; };
  std::vector<ProfSpan> spans;
  std::vector<cudaEvent_t> event_pool;
};

struct sdbg_segment {
  sdbg_ctx* ctx = nullptr;
  uint32_t n_docs = 0;
  // postings
  void* d_arena = nullptr;
  void* d_blocks = nullptr;
  void* d_anchor = nullptr;
  void* d_blkmax = nullptr;
  std::vector<uint32_t> term_blk_begin, term_docs;
  std::vector<MaxPair> term_max;
  std::vector<uint8_t> term_probe;
  std::vector<uint64_t> term_bytes;
  uint64_t arena_bytes = 0, n_blocks = 0, n_postings = 0;
  bool has_wand = false;
  float wand_b = 0.75f;   // the b the block-max (freq, norm) pairs were chosen for (BM25 default unless told otherwise)
  // norms
  void* d_norms = nullptr;
  uint32_t norm_width = 0;
  // deleted docs (DocumentMask) as a bitmap over doc ids 0..n_docs
  void* d_deleted = nullptr;
  uint64_t n_deleted = 0;
  // columns
  std::map<uint64_t, ColumnObj> cols;
};

struct sdbg_writer {
  std::unique_ptr<PostingWriter> w;
  std::vector<sdbg_term_meta> metas;
};

namespace {

int fail(sdbg_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}
#define CU(ctx, expr)                                                                                   \
  do {                                                                                                  \
    cudaError_t e_ = (expr);                                                                            \
    if (e_ != cudaSuccess)                                                                              \
      return fail((ctx), SDBG_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));              \
  } while (0)

int ensure(sdbg_ctx* c, DevBuf& b, size_t bytes) {
  if (b.cap >= bytes) return SDBG_OK;
  if (b.p) { cudaStreamSynchronize(c->stream); cudaFree(b.p); b.p = nullptr; b.cap = 0; }
  const size_t want = std::max(bytes, size_t(1) << 20);
  CU(c, cudaMalloc(&b.p, want));
  b.cap = want;
  return SDBG_OK;
}
int ensure_pinned(sdbg_ctx* c, size_t bytes) {
  if (c->h_pinned_cap >= bytes) return SDBG_OK;
  if (c->h_pinned) { cudaStreamSynchronize(c->stream); cudaFreeHost(c->h_pinned); c->h_pinned = nullptr; c->h_pinned_cap = 0; }
  const size_t want = std::max(bytes, size_t(1) << 20);
  CU(c, cudaMallocHost(&c->h_pinned, want));
  c->h_pinned_cap = want;
  return SDBG_OK;
}

__global__ void fill_u64_kernel(unsigned long long* p, size_t n, unsigned long long v) {
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) p[i] = v;
}
__global__ void flush_kernel(uint4* p, size_t n, uint32_t v) {
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
    p[i] = make_uint4(v, v, v, v);
}

// Kernel ids for sdbg_profile_*: the kernels a roofline is reported for.
enum { kProfGroupBy = 0, kProfTopk = 1, kProfMerge = 2, kProfCountSum = 3, kProfBitmap = 4, kProfIds = 5 };

struct ProfScope {  // records an event pair around a launch when profiling is on
  sdbg_ctx* c; int idx = -1;
  ProfScope(sdbg_ctx* ctx, int id) : c(ctx) {
    if (!c->profiling) return;
    auto get = [&]() { cudaEvent_t e; if (!c->event_pool.empty()) { e = c->event_pool.back(); c->event_pool.pop_back(); } else cudaEventCreate(&e); return e; };
    sdbg_ctx::ProfSpan sp{id, get(), get()};
    cudaEventRecord(sp.a, c->stream);
    c->spans.push_back(sp);
    idx = int(c->spans.size()) - 1;
  }
  ~ProfScope() { if (idx >= 0) cudaEventRecord(c->spans[size_t(idx)].b, c->stream); }
};

uint32_t next_pow2(uint32_t v) { uint32_t p = 1; while (p < v) p <<= 1; return p; }

int env_int(const char* name, int dflt) {
  const char* s = std::getenv(name);
  return s && *s ? std::atoi(s) : dflt;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// lifecycle
// ------------------------------------------------------------------------------------------
extern "C" const char* sdbg_version(void) { return "serenedb-b200 0.1 (sm_100a)"; }

extern "C" int sdbg_init(int device, sdbg_ctx** out) {
  if (!out) return SDBG_EINVAL;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return SDBG_ENODEVICE;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SDBG_ENODEVICE;
  if (prop.major != 10) return SDBG_ENODEVICE;  // kernels are built for sm_100a only
  auto* c = new sdbg_ctx;
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->wand = std::max(0, std::min(2, env_int("SDBG_WAND", 2)));
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess) {
    delete c;
    return SDBG_ECUDA;
  }
  *out = c;
  return SDBG_OK;
}

extern "C" int sdbg_dist_destroy(sdbg_ctx* c);
extern "C" void sdbg_destroy(sdbg_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (auto& b : c->scratch) if (b.p) cudaFree(b.p);
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  if (c->h_oor) cudaFreeHost(c->h_oor);
  if (c->d_zone_skipped) cudaFree(c->d_zone_skipped);
  if (c->h_result) cudaFreeHost(c->h_result);
  if (c->flush) cudaFree(c->flush);
  cudaStreamSynchronize(c->stream2);
  sdbg_dist_destroy(c);
  cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1);
  cudaEventDestroy(c->ev_fork); cudaEventDestroy(c->ev_join);
  for (cudaEvent_t e : c->ev_copy) if (e) cudaEventDestroy(e);
  cudaStreamDestroy(c->stream2);
  cudaStreamDestroy(c->stream);
  delete c;
}
extern "C" const char* sdbg_last_error(const sdbg_ctx* c) { return c ? c->err.c_str() : "null context"; }
extern "C" int sdbg_timer_start(sdbg_ctx* c) { CU(c, cudaSetDevice(c->device)); CU(c, cudaEventRecord(c->ev0, c->stream)); return SDBG_OK; }
extern "C" int sdbg_timer_stop(sdbg_ctx* c, float* ms) {
  CU(c, cudaEventRecord(c->ev1, c->stream));
  CU(c, cudaEventSynchronize(c->ev1));
  CU(c, cudaEventElapsedTime(ms, c->ev0, c->ev1));
  return SDBG_OK;
}
extern "C" int sdbg_sync(sdbg_ctx* c) {
  CU(c, cudaStreamSynchronize(c->stream));
  if (c->oor_pending) {
    c->oor_pending = false;
    if (*c->h_oor) return fail(c, SDBG_EINVAL, "GROUP BY key outside [key_min, key_min + span)");
  }
  return SDBG_OK;
}
extern "C" uint64_t sdbg_launch_count(const sdbg_ctx* c) { return c ? c->launches : 0; }
extern "C" int sdbg_flush_l2(sdbg_ctx* c) {
  CU(c, cudaSetDevice(c->device));
  if (!c->flush) { c->flush_bytes = size_t(256) << 20; CU(c, cudaMalloc(&c->flush, c->flush_bytes)); }
  static uint32_t tick = 0;
  flush_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(static_cast<uint4*>(c->flush), c->flush_bytes / 16, ++tick);
  CU(c, cudaGetLastError());
  return SDBG_OK;
}

extern "C" int sdbg_set_wand(sdbg_ctx* c, int enabled) {
  if (!c) return SDBG_EINVAL;
  c->wand = enabled < 0 ? 0 : enabled > 2 ? 2 : enabled;
  return SDBG_OK;
}
extern "C" int sdbg_profile_enable(sdbg_ctx* c, int on) {
  if (!c) return SDBG_EINVAL;
  CU(c, cudaStreamSynchronize(c->stream));
  for (auto& sp : c->spans) { c->event_pool.push_back(sp.a); c->event_pool.push_back(sp.b); }
  c->spans.clear();
  c->profiling = on != 0;
  return SDBG_OK;
}
extern "C" int sdbg_profile_read(sdbg_ctx* c, int kernel_id, double* total_ms, uint64_t* launches) {
  if (!c || !total_ms || !launches || kernel_id < 0 || kernel_id >= kProfIds) return SDBG_EINVAL;
  CU(c, cudaStreamSynchronize(c->stream));
  double ms = 0; uint64_t n = 0;
  for (auto& sp : c->spans) {
    if (sp.id != kernel_id) continue;
    float t = 0;
    CU(c, cudaEventElapsedTime(&t, sp.a, sp.b));
    ms += t; ++n;
  }
  *total_ms = ms; *launches = n;
  return SDBG_OK;
}

// ------------------------------------------------------------------------------------------
// staging
// ------------------------------------------------------------------------------------------
extern "C" int sdbg_segment_create(sdbg_ctx* c, uint32_t docs_count, sdbg_segment** out) {
  if (!c || !out) return SDBG_EINVAL;
  auto* s = new sdbg_segment;
  s->ctx = c; s->n_docs = docs_count;
  *out = s;
  return SDBG_OK;
}

namespace {
void free_postings(sdbg_segment* s) {
  if (s->d_arena) cudaFree(s->d_arena);
  if (s->d_blocks) cudaFree(s->d_blocks);
  if (s->d_anchor) cudaFree(s->d_anchor);
  s->d_anchor = nullptr;
  if (s->d_blkmax) cudaFree(s->d_blkmax);
  s->d_arena = s->d_blocks = s->d_blkmax = nullptr;
}
void free_column(ColumnObj& c) {
  if (c.d_zone) { cudaFree(c.d_zone); c.d_zone = nullptr; }
  if (c.owned && c.d_values) cudaFree(c.d_values);
  if (c.d_validity) cudaFree(c.d_validity);
  c.d_values = nullptr; c.d_validity = nullptr;
}
}  // namespace

extern "C" void sdbg_segment_destroy(sdbg_segment* s) {
  if (!s) return;
  cudaSetDevice(s->ctx->device);
  cudaStreamSynchronize(s->ctx->stream);
  free_postings(s);
  if (s->d_norms) cudaFree(s->d_norms);
  if (s->d_deleted) cudaFree(s->d_deleted);
  for (auto& kv : s->cols) free_column(kv.second);
  delete s;
}

extern "C" sdbg_ctx* sdbg_segment_context(const sdbg_segment* s) { return s ? s->ctx : nullptr; }

// Zonemap effect of the last GROUP BY scan on this context: 2048-row blocks looked at by the verdict pass and how many
// of them were proven dead (never copied, never evaluated). Synchronises the stream.
extern "C" int sdbg_scan_stats(sdbg_ctx* c, uint64_t* blocks_total, uint64_t* blocks_skipped) {
  if (!c) return SDBG_EINVAL;
  unsigned long long skipped = 0;
  if (c->d_zone_skipped) {
    CU(c, cudaMemcpyAsync(&skipped, c->d_zone_skipped, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
  }
  if (blocks_total) *blocks_total = c->zone_blocks_total;
  if (blocks_skipped) *blocks_skipped = skipped;
  return SDBG_OK;
}

extern "C" int sdbg_segment_set_wand_b(sdbg_segment* s, float wand_b) {
  if (!s) return SDBG_EINVAL;
  s->wand_b = wand_b;
  return SDBG_OK;
}

extern "C" int sdbg_stage_docs_mask(sdbg_segment* s, const uint32_t* deleted_docs, size_t n) {
  if (!s || (!deleted_docs && n)) return SDBG_EINVAL;
  sdbg_ctx* c = s->ctx;
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaStreamSynchronize(c->stream));
  if (s->d_deleted) { CU(c, cudaFree(s->d_deleted)); s->d_deleted = nullptr; }
  s->n_deleted = 0;
  if (!n) return SDBG_OK;
  std::vector<uint32_t> bits((size_t(s->n_docs) + 32) / 32 + 1, 0u);
  for (size_t i = 0; i < n; ++i) {
    const uint32_t d = deleted_docs[i];
    if (d == 0 || d > s->n_docs) return fail(c, SDBG_EINVAL, "deleted doc id outside 1..docs_count");
    if (!((bits[d >> 5] >> (d & 31u)) & 1u)) ++s->n_deleted;
    bits[d >> 5] |= 1u << (d & 31u);
  }
  CU(c, cudaMalloc(&s->d_deleted, bits.size() * 4));
  CU(c, cudaMemcpyAsync(s->d_deleted, bits.data(), bits.size() * 4, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return SDBG_OK;
}

namespace {
int upload_postings(sdbg_segment* s, const StagedPostings& sp) {
  sdbg_ctx* c = s->ctx;
  CU(c, cudaSetDevice(c->device));
  free_postings(s);
  CU(c, cudaMalloc(&s->d_arena, std::max<size_t>(sp.arena.size(), 16)));
  // one sentinel descriptor behind the last block: off16 = end of the payloads, so that "next offset - own offset" is
  // the payload size of every block (the stream kernel's prefetch size)
  CU(c, cudaMalloc(&s->d_blocks, (sp.blocks.size() + 1) * sizeof(BlockDesc)));
  {
    BlockDesc sentinel;
    sentinel.off16 = uint32_t((sp.arena.size() >= 1024 ? sp.arena.size() - 1024 : 0) / 16);
    sentinel.last_doc = 0xFFFFFFFFu; sentinel.prev_last = 0xFFFFFFFFu; sentinel.packed = 0;
    CU(c, cudaMemcpyAsync(static_cast<BlockDesc*>(s->d_blocks) + sp.blocks.size(), &sentinel, sizeof sentinel, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
  }
  CU(c, cudaMalloc(&s->d_blkmax, std::max<size_t>(sp.blk_max.size() * sizeof(MaxPair), 16)));
  CU(c, cudaMalloc(&s->d_anchor, std::max<size_t>(sp.blk_anchor.size() * 4, 16)));
  CU(c, cudaMemcpyAsync(s->d_anchor, sp.blk_anchor.data(), sp.blk_anchor.size() * 4, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaMemcpyAsync(s->d_arena, sp.arena.data(), sp.arena.size(), cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaMemcpyAsync(s->d_blocks, sp.blocks.data(), sp.blocks.size() * sizeof(BlockDesc), cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaMemcpyAsync(s->d_blkmax, sp.blk_max.data(), sp.blk_max.size() * sizeof(MaxPair), cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  s->term_blk_begin = sp.term_blk_begin;
  s->term_docs = sp.term_docs;
  s->term_max = sp.term_max;
  s->term_probe = sp.term_probe;
  s->term_bytes = sp.term_bytes;
  s->arena_bytes = sp.arena.size();
  s->n_blocks = sp.blocks.size();
  s->n_postings = sp.n_postings;
  s->has_wand = sp.has_wand;
  return SDBG_OK;
}
}  // namespace

extern "C" int sdbg_stage_postings(sdbg_segment* s, const uint8_t* doc_file, size_t n, const sdbg_term_meta* terms,
                                   size_t n_terms, int has_wand) {
  if (!s || (!doc_file && n) || (!terms && n_terms)) return SDBG_EINVAL;
  static_assert(sizeof(sdbg_term_meta) == sizeof(TermMeta), "ABI term meta mirrors the host struct");
  StagedPostings sp;
  const std::string e = stage_postings(doc_file, n, reinterpret_cast<const TermMeta*>(terms), n_terms, has_wand != 0, &sp);
  if (!e.empty()) return fail(s->ctx, SDBG_EFORMAT, e);
  for (size_t t = 0; t < n_terms; ++t)
    if (terms[t].docs_count && sp.blocks[sp.term_blk_begin[t + 1] - 1].last_doc > s->n_docs)
      return fail(s->ctx, SDBG_EFORMAT, "doc id beyond segment size");
  return upload_postings(s, sp);
}

extern "C" int sdbg_stage_norms(sdbg_segment* s, const uint8_t* bytes, size_t n, const sdbg_norm_rg* rgs, size_t n_rg) {
  if (!s || !bytes || !rgs || !n_rg) return SDBG_EINVAL;
  sdbg_ctx* c = s->ctx;
  // Row groups may differ in width (norm_column_reader.hpp:43-48); HBM holds one uniform width per
  // segment (the widest) so a gather is a single indexed load.
  uint32_t width = 1; uint64_t rows = 0;
  for (size_t i = 0; i < n_rg; ++i) {
    if (rgs[i].byte_size != 1 && rgs[i].byte_size != 2 && rgs[i].byte_size != 4) return fail(c, SDBG_EINVAL, "norm width must be 1, 2 or 4");
    if (rgs[i].file_offset + uint64_t(rgs[i].row_count) * rgs[i].byte_size > n) return fail(c, SDBG_EINVAL, "norm row group beyond buffer");
    width = std::max<uint32_t>(width, rgs[i].byte_size);
    rows += rgs[i].row_count;
  }
  if (rows != s->n_docs) return fail(c, SDBG_EINVAL, "norm rows != segment docs");
  std::vector<uint8_t> flat;
  const uint8_t* src = bytes + rgs[0].file_offset;
  if (!(n_rg == 1 && rgs[0].byte_size == width)) {
    flat.resize(rows * width);
    uint64_t r = 0;
    for (size_t i = 0; i < n_rg; ++i)
      for (uint32_t j = 0; j < rgs[i].row_count; ++j, ++r) {
        uint32_t v = 0;
        std::memcpy(&v, bytes + rgs[i].file_offset + size_t(j) * rgs[i].byte_size, rgs[i].byte_size);
        std::memcpy(flat.data() + r * width, &v, width);
      }
    src = flat.data();
  }
  CU(c, cudaSetDevice(c->device));
  if (s->d_norms) { cudaFree(s->d_norms); s->d_norms = nullptr; }
  CU(c, cudaMalloc(&s->d_norms, rows * width + 16));
  CU(c, cudaMemcpyAsync(s->d_norms, src, rows * width, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  s->norm_width = width;
  return SDBG_OK;
}

namespace {
size_t type_width(int t) { return t == SDBG_I32 ? 4 : 8; }
}

extern "C" int sdbg_stage_column(sdbg_segment* s, uint64_t field, sdbg_type t, const void* values,
                                 const uint64_t* validity, uint64_t rows) {
  if (!s || !values || t < 0 || t > 2) return SDBG_EINVAL;
  sdbg_ctx* c = s->ctx;
  CU(c, cudaSetDevice(c->device));
  ColumnObj& col = s->cols[field];
  const size_t bytes = rows * type_width(t);
  if (!(col.owned && col.d_values && col.rows == rows && col.type == t)) {
    free_column(col);
    col = ColumnObj{};
    CU(c, cudaMalloc(&col.d_values, bytes + 64));  // slack: row pairs are loaded as one 16-byte vector
    CU(c, cudaMemsetAsync(static_cast<char*>(col.d_values) + bytes, 0, 64, c->stream));
  }
  col.type = t; col.rows = rows; col.owned = true; col.has_minmax = false; col.has_absmax = false;
  CU(c, cudaMemcpyAsync(col.d_values, values, bytes, cudaMemcpyHostToDevice, c->stream));
  if (validity) {
    const size_t vb = ((rows + 63) / 64) * 8;
    if (!col.d_validity) CU(c, cudaMalloc(reinterpret_cast<void**>(&col.d_validity), vb + 16));
    CU(c, cudaMemcpyAsync(col.d_validity, validity, vb, cudaMemcpyHostToDevice, c->stream));
  } else if (col.d_validity) {
    cudaFree(col.d_validity); col.d_validity = nullptr;
  }
  return SDBG_OK;  // asynchronous on the context stream; consumers run on the same stream
}

// ---- frame-of-reference bit-packed int64 columns (DuckDB's bitpacking codec in FOR mode is this algorithm: per group
// of 2048 values a base and a bit width, formats/column/column_reader.hpp:90-96 ColumnBlockMeta::codec; DuckDB itself is
// not vendored in the reference tree, so the byte layout below is this library's own) ----
static_assert(sizeof(sdbg_for_block) == sizeof(ForBlockDev), "header layout");

extern "C" int sdbg_pack_for(const int64_t* values, uint64_t rows, sdbg_for_block* headers, uint64_t* words, uint64_t cap_words,
                             uint64_t* n_words) {
  if (!values || !headers || !n_words || (cap_words && !words)) return SDBG_EINVAL;
  const uint64_t n_groups = (rows + kForGroupRows - 1) / kForGroupRows;
  // pass 1 (threaded): base = min, bits = width of max - min, word count per group
  const size_t n_thr = std::max<size_t>(1, std::min<size_t>(size_t(env_int("SDBG_HOST_THREADS", 16)), n_groups / 64 + 1));
  auto stats = [&](uint64_t g0, uint64_t g1) {
    for (uint64_t g = g0; g < g1; ++g) {
      const uint64_t r0 = g * kForGroupRows, r1 = std::min<uint64_t>(rows, r0 + kForGroupRows);
      int64_t mn = values[r0], mx = values[r0];
      for (uint64_t r = r0 + 1; r < r1; ++r) { mn = std::min(mn, values[r]); mx = std::max(mx, values[r]); }
      const uint64_t span = uint64_t(mx) - uint64_t(mn);
      headers[g].base = mn;
      headers[g].bits = span == 0 ? 0u : uint32_t(64 - __builtin_clzll(span));
    }
  };
  {
    std::vector<std::thread> pool;
    for (size_t t = 0; t < n_thr; ++t) pool.emplace_back(stats, n_groups * t / n_thr, n_groups * (t + 1) / n_thr);
    for (auto& th : pool) th.join();
  }
  uint64_t off = 0;
  for (uint64_t g = 0; g < n_groups; ++g) {
    const uint64_t n = std::min<uint64_t>(kForGroupRows, rows - g * kForGroupRows);
    if (off > 0xFFFFFFFFull) return SDBG_EUNSUPPORTED;            // 32-bit word offsets: 32 GiB of packed data per column
    headers[g].off8 = uint32_t(off);
    off += (n * headers[g].bits + 63) / 64;
  }
  *n_words = off + 1;                                             // one word of slack: the decoder may read one past a value
  if (*n_words > cap_words) return SDBG_ECAPACITY;
  auto pack = [&](uint64_t g0, uint64_t g1) {
    for (uint64_t g = g0; g < g1; ++g) {
      const uint32_t bits = headers[g].bits;
      if (!bits) continue;
      const uint64_t r0 = g * kForGroupRows, n = std::min<uint64_t>(kForGroupRows, rows - r0);
      uint64_t* w = words + headers[g].off8;
      const uint64_t nw = (n * bits + 63) / 64;
      std::fill(w, w + nw, 0ull);
      const uint64_t base = uint64_t(headers[g].base);
      for (uint64_t i = 0; i < n; ++i) {
        const uint64_t v = uint64_t(values[r0 + i]) - base, bit = i * bits;
        w[bit >> 6] |= v << (bit & 63);
        if ((bit & 63) + bits > 64) w[(bit >> 6) + 1] |= v >> (64 - (bit & 63));
      }
    }
  };
  {
    std::vector<std::thread> pool;
    for (size_t t = 0; t < n_thr; ++t) pool.emplace_back(pack, n_groups * t / n_thr, n_groups * (t + 1) / n_thr);
    for (auto& th : pool) th.join();
  }
  words[off] = 0;
  return SDBG_OK;
}

extern "C" int sdbg_stage_column_for(sdbg_segment* s, uint64_t field, const sdbg_for_block* headers, const uint64_t* words,
                                     uint64_t n_words, uint64_t rows) {
  if (!s || !headers || !words || !rows || !n_words) return SDBG_EINVAL;
  sdbg_ctx* c = s->ctx;
  CU(c, cudaSetDevice(c->device));
  const uint64_t n_groups = (rows + kForGroupRows - 1) / kForGroupRows;
  for (uint64_t g = 0; g < n_groups; ++g) {                       // the stream is untrusted input: every group must lie inside it
    const uint64_t n = std::min<uint64_t>(kForGroupRows, rows - g * kForGroupRows);
    if (headers[g].bits > 64u || uint64_t(headers[g].off8) + (n * headers[g].bits + 63) / 64 + 1 > n_words)
      return fail(c, SDBG_EFORMAT, "bit-packed column: group outside the word stream");
  }
  ColumnObj& col = s->cols[field];
  const size_t bytes = rows * 8;
  if (!(col.owned && col.d_values && col.rows == rows && col.type == SDBG_I64)) {
    free_column(col);
    col = ColumnObj{};
    CU(c, cudaMalloc(&col.d_values, bytes + 64));
    CU(c, cudaMemsetAsync(static_cast<char*>(col.d_values) + bytes, 0, 64, c->stream));
  }
  if (col.d_validity) { cudaFree(col.d_validity); col.d_validity = nullptr; }
  col.type = SDBG_I64; col.rows = rows; col.owned = true; col.has_minmax = false; col.has_absmax = false;
  if (col.d_zone) { cudaFree(col.d_zone); col.d_zone = nullptr; }
  DevBuf& buf = c->scratch[12];
  const size_t hdr_bytes = (n_groups * sizeof(ForBlockDev) + 255) & ~size_t(255);
  int rc = ensure(c, buf, hdr_bytes + n_words * 8);
  if (rc) return rc;
  char* base = static_cast<char*>(buf.p);
  CU(c, cudaMemcpyAsync(base, headers, n_groups * sizeof(ForBlockDev), cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaMemcpyAsync(base + hdr_bytes, words, n_words * 8, cudaMemcpyHostToDevice, c->stream));
  const unsigned grid = unsigned(std::min<uint64_t>((n_groups * 32 + 255) / 256, uint64_t(c->sm_count) * 16));
  for_unpack_kernel<<<std::max(grid, 1u), 256, 0, c->stream>>>(reinterpret_cast<const ForBlockDev*>(base),
                                                               reinterpret_cast<const unsigned long long*>(base + hdr_bytes), rows,
                                                               static_cast<long long*>(col.d_values));
  ++c->launches;
  CU(c, cudaGetLastError());
  return SDBG_OK;   // asynchronous on the context stream, like sdbg_stage_column
}

extern "C" int sdbg_stage_column_device(sdbg_segment* s, uint64_t field, sdbg_type t, const void* d_values, uint64_t rows) {
  if (!s || !d_values || t < 0 || t > 2) return SDBG_EINVAL;
  if (rows & 1) return fail(s->ctx, SDBG_EINVAL, "borrowed device columns need an even row count (16-byte row pairs)");
  ColumnObj& col = s->cols[field];
  free_column(col);
  col = ColumnObj{};
  col.d_values = const_cast<void*>(d_values); col.type = t; col.rows = rows; col.owned = false;
  return SDBG_OK;
}

extern "C" int sdbg_column_device_ptr(sdbg_segment* s, uint64_t field, void** d_values, uint64_t* rows) {
  if (!s) return SDBG_EINVAL;
  auto it = s->cols.find(field);
  if (it == s->cols.end()) return fail(s->ctx, SDBG_ENOTFOUND, "unknown column");
  if (d_values) *d_values = it->second.d_values;
  if (rows) *rows = it->second.rows;
  return SDBG_OK;
}

extern "C" int sdbg_column_to_host(sdbg_segment* s, uint64_t field, void* host_dst, uint64_t rows) {
  if (!s || !host_dst) return SDBG_EINVAL;
  auto it = s->cols.find(field);
  if (it == s->cols.end()) return fail(s->ctx, SDBG_ENOTFOUND, "unknown column");
  if (rows > it->second.rows) return fail(s->ctx, SDBG_EINVAL, "more rows requested than staged");
  sdbg_ctx* c = s->ctx;
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaMemcpyAsync(host_dst, it->second.d_values, rows * type_width(it->second.type), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return SDBG_OK;
}

// Values of one column for a set of hit docs (late materialisation: HitBatcher::MaterializeColumn, hit_batcher.hpp:39;
// FinalizeBatch in duckdb_search_full_scan.cpp fetches the projected columns for the emitted doc ids only).
extern "C" int sdbg_gather_column(sdbg_segment* s, uint64_t field, const uint32_t* docs, size_t n, void* out_values, uint8_t* out_valid) {
  if (!s || (n && (!docs || !out_values))) return SDBG_EINVAL;
  sdbg_ctx* c = s->ctx;
  auto it = s->cols.find(field);
  if (it == s->cols.end()) return fail(c, SDBG_ENOTFOUND, "unknown column");
  if (!n) return SDBG_OK;
  CU(c, cudaSetDevice(c->device));
  const ColumnObj& co = it->second;
  const size_t w = type_width(co.type);
  const size_t docs_bytes = (n * 4 + 255) & ~size_t(255), val_bytes = (n * w + 255) & ~size_t(255);
  DevBuf& buf = c->scratch[12];
  int rc = ensure(c, buf, docs_bytes + val_bytes + n);
  if (rc) return rc;
  char* base = static_cast<char*>(buf.p);
  auto* d_docs = reinterpret_cast<uint32_t*>(base);
  void* d_out = base + docs_bytes;
  auto* d_valid = reinterpret_cast<unsigned char*>(base + docs_bytes + val_bytes);
  CU(c, cudaMemcpyAsync(d_docs, docs, n * 4, cudaMemcpyHostToDevice, c->stream));
  const unsigned grid = unsigned(std::min<size_t>((n + 255) / 256, size_t(c->sm_count) * 8));
  const auto* valid = reinterpret_cast<const unsigned long long*>(co.d_validity);
  if (w == 4) gather_rows_kernel<uint32_t><<<grid, 256, 0, c->stream>>>(static_cast<const uint32_t*>(co.d_values), valid, d_docs, n, co.rows, static_cast<uint32_t*>(d_out), out_valid ? d_valid : nullptr);
  else gather_rows_kernel<unsigned long long><<<grid, 256, 0, c->stream>>>(static_cast<const unsigned long long*>(co.d_values), valid, d_docs, n, co.rows, static_cast<unsigned long long*>(d_out), out_valid ? d_valid : nullptr);
  ++c->launches;
  CU(c, cudaGetLastError());
  CU(c, cudaMemcpyAsync(out_values, d_out, n * w, cudaMemcpyDeviceToHost, c->stream));
  if (out_valid) CU(c, cudaMemcpyAsync(out_valid, d_valid, n, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return SDBG_OK;
}

extern "C" int sdbg_segment_posting_stats(const sdbg_segment* s, uint64_t* payload_bytes, uint64_t* table_bytes,
                                          uint64_t* n_blocks, uint64_t* n_postings) {
  if (!s) return SDBG_EINVAL;
  if (payload_bytes) *payload_bytes = s->arena_bytes;
  if (table_bytes) *table_bytes = s->n_blocks * (sizeof(BlockDesc) + sizeof(MaxPair));
  if (n_blocks) *n_blocks = s->n_blocks;
  if (n_postings) *n_postings = s->n_postings;
  return SDBG_OK;
}

extern "C" int sdbg_segment_term_bytes(const sdbg_segment* s, uint64_t* bytes_out, size_t n_terms) {
  if (!s || !bytes_out || n_terms > s->term_bytes.size()) return SDBG_EINVAL;
  std::copy(s->term_bytes.begin(), s->term_bytes.begin() + long(n_terms), bytes_out);
  return SDBG_OK;
}

// ------------------------------------------------------------------------------------------
// BM25 top-k
// ------------------------------------------------------------------------------------------
namespace {

PostingsDev postings_view(const sdbg_segment* s, uint32_t ordinal_base) {
  PostingsDev p;
  p.arena = static_cast<const uint4*>(s->d_arena);
  p.blocks = static_cast<const uint4*>(s->d_blocks);
  p.blk_max = static_cast<const uint2*>(s->d_blkmax);
  p.anchors = static_cast<const uint4*>(s->d_anchor);
  p.norms = static_cast<const uint8_t*>(s->d_norms);
  p.norm_width = s->norm_width;
  p.deleted = static_cast<const uint32_t*>(s->d_deleted);
  p.n_docs = s->n_docs;
  p.ordinal_base = ordinal_base;
  return p;
}

int filter_view(sdbg_segment* s, const sdbg_col_pred* f, FilterDev* out) {
  std::memset(out, 0, sizeof *out);
  if (!f) return SDBG_OK;
  auto it = s->cols.find(f->field);
  if (it == s->cols.end()) return fail(s->ctx, SDBG_ENOTFOUND, "filter column not staged");
  if (it->second.rows < s->n_docs) return fail(s->ctx, SDBG_EINVAL, "filter column shorter than segment");
  out->values = it->second.d_values; out->validity = it->second.d_validity; out->type = it->second.type;
  out->op = f->op; out->lo_i = f->lo_i; out->hi_i = f->hi_i; out->lo_f = f->lo_f; out->hi_f = f->hi_f;
  return SDBG_OK;
}

constexpr int kMaxCopyEvents = 16;
constexpr float kTfidfK1 = -1.f;   // internal selector of the TFIDF scorer (it has no k / b): see sdbg_tfidf_topk_batch

// Device-side descriptor of one query term over one segment (the caller has checked the term id).
void fill_qterm(const sdbg_segment* s, const sdbg_bm25_term& t, float k1, float b, QTermDev& d) {
  d.blk_begin = s->term_blk_begin[t.term];
  d.nblk = s->term_blk_begin[t.term + 1] - d.blk_begin;
  d.c0 = t.boost * (k1 + 1) * t.idf;  // bm25.cpp:224
  d.norm_const = t.norm_const; d.norm_length = t.norm_length;
  if (k1 == kTfidfK1) {                                                  // TFIDF (tfidf.cpp:59-80, 101): c0 = boost * idf
    d.c0 = t.boost * t.idf;
    d.norm_const = std::numeric_limits<float>::quiet_NaN();               // device-side marker, see bm25()
    d.norm_length = b != 0.f ? 1.f : 0.f;                                 // normalised by sqrt(doc length) or not
  }
  else if (k1 == 0.f) d.c0 = 0.f;                                       // BM1: Bm1Score without a filter boost scores 0 (bm25.cpp:118-126)
  else if (b == 0.f) d.norm_length = std::numeric_limits<float>::quiet_NaN();   // BM15 form (device-side marker, see bm25())
  d.docs_count = s->term_docs[t.term];
  d.root_freq = s->term_max[t.term].freq & 0x7FFFFFFFu; d.root_norm = s->term_max[t.term].norm;
  if (t.term < s->term_probe.size() && s->term_probe[t.term]) d.root_freq |= 0x80000000u;   // probe-friendly list (driver mode)
}

struct TopkPlan {
  uint32_t G, cap, k, budget;
  size_t smem;
};

// Device-side result of a batch: keys_out[Q][k] (sorted desc, 0 = empty), n_out[Q], total[Q].
struct TopkDevOut { unsigned long long* keys; uint32_t* n_out; unsigned long long* total; };

int topk_run(sdbg_segment* const* segs, size_t n_segs, int kind, const sdbg_bm25_term* terms, const uint32_t* term_off,
             size_t nq, float k1, const float b, const sdbg_col_pred* filt, uint32_t k, float threshold_in, TopkDevOut* dev) {
  if (!segs || !n_segs || !terms || !term_off || !nq || !k) return SDBG_EINVAL;
  sdbg_ctx* c = segs[0]->ctx;
  if (k > 8192) return fail(c, SDBG_EUNSUPPORTED, "k > 8192");
  if (nq > 65535) return fail(c, SDBG_EUNSUPPORTED, "more than 65535 queries per batch");
  CU(c, cudaSetDevice(c->device));
  const uint32_t total_terms = term_off[nq];
  for (size_t q = 0; q < nq; ++q) {
    const uint32_t nt = term_off[q + 1] - term_off[q];
    if (nt == 0 || nt > kMaxQueryTerms) return fail(c, SDBG_EUNSUPPORTED, "a query needs 1..16 terms");
  }
  uint64_t ord = 0;
  uint32_t max_docs = 0;
  for (size_t si = 0; si < n_segs; ++si) {
    if (segs[si]->ctx != c) return fail(c, SDBG_EINVAL, "segments of one call must share a context");
    if (!segs[si]->d_blocks) return fail(c, SDBG_EINVAL, "segment has no staged postings");
    ord += segs[si]->n_docs;
    max_docs = std::max(max_docs, segs[si]->n_docs);
  }
  if (ord >= 0xFFFFFFFFull) return fail(c, SDBG_EUNSUPPORTED, "more than 2^32-1 docs per GPU");

  TopkPlan pl;
  pl.k = k;
  pl.budget = uint32_t(env_int("SDBG_TOPK_BUDGET", 32));
  if (pl.budget != 16 && pl.budget != 32) return fail(c, SDBG_EINVAL, "SDBG_TOPK_BUDGET must be 16 or 32");
  const uint32_t entries = pl.budget * 128u;
  pl.cap = std::max(next_pow2(k + 1024), uint32_t(env_int("SDBG_TOPK_CAP", 2048)));  // selection is O(n): buffer size trades shared memory (occupancy) against selection count
  // Enough CTAs to fill the machine a few times over; a query is split into chains (contiguous doc
  // ranges) only when the batch alone cannot do that.
  const uint32_t target_ctas = uint32_t(c->sm_count) * 8u;
  pl.G = uint32_t(std::max<size_t>(1, (target_ctas + nq - 1) / nq));
  const uint32_t max_chains = uint32_t(env_int("SDBG_TOPK_MAX_CHAINS", 296));
  // Work list: (segment, query) pairs get max(G, postings / target) chains, so that a query over a 5 M-doc
  // list is not one CTA-long critical path next to thousands of short ones; largest chains are issued first.
  uint64_t batch_postings = 0;
  for (size_t si = 0; si < n_segs; ++si)
    for (uint32_t i = 0; i < total_terms; ++i)
      if (terms[i].term < segs[si]->term_docs.size()) batch_postings += segs[si]->term_docs[terms[i].term];
  const uint64_t chain_target = std::max<uint64_t>(uint64_t(env_int("SDBG_TOPK_CHAIN_MIN", 65536)), batch_postings / (uint64_t(c->sm_count) * uint64_t(std::max(1, env_int("SDBG_TOPK_CHAIN_DIV", 4)))));
  // Work classes (one launch each):
  //   0 / 1        legacy window kernel (driver mode / plain): > 4-term disjunctions, BM15 / BM1 forms
  //   2 + (T-1)    exhaustive warp-autonomous merge of T = 1..4 lists (bm25_merge.cuh): pruning off or not applicable
  //   6 + (T-1)    stream kernel, disjunction with MaxScore demotion / single-list block-max skip (bm25_stream.cuh)
  //   10           conjunctions: lead list + probes (stream kernel, kModeAnd), hybrid filter and deleted docs included
  //   11           first slice of a two-term disjunction whose long list is worth probing instead of scanning: merged
  //                exhaustively FIRST so that the query has a threshold when the rest of its range starts
  //   12           the rest of such a query: launched into BOTH the merge kernel and the stream kernel in lead mode
  //                (short list streamed, long list probed per candidate); the first CTA to arrive decides from the
  //                threshold which of the two runs the item (TopkParams::claim)
  struct WorkItem { uint32_t q, lo, len, list; uint64_t weight; uint32_t cls; };
  constexpr uint32_t kClsMerge = 2, kClsStream = 2 + kStreamMaxTerms, kClsAnd = 2 + 2 * kStreamMaxTerms, kClsSlice = kClsAnd + 1,
                     kClsLead = kClsAnd + 2, kClasses = kClsAnd + 3;
  const bool level2 = c->wand >= 2 && kind != SDBG_QUERY_AND && k1 != 0.f && k1 != kTfidfK1 && b != 0.f;
  const bool stream_ok = env_int("SDBG_STREAM", 1) != 0 && k1 != 0.f && b != 0.f && k1 != kTfidfK1 &&
                         size_t(pl.cap) * 8 + size_t(kStreamMaxTerms) * (kLutFreqs * 1024 + kTopkWarps * kStreamTermBytes) <= 200 * 1024;
  const bool lead_ok = env_int("SDBG_STREAM_LEAD", 1) != 0;
  // The staged block-max pairs are maximisers for BM25 with the index-time b only (FreqNormProducer::CmpBm25,
  // wand_writer.hpp:142-175; the order of two pairs does not depend on k); the reference enables WAND only when
  // Scorer::equals matches (PostingsReaderImpl::WandIterator, reader.hpp:457-501). Any other b: exhaustive.
  auto seg_wand = [&](const sdbg_segment* s) { return (c->wand && s->has_wand && k1 != 0.f && k1 != kTfidfK1 && b != 0.f && b == s->wand_b) ? c->wand : 0; };
  std::vector<std::array<size_t, kClasses>> n_cls(n_segs);
  for (auto& a : n_cls) a.fill(0);
  std::vector<std::vector<WorkItem>> seg_work(n_segs);
  std::vector<uint32_t> list_off(nq + 1, 0);
  for (size_t q = 0; q < nq; ++q) {          // lists of one query are contiguous: [segment 0 chains | segment 1 chains | ...]
    uint32_t lists = 0;
    for (size_t si = 0; si < n_segs; ++si) {
      const sdbg_segment* s = segs[si];
      uint64_t postings = 0;
      uint32_t largest = 0, largest_term = UINT32_MAX, smallest = UINT32_MAX;
      for (uint32_t i = term_off[q]; i < term_off[q + 1]; ++i)
        if (terms[i].term < s->term_docs.size()) {
          const uint32_t dc = s->term_docs[terms[i].term];
          postings += dc;
          smallest = std::min(smallest, dc);
          if (dc >= largest) { largest = dc; largest_term = terms[i].term; }
        }
      const uint32_t nt = term_off[q + 1] - term_off[q];
      const int wand = seg_wand(s);
      // Driver mode (pruning level 2) pays only when the largest list can be probed without decoding blocks; the
      // other queries run the plain kernel, which is lighter (fewer registers, no probe buffers, level-1 planner).
      const bool drive_q = level2 && wand && nt >= 2 && largest_term < s->term_probe.size() && s->term_probe[largest_term] != 0;
      uint32_t cls = drive_q ? 0u : 1u;
      uint32_t slice_docs = 0;    // > 0: lead candidate
      if (stream_ok && kind == SDBG_QUERY_AND && env_int("SDBG_STREAM_AND", 1) != 0) cls = kClsAnd;
      else if (stream_ok && kind != SDBG_QUERY_AND && nt <= kStreamMaxTerms) {
        const bool plain = !filt && !s->d_deleted;                  // the merge kernel has no per-doc checks
        cls = (!plain || (wand && (nt != 2 || !lead_ok))) ? kClsStream + (nt - 1u) : kClsMerge + (nt - 1u);
        if (wand && nt == 2 && lead_ok && plain && uint64_t(largest) >= 4ull * smallest && smallest >= 3u * k) {
          // Lead mode needs the threshold above the long list's bound. That happens when a typical posting of the short
          // list (freq 1, average length) already outscores the best posting of the long one; then about k docs of the
          // short list, i.e. the first 1.5 k / |short| of the doc range, are enough to get there.
          const sdbg_bm25_term* ta = &terms[term_off[q]];
          const sdbg_bm25_term* tb = ta + 1;
          if (s->term_docs[ta->term] > s->term_docs[tb->term]) std::swap(ta, tb);       // ta = short list
          const MaxPair root = tb->term < s->term_max.size() ? s->term_max[tb->term] : MaxPair{0, 0};
          if (root.freq != 0) {
            const float c0b = tb->boost * (k1 + 1) * tb->idf, c1b = tb->norm_const + tb->norm_length * float(root.norm);
            const float ub_b = c0b - c0b * c1b / (c1b + float(root.freq));
            const float c0a = ta->boost * (k1 + 1) * ta->idf, c1a = ta->norm_const + ta->norm_length * (ta->norm_length > 0.f ? (k1 * b) / ta->norm_length : 1.f);
            const float typ_a = c0a - c0a * c1a / (c1a + 1.f);                           // freq 1 at the average length
            if (ub_b * 1.05f < typ_a) {
              const double frac = 1.5 * double(k) / double(smallest);
              slice_docs = uint32_t(std::min<double>(double(s->n_docs), std::max(4096.0, std::ceil(double(s->n_docs) * frac))));
              if (slice_docs > s->n_docs / 2) slice_docs = 0;
            }
          }
        }
      }
      const uint32_t first = slice_docs ? slice_docs + 1u : 1u;             // first doc of the chained range
      const uint32_t rest_docs = s->n_docs - (first - 1u);
      const uint64_t rest_postings = slice_docs ? uint64_t(double(postings) * double(rest_docs) / double(s->n_docs)) : postings;
      if (slice_docs) {
        seg_work[si].push_back({uint32_t(q), 1u, slice_docs, list_off[q] + lists, postings - rest_postings, kClsSlice});
        ++n_cls[si][kClsSlice];
        ++lists;
        cls = kClsLead;
      }
      uint32_t g = uint32_t(std::max<uint64_t>(pl.G, (rest_postings + chain_target - 1) / chain_target));
      // lead mode is latency-bound (dependent loads per probe), not throughput-bound: more, shorter chains
      if (slice_docs) g = std::max(g, std::min<uint32_t>(uint32_t(env_int("SDBG_STREAM_LEAD_CHAINS", 16)), std::max(1u, smallest / 8192u)));
      g = std::min(g, max_chains);
      g = std::min(g, std::max(1u, rest_docs / 4096u));
      const uint32_t chunk = (rest_docs + g - 1) / g;
      for (uint32_t j = 0; j < g; ++j)
        seg_work[si].push_back({uint32_t(q), first + j * chunk, chunk, list_off[q] + lists + j, rest_postings / g, cls});
      n_cls[si][cls] += g;
      lists += g;
    }
    list_off[q + 1] = list_off[q] + lists;
  }
  const uint32_t total_lists = list_off[nq];
  size_t total_work = 0;
  for (auto& w : seg_work) {
    std::stable_sort(w.begin(), w.end(), [](const WorkItem& x, const WorkItem& y) { return x.cls != y.cls ? x.cls < y.cls : x.weight > y.weight; });
    total_work += w.size();
  }
  pl.smem = size_t(entries) * 8 + (kind == SDBG_QUERY_AND ? entries : 0) + size_t(pl.cap) * 8;
  const size_t smem_drive = pl.smem + size_t(entries) * 4;   // probe list | decode-fallback list
  if (smem_drive > 200 * 1024) return fail(c, SDBG_EUNSUPPORTED, "hash window + candidate buffer exceed shared memory");

  // host-side query descriptors, per segment, sorted by ascending docs_count (conjunction.hpp:520-523)
  const size_t qt_bytes = size_t(total_terms) * sizeof(QTermDev) * n_segs;
  const size_t off_bytes = (nq + 1) * sizeof(uint32_t);
  const size_t work_bytes = total_work * sizeof(uint4);
  const size_t qt_pad = (qt_bytes + off_bytes + 15) & ~size_t(15);   // work items are 16-byte loads
  int rc = ensure_pinned(c, qt_pad + work_bytes + off_bytes);
  if (rc) return rc;
  auto* h_qt = static_cast<QTermDev*>(c->h_pinned);
  auto* h_off = reinterpret_cast<uint32_t*>(static_cast<char*>(c->h_pinned) + qt_bytes);
  std::memcpy(h_off, term_off, off_bytes);
  {
    auto* h_work = reinterpret_cast<uint4*>(static_cast<char*>(c->h_pinned) + qt_pad);
    for (auto& w : seg_work) for (const WorkItem& it : w) *h_work++ = make_uint4(it.q, it.lo, it.len, it.list);
    std::memcpy(static_cast<char*>(c->h_pinned) + qt_pad + work_bytes, list_off.data(), off_bytes);
  }
  for (size_t si = 0; si < n_segs; ++si) {
    const sdbg_segment* s = segs[si];
    QTermDev* dst = h_qt + si * total_terms;
    for (size_t q = 0; q < nq; ++q) {
      const uint32_t t_begin = term_off[q], t_end = term_off[q + 1];
      for (uint32_t i = t_begin; i < t_end; ++i) {
        const sdbg_bm25_term& t = terms[i];
        if (t.term + 1 >= s->term_blk_begin.size()) return fail(c, SDBG_EINVAL, "term id out of range");
        fill_qterm(s, t, k1, b, dst[i]);
      }
      std::stable_sort(dst + t_begin, dst + t_end, [](const QTermDev& x, const QTermDev& y) { return x.docs_count < y.docs_count; });
    }
  }
  DevBuf& b_qt = c->scratch[0]; DevBuf& b_theta = c->scratch[1]; DevBuf& b_cand = c->scratch[2];
  DevBuf& b_candn = c->scratch[3]; DevBuf& b_keys = c->scratch[4]; DevBuf& b_small = c->scratch[5];
  if ((rc = ensure(c, b_qt, qt_pad + work_bytes + off_bytes))) return rc;
  if ((rc = ensure(c, b_theta, nq * 16))) return rc;  // theta[nq] | total[nq]
  if ((rc = ensure(c, b_cand, size_t(total_lists) * pl.cap * 8))) return rc;
  if ((rc = ensure(c, b_candn, size_t(total_lists) * 4))) return rc;
  if ((rc = ensure(c, b_keys, nq * size_t(k) * 8))) return rc;
  if ((rc = ensure(c, b_small, nq * 4))) return rc;
  CU(c, cudaMemcpyAsync(b_qt.p, c->h_pinned, qt_pad + work_bytes + off_bytes, cudaMemcpyHostToDevice, c->stream));
  auto* d_theta = static_cast<unsigned long long*>(b_theta.p);
  auto* d_total = d_theta + nq;
  uint32_t thr_bits; std::memcpy(&thr_bits, &threshold_in, 4);
  if (!(threshold_in >= 0.f)) thr_bits = 0;  // negative / NaN seeds accept every positive score
  fill_u64_kernel<<<64, 256, 0, c->stream>>>(d_theta, nq, (static_cast<unsigned long long>(thr_bits) << 32) | 0xFFFFFFFFull);
  ++c->launches;
  CU(c, cudaMemsetAsync(d_total, 0, nq * 8, c->stream));

  // shared memory of the stream kernel: candidates | score table | per warp (live blocks + prefetch slots)
  bool use_lut[n_segs ? n_segs : 1];
  for (size_t si = 0; si < n_segs; ++si) use_lut[si] = segs[si]->norm_width == 1 && env_int("SDBG_STREAM_LUT", 0) != 0;
  auto stream_smem = [&](uint32_t T, bool lut, bool conj) {
    return size_t(pl.cap) * 8 + (lut ? size_t(T) * kLutFreqs * 1024 : 0) + size_t(kTopkWarps) * (T * kStreamTermBytes + (conj ? 1024 : 0));
  };
  if (!c->topk_attr_set) {
    CU(c, cudaFuncSetAttribute(bm25_topk_kernel<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CU(c, cudaFuncSetAttribute(bm25_topk_kernel<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CU(c, cudaFuncSetAttribute(bm25_topk_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CU(c, cudaFuncSetAttribute(bm25_topk_kernel<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
#define SDBG_STREAM_ATTR(TT) \
    CU(c, cudaFuncSetAttribute(bm25_merge_kernel<TT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
    CU(c, cudaFuncSetAttribute(bm25_merge_kernel<TT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
    CU(c, cudaFuncSetAttribute(bm25_stream_kernel<TT, false, 3, kModeOr>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
    CU(c, cudaFuncSetAttribute(bm25_stream_kernel<TT, true, 3, kModeOr>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024))
    SDBG_STREAM_ATTR(1); SDBG_STREAM_ATTR(2); SDBG_STREAM_ATTR(3); SDBG_STREAM_ATTR(4);
#undef SDBG_STREAM_ATTR
    CU(c, cudaFuncSetAttribute(bm25_stream_kernel<1, false, 3, kModeAnd>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CU(c, cudaFuncSetAttribute(bm25_stream_kernel<1, true, 3, kModeAnd>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CU(c, cudaFuncSetAttribute(bm25_stream_kernel<1, false, 3, kModeLead>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CU(c, cudaFuncSetAttribute(bm25_stream_kernel<1, true, 3, kModeLead>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CU(c, cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    c->topk_attr_set = true;
  }
  // one claim word per work item of class kClsLead (zeroed per call)
  size_t n_lead_total = 0;
  for (size_t si = 0; si < n_segs; ++si) n_lead_total += n_cls[si][kClsLead];
  DevBuf& b_claim = c->scratch[11];
  if (n_lead_total) {
    if ((rc = ensure(c, b_claim, n_lead_total * 4))) return rc;
    CU(c, cudaMemsetAsync(b_claim.p, 0, n_lead_total * 4, c->stream));
  }
  uint32_t base = 0;
  size_t work_done = 0, lead_done = 0;
  // Work classes are separate launches; with more than one present they alternate between two streams (forked from
  // and joined back into the context's stream) so that no class waits for another's tail. First slices (class 11) go
  // first, and everything that depends on their thresholds is ordered behind them.
  uint32_t classes_present = 0;
  for (size_t si = 0; si < n_segs; ++si) for (uint32_t k2 = 0; k2 < kClasses; ++k2) if (n_cls[si][k2]) classes_present |= 1u << k2;
  const bool two_lanes = (classes_present & (classes_present - 1u)) != 0u;
  {
    ProfScope ps_(c, kProfTopk);   // one span for all top-k launches of the call
    uint32_t lane_no = 0;
    for (size_t si = 0; si < n_segs; ++si) {
      sdbg_segment* s = segs[si];
      TopkParams P;
      P.seg = postings_view(s, base);
      if ((rc = filter_view(s, filt, &P.filt))) return rc;
      P.qterms = static_cast<const QTermDev*>(b_qt.p) + si * total_terms;
      P.qterm_off = reinterpret_cast<const uint32_t*>(static_cast<const char*>(b_qt.p) + qt_bytes);
      P.theta = d_theta; P.total = d_total;
      P.cand = static_cast<unsigned long long*>(b_cand.p);
      P.cand_n = static_cast<uint32_t*>(b_candn.p);
      P.k = k; P.cap = pl.cap; P.conjunction = kind == SDBG_QUERY_AND ? 1 : 0;
      P.claim = nullptr;
      const int wand = seg_wand(s);
      const bool lut = use_lut[si];
      const uint4* const work0 = reinterpret_cast<const uint4*>(static_cast<const char*>(b_qt.p) + qt_pad) + work_done;
      work_done += seg_work[si].size();
      std::array<size_t, kClasses> cls_off{};
      { size_t o = 0; for (uint32_t cls = 0; cls < kClasses; ++cls) { cls_off[cls] = o; o += n_cls[si][cls]; } }
      auto launch_merge = [&](uint32_t T, size_t n, cudaStream_t st) {
        const size_t sm = size_t(pl.cap) * 8 + (lut ? size_t(T) * kLutFreqs * 1024 : 0) + size_t(kTopkWarps) * T * kMergeTermBytes;
#define SDBG_MERGE_LAUNCH(TT) \
        if (lut) bm25_merge_kernel<TT, true><<<unsigned(n), kTopkThreads, sm, st>>>(P); \
        else bm25_merge_kernel<TT, false><<<unsigned(n), kTopkThreads, sm, st>>>(P)
        switch (T) {
          case 1: SDBG_MERGE_LAUNCH(1); break;
          case 2: SDBG_MERGE_LAUNCH(2); break;
          case 3: SDBG_MERGE_LAUNCH(3); break;
          default: SDBG_MERGE_LAUNCH(4); break;
        }
#undef SDBG_MERGE_LAUNCH
        ++c->launches;
      };
      // first slices: before everything else of this segment, on the main stream
      if (n_cls[si][kClsSlice]) {
        P.work = work0 + cls_off[kClsSlice]; P.wand = 0;
        launch_merge(2, n_cls[si][kClsSlice], c->stream);
      }
      if (two_lanes) {
        CU(c, cudaEventRecord(c->ev_fork, c->stream));
        CU(c, cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
      }
      for (uint32_t cls = 0; cls < kClasses; ++cls) {
        const size_t n = n_cls[si][cls];
        if (!n || cls == kClsSlice) continue;
        cudaStream_t st = (two_lanes && (lane_no++ & 1u)) ? c->stream2 : c->stream;
        P.work = work0 + cls_off[cls];
        P.claim = nullptr;
        if (cls == 0) {
          P.wand = wand;
          if (pl.budget == 16) bm25_topk_kernel<16, true><<<unsigned(n), kTopkThreads, smem_drive, st>>>(P);
          else bm25_topk_kernel<32, true><<<unsigned(n), kTopkThreads, smem_drive, st>>>(P);
          ++c->launches;
        } else if (cls == 1) {
          P.wand = std::min(wand, 1);
          if (pl.budget == 16) bm25_topk_kernel<16, false><<<unsigned(n), kTopkThreads, pl.smem, st>>>(P);
          else bm25_topk_kernel<32, false><<<unsigned(n), kTopkThreads, pl.smem, st>>>(P);
          ++c->launches;
        } else if (cls == kClsAnd) {
          const size_t sm = stream_smem(1, lut, true);
          P.wand = 0;                                  // conjunctions are exact: every candidate of the lead list is probed
          if (lut) bm25_stream_kernel<1, true, 3, kModeAnd><<<unsigned(n), kTopkThreads, sm, st>>>(P);
          else bm25_stream_kernel<1, false, 3, kModeAnd><<<unsigned(n), kTopkThreads, sm, st>>>(P);
          ++c->launches;
        } else if (cls == kClsLead) {
          // both kernels over the same items; each item is run by exactly one of them (claim word)
          P.claim = static_cast<uint32_t*>(b_claim.p) + lead_done;
          lead_done += n;
          P.wand = 0;
          launch_merge(2, n, c->stream);
          const size_t sm = stream_smem(1, lut, true);
          P.wand = wand | (env_int("SDBG_STREAM_DBG", 0) & 0xF0);
          cudaStream_t st2 = two_lanes ? c->stream2 : c->stream;
          if (lut) bm25_stream_kernel<1, true, 3, kModeLead><<<unsigned(n), kTopkThreads, sm, st2>>>(P);
          else bm25_stream_kernel<1, false, 3, kModeLead><<<unsigned(n), kTopkThreads, sm, st2>>>(P);
          ++c->launches;
        } else if (cls >= kClsStream) {
          const uint32_t T = cls - kClsStream + 1u;
          const size_t sm = stream_smem(T, lut, false);
          P.wand = wand ? (wand | (env_int("SDBG_STREAM_DBG", 0) & 0xF0)) : 0;   // debug bits 16/32/64/128 switch parts of the pruning off
#define SDBG_STREAM_LAUNCH(TT) \
          if (lut) bm25_stream_kernel<TT, true, 3, kModeOr><<<unsigned(n), kTopkThreads, sm, st>>>(P); \
          else bm25_stream_kernel<TT, false, 3, kModeOr><<<unsigned(n), kTopkThreads, sm, st>>>(P)
          switch (T) {
            case 1: SDBG_STREAM_LAUNCH(1); break;
            case 2: SDBG_STREAM_LAUNCH(2); break;
            case 3: SDBG_STREAM_LAUNCH(3); break;
            default: SDBG_STREAM_LAUNCH(4); break;
          }
#undef SDBG_STREAM_LAUNCH
          ++c->launches;
        } else {
          P.wand = 0;
          launch_merge(cls - kClsMerge + 1u, n, st);
        }
      }
      CU(c, cudaGetLastError());
      base += s->n_docs;
      if (two_lanes) {
        CU(c, cudaEventRecord(c->ev_join, c->stream2));
        CU(c, cudaStreamWaitEvent(c->stream, c->ev_join, 0));
      }
    }
  }
  MergeParams M;
  M.cand = static_cast<const unsigned long long*>(b_cand.p);
  M.cand_n = static_cast<const uint32_t*>(b_candn.p);
  M.list_off = reinterpret_cast<const uint32_t*>(static_cast<const char*>(b_qt.p) + qt_pad + work_bytes);
  M.G = 0; M.stride = pl.cap; M.k = k; M.cap = pl.cap;
  M.keys_out = static_cast<unsigned long long*>(b_keys.p);
  M.n_out = static_cast<uint32_t*>(b_small.p);
  { ProfScope ps_(c, kProfMerge);
    topk_merge_kernel<<<unsigned(nq), kTopkThreads, size_t(pl.cap) * 8, c->stream>>>(M); }
  ++c->launches;
  CU(c, cudaGetLastError());
  dev->keys = M.keys_out; dev->n_out = M.n_out; dev->total = d_total;
  return SDBG_OK;
}

// keys -> hits on the host. `bases` = first ordinal of each segment.
void keys_to_hits(const unsigned long long* keys, uint32_t n, const std::vector<uint64_t>& bases, sdbg_hit* out) {
  if (bases.size() == 1) {                      // one segment: no search for the owner of an ordinal
    const uint32_t b0 = uint32_t(bases[0]);
    for (uint32_t i = 0; i < n; ++i) {
      const uint32_t bits = uint32_t(keys[i] >> 32);
      std::memcpy(&out[i].score, &bits, 4);
      out[i].seg = 0;
      out[i].doc = ~uint32_t(keys[i]) - b0;
    }
    return;
  }
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t bits = uint32_t(keys[i] >> 32);
    const uint32_t ordinal = ~uint32_t(keys[i]);
    size_t seg = std::upper_bound(bases.begin(), bases.end(), uint64_t(ordinal) - 1) - bases.begin() - 1;
    std::memcpy(&out[i].score, &bits, 4);
    out[i].seg = uint32_t(seg);
    out[i].doc = uint32_t(ordinal - bases[seg]);
  }
}

}  // namespace

extern "C" int sdbg_bm25_collect(uint64_t docs_with_field, uint64_t total_term_freq, uint64_t docs_with_term, float k, float b,
                                 sdbg_bm25_term* out) {
  if (!out || docs_with_term > docs_with_field) return SDBG_EINVAL;
  // bm25.cpp:288-309, operation for operation (host code is built with -ffp-contract=off)
  out->idf = float(std::log1p((double(docs_with_field - docs_with_term) + 0.5) / (double(docs_with_term) + 0.5)));
  const float kb = k * b;
  out->norm_const = k - kb;
  if (total_term_freq && docs_with_field) {
    const float avg_dl = float(total_term_freq) / float(docs_with_field);
    out->norm_length = kb / avg_dl;
  } else {
    out->norm_length = kb;
  }
  out->boost = 1.f;
  return SDBG_OK;
}

// TFIDF (search/tfidf.cpp): idf = (float) log1p((docs_with_field + 1.0) / (docs_with_term + 1.0)), :149-150.
extern "C" int sdbg_tfidf_collect(uint64_t docs_with_field, uint64_t docs_with_term, sdbg_bm25_term* out) {
  if (!out) return SDBG_EINVAL;
  out->idf = float(std::log1p((double(docs_with_field) + 1.0) / (double(docs_with_term) + 1.0)));
  out->norm_const = 0.f; out->norm_length = 0.f; out->boost = 1.f;
  return SDBG_OK;
}

extern "C" int sdbg_bm25_topk_batch(sdbg_segment* const* segs, size_t n_segs, int kind, const sdbg_bm25_term* terms,
                                    const uint32_t* term_off, size_t nq, float k1, float b, const sdbg_col_pred* filt,
                                    uint32_t k, float threshold_in, sdbg_hit* out, uint32_t* n_out, uint64_t* total_matches);
// Same scan, scored with TFIDF: sqrt(freq) * boost * idf, divided by sqrt(doc length) when `normalize` (tfidf.cpp:59-80).
// Exhaustive (the segment's block-max entries belong to BM25).
extern "C" int sdbg_tfidf_topk_batch(sdbg_segment* const* segs, size_t n_segs, int kind, const sdbg_bm25_term* terms,
                                     const uint32_t* term_off, size_t nq, int normalize, const sdbg_col_pred* filt, uint32_t k,
                                     float threshold_in, sdbg_hit* out, uint32_t* n_out, uint64_t* total_matches) {
  return sdbg_bm25_topk_batch(segs, n_segs, kind, terms, term_off, nq, kTfidfK1, normalize ? 1.f : 0.f, filt, k, threshold_in, out, n_out,
                              total_matches);
}

extern "C" int sdbg_bm25_topk_batch(sdbg_segment* const* segs, size_t n_segs, int kind, const sdbg_bm25_term* terms,
                                    const uint32_t* term_off, size_t nq, float k1, float b, const sdbg_col_pred* filt,
                                    uint32_t k, float threshold_in, sdbg_hit* out, uint32_t* n_out,
                                    uint64_t* total_matches) {
  if (!out || !n_out) return SDBG_EINVAL;
  TopkDevOut dev{};
  int rc = topk_run(segs, n_segs, kind, terms, term_off, nq, k1, b, filt, k, threshold_in, &dev);
  if (rc) return rc;
  sdbg_ctx* c = segs[0]->ctx;
  const size_t kb = nq * size_t(k) * 8, nb = nq * 4, tb = nq * 8;
  // results come back through a dedicated pinned block (the query descriptors are done with by now)
  if ((rc = ensure_pinned(c, kb + nb + tb + 64))) return rc;
  char* h = static_cast<char*>(c->h_pinned);
  // key -> {segment, doc, score} is a few ns per hit; a large batch (millions of hits) is cut into per-thread query
  // ranges whose keys are copied back one after the other, each followed by an event: a thread converts its range as
  // soon as it has landed, while the later ranges are still on the wire.
  const size_t n_thr = std::max<size_t>(1, std::min<size_t>({size_t(env_int("SDBG_HOST_THREADS", 16)), (nq * size_t(k)) / 65536, size_t(kMaxCopyEvents)}));
  CU(c, cudaMemcpyAsync(h + kb, dev.total, tb, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(h + kb + tb, dev.n_out, nb, cudaMemcpyDeviceToHost, c->stream));
  for (size_t t = 0; t < n_thr; ++t) {
    const size_t q0 = nq * t / n_thr, q1 = nq * (t + 1) / n_thr;
    CU(c, cudaMemcpyAsync(h + q0 * k * 8, dev.keys + q0 * k, (q1 - q0) * k * 8, cudaMemcpyDeviceToHost, c->stream));
    if (n_thr > 1) {
      if (!c->ev_copy[t]) CU(c, cudaEventCreateWithFlags(&c->ev_copy[t], cudaEventDisableTiming));
      CU(c, cudaEventRecord(c->ev_copy[t], c->stream));
    }
  }
  std::vector<uint64_t> bases(n_segs);
  uint64_t ord0 = 0;
  for (size_t si = 0; si < n_segs; ++si) { bases[si] = ord0; ord0 += segs[si]->n_docs; }
  const auto* keys = reinterpret_cast<const unsigned long long*>(h);
  const auto* tot = reinterpret_cast<const unsigned long long*>(h + kb);
  const auto* cnt = reinterpret_cast<const uint32_t*>(h + kb + tb);
  auto convert = [&](size_t q0, size_t q1) {
    for (size_t q = q0; q < q1; ++q) {
      n_out[q] = cnt[q];
      keys_to_hits(keys + q * k, cnt[q], bases, out + q * k);
      if (total_matches) total_matches[q] = tot[q];
    }
  };
  if (n_thr <= 1) {
    CU(c, cudaStreamSynchronize(c->stream));
    convert(0, nq);
  } else {
    std::atomic<int> err{0};
    std::vector<std::thread> pool;
    for (size_t t = 0; t < n_thr; ++t)
      pool.emplace_back([&, t] {
        if (cudaSetDevice(c->device) != cudaSuccess || cudaEventSynchronize(c->ev_copy[t]) != cudaSuccess) { err = 1; return; }
        convert(nq * t / n_thr, nq * (t + 1) / n_thr);
      });
    for (auto& th : pool) th.join();
    CU(c, cudaStreamSynchronize(c->stream));
    if (err) return fail(c, SDBG_ECUDA, "copying the hits back failed");
  }
  return SDBG_OK;
}

// Streaming mode (duckdb_search_full_scan.cpp RunStreamingScan :2370-2403; DocIterator::EmitScoredDocs,
// iterators.hpp:202-204): every match of one query in docs [doc_min, doc_max) of one segment with its score, ascending
// by doc. Same stream kernel as the top-k scan with pruning off; its sink writes through a global cursor and the
// pairs are then radix-sorted by doc id on the device.
extern "C" int sdbg_bm25_scan(sdbg_segment* s, int kind, const sdbg_bm25_term* terms, size_t n_terms, float k1, float b,
                              const sdbg_col_pred* filt, uint32_t doc_min, uint32_t doc_max, uint32_t* out_docs, float* out_scores,
                              uint64_t cap, uint64_t* n_out) {
  if (!s || !terms || !n_terms || !n_out || (cap && (!out_docs || !out_scores))) return SDBG_EINVAL;
  sdbg_ctx* c = s->ctx;
  CU(c, cudaSetDevice(c->device));
  if (!s->d_blocks) return fail(c, SDBG_EINVAL, "segment has no staged postings");
  const bool conj = kind == SDBG_QUERY_AND;
  if (n_terms > (conj ? size_t(kMaxQueryTerms) : size_t(kStreamMaxTerms)))
    return fail(c, SDBG_EUNSUPPORTED, "scored scan: disjunctions take 1..4 terms, conjunctions 1..16");
  if (k1 == 0.f || b == 0.f || k1 == kTfidfK1) return fail(c, SDBG_EUNSUPPORTED, "scored scan: BM25 form only");
  *n_out = 0;
  doc_min = std::max(doc_min, 1u);                                   // doc ids start at doc_limits::min()
  doc_max = uint32_t(std::min<uint64_t>(doc_max, uint64_t(s->n_docs) + 1u));
  if (doc_min >= doc_max) return SDBG_OK;
  const uint32_t range = doc_max - doc_min;
  const uint32_t T = uint32_t(n_terms);
  const uint32_t scan_cap = 1024;                                    // candidate buffer of the kernel: unused here, kept minimal
  const uint32_t g = std::max(1u, std::min(uint32_t(c->sm_count) * 3u, range / 4096u));
  const uint32_t chunk = (range + g - 1) / g;
  // descriptors: [QTermDev x T][term_off x 2][pad][work x g]
  const size_t qt_bytes = size_t(T) * sizeof(QTermDev), off_bytes = 2 * sizeof(uint32_t);
  const size_t qt_pad = (qt_bytes + off_bytes + 15) & ~size_t(15), work_bytes = size_t(g) * sizeof(uint4);
  int rc = ensure_pinned(c, qt_pad + work_bytes);
  if (rc) return rc;
  auto* h_qt = static_cast<QTermDev*>(c->h_pinned);
  for (uint32_t i = 0; i < T; ++i) {
    if (terms[i].term + 1 >= s->term_blk_begin.size()) return fail(c, SDBG_EINVAL, "term id out of range");
    fill_qterm(s, terms[i], k1, b, h_qt[i]);
  }
  std::stable_sort(h_qt, h_qt + T, [](const QTermDev& x, const QTermDev& y) { return x.docs_count < y.docs_count; });
  auto* h_off = reinterpret_cast<uint32_t*>(static_cast<char*>(c->h_pinned) + qt_bytes);
  h_off[0] = 0; h_off[1] = T;
  auto* h_work = reinterpret_cast<uint4*>(static_cast<char*>(c->h_pinned) + qt_pad);
  for (uint32_t j = 0; j < g; ++j) {
    const uint32_t lo = doc_min + j * chunk;
    h_work[j] = make_uint4(0u, lo, lo < doc_max ? std::min(chunk, doc_max - lo) : 0u, j);
    if (lo >= doc_max) h_work[j].y = s->n_docs + 1u;               // empty chain
  }
  DevBuf& b_qt = c->scratch[0]; DevBuf& b_theta = c->scratch[1]; DevBuf& b_cand = c->scratch[2]; DevBuf& b_candn = c->scratch[3];
  DevBuf& b_emit = c->scratch[12];
  const uint64_t room = std::max<uint64_t>(cap, 1);
  size_t sort_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, static_cast<const uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr),
                                  static_cast<const float*>(nullptr), static_cast<float*>(nullptr), room, 0, 32, c->stream);
  const size_t pair_bytes = (size_t(room) * 4 + 255) & ~size_t(255);
  if ((rc = ensure(c, b_qt, qt_pad + work_bytes))) return rc;
  if ((rc = ensure(c, b_theta, 32))) return rc;                      // theta | total | cursor
  if ((rc = ensure(c, b_cand, size_t(g) * scan_cap * 8))) return rc;
  if ((rc = ensure(c, b_candn, size_t(g) * 4))) return rc;
  if ((rc = ensure(c, b_emit, 4 * pair_bytes + sort_bytes))) return rc;
  CU(c, cudaMemcpyAsync(b_qt.p, c->h_pinned, qt_pad + work_bytes, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaMemsetAsync(b_theta.p, 0, 32, c->stream));
  TopkParams P;
  P.seg = postings_view(s, 0);
  if ((rc = filter_view(s, filt, &P.filt))) return rc;
  P.qterms = static_cast<const QTermDev*>(b_qt.p);
  P.qterm_off = reinterpret_cast<const uint32_t*>(static_cast<const char*>(b_qt.p) + qt_bytes);
  P.work = reinterpret_cast<const uint4*>(static_cast<const char*>(b_qt.p) + qt_pad);
  P.theta = static_cast<unsigned long long*>(b_theta.p);
  P.total = P.theta + 1;
  P.emit_count = P.theta + 2;
  P.cand = static_cast<unsigned long long*>(b_cand.p);
  P.cand_n = static_cast<uint32_t*>(b_candn.p);
  P.claim = nullptr;
  P.k = 1; P.cap = scan_cap; P.conjunction = conj ? 1 : 0; P.wand = 0;
  char* e = static_cast<char*>(b_emit.p);
  P.emit_docs = reinterpret_cast<uint32_t*>(e);
  P.emit_scores = reinterpret_cast<float*>(e + pair_bytes);
  P.emit_cap = cap;
  auto* sorted_docs = reinterpret_cast<uint32_t*>(e + 2 * pair_bytes);
  auto* sorted_scores = reinterpret_cast<float*>(e + 3 * pair_bytes);
  const bool lut = s->norm_width == 1 && env_int("SDBG_STREAM_LUT", 0) != 0;
  const uint32_t Tl = conj ? 1u : T;
  const size_t sm = size_t(scan_cap) * 8 + (lut ? size_t(Tl) * kLutFreqs * 1024 : 0) + size_t(kTopkWarps) * (Tl * kStreamTermBytes + (conj ? 1024 : 0));
  if (!c->scan_attr_set) {
#define SDBG_SCAN_ATTR(TT) \
    CU(c, cudaFuncSetAttribute(bm25_stream_kernel<TT, false, 3, kModeOr>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
    CU(c, cudaFuncSetAttribute(bm25_stream_kernel<TT, true, 3, kModeOr>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024))
    SDBG_SCAN_ATTR(1); SDBG_SCAN_ATTR(2); SDBG_SCAN_ATTR(3); SDBG_SCAN_ATTR(4);
#undef SDBG_SCAN_ATTR
    CU(c, cudaFuncSetAttribute(bm25_stream_kernel<1, false, 3, kModeAnd>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CU(c, cudaFuncSetAttribute(bm25_stream_kernel<1, true, 3, kModeAnd>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    c->scan_attr_set = true;
  }
  {
    ProfScope ps_(c, kProfTopk);
    if (conj) {
      if (lut) bm25_stream_kernel<1, true, 3, kModeAnd><<<g, kTopkThreads, sm, c->stream>>>(P);
      else bm25_stream_kernel<1, false, 3, kModeAnd><<<g, kTopkThreads, sm, c->stream>>>(P);
    } else {
#define SDBG_SCAN_LAUNCH(TT) \
      if (lut) bm25_stream_kernel<TT, true, 3, kModeOr><<<g, kTopkThreads, sm, c->stream>>>(P); \
      else bm25_stream_kernel<TT, false, 3, kModeOr><<<g, kTopkThreads, sm, c->stream>>>(P)
      switch (T) {
        case 1: SDBG_SCAN_LAUNCH(1); break;
        case 2: SDBG_SCAN_LAUNCH(2); break;
        case 3: SDBG_SCAN_LAUNCH(3); break;
        default: SDBG_SCAN_LAUNCH(4); break;
      }
#undef SDBG_SCAN_LAUNCH
    }
  }
  ++c->launches;
  CU(c, cudaGetLastError());
  unsigned long long found = 0;
  CU(c, cudaMemcpyAsync(&found, P.emit_count, 8, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  *n_out = found;
  if (found > cap) return fail(c, SDBG_ECAPACITY, "scored scan: more matches than the output has room for (*n_out = needed)");
  if (!found) return SDBG_OK;
  cub::DeviceRadixSort::SortPairs(e + 4 * pair_bytes, sort_bytes, P.emit_docs, sorted_docs, P.emit_scores, sorted_scores, found, 0, 32,
                                  c->stream);
  ++c->launches;
  CU(c, cudaGetLastError());
  CU(c, cudaMemcpyAsync(out_docs, sorted_docs, size_t(found) * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(out_scores, sorted_scores, size_t(found) * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return SDBG_OK;
}

extern "C" int sdbg_bm25_topk(sdbg_segment* const* segs, size_t n_segs, int kind, const sdbg_bm25_term* terms,
                              size_t n_terms, float k1, float b, const sdbg_col_pred* filt, uint32_t k, float threshold_in,
                              sdbg_hit* out, uint32_t* n_out, uint64_t* total_matches, float* threshold_out) {
  const uint32_t off[2] = {0, uint32_t(n_terms)};
  uint64_t tot = 0;
  const int rc = sdbg_bm25_topk_batch(segs, n_segs, kind, terms, off, 1, k1, b, filt, k, threshold_in, out, n_out, &tot);
  if (rc) return rc;
  if (total_matches) *total_matches = tot;
  if (threshold_out) *threshold_out = (*n_out == k) ? out[k - 1].score : threshold_in;
  return SDBG_OK;
}

namespace {
__global__ void shift_keys_kernel(const unsigned long long* __restrict__ in, unsigned long long* __restrict__ out, size_t n,
                                  uint32_t add) {
  // Re-bases ordinals for a cross-rank gather: ordinal' = ordinal + add (keys keep their order within a rank).
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const unsigned long long k = in[i];
    out[i] = k ? ((k & 0xFFFFFFFF00000000ull) | static_cast<unsigned long long>(~(~uint32_t(k) + add))) : 0ull;
  }
}
}  // namespace

namespace {
int topk_batch_device_impl(sdbg_segment* const* segs, size_t n_segs, int kind, const sdbg_bm25_term* terms, const uint32_t* term_off,
                           size_t nq, float k1, float b, const sdbg_col_pred* filt, uint32_t k, float threshold_in, uint32_t rank,
                           void* d_keys, void* d_totals, bool sync);
}
extern "C" int sdbg_bm25_topk_batch_device(sdbg_segment* const* segs, size_t n_segs, int kind, const sdbg_bm25_term* terms,
                                           const uint32_t* term_off, size_t nq, float k1, float b, const sdbg_col_pred* filt,
                                           uint32_t k, float threshold_in, uint32_t rank, void* d_keys, void* d_totals) {
  return topk_batch_device_impl(segs, n_segs, kind, terms, term_off, nq, k1, b, filt, k, threshold_in, rank, d_keys, d_totals, true);
}
namespace {
int topk_batch_device_impl(sdbg_segment* const* segs, size_t n_segs, int kind, const sdbg_bm25_term* terms, const uint32_t* term_off,
                           size_t nq, float k1, float b, const sdbg_col_pred* filt, uint32_t k, float threshold_in, uint32_t rank,
                           void* d_keys, void* d_totals, bool sync) {
  if (!d_keys) return SDBG_EINVAL;
  TopkDevOut dev{};
  int rc = topk_run(segs, n_segs, kind, terms, term_off, nq, k1, b, filt, k, threshold_in, &dev);
  if (rc) return rc;
  sdbg_ctx* c = segs[0]->ctx;
  // Each rank owns a 2^28-ordinal slot in the merged key space: rank r's docs sort after rank r-1's on ties.
  uint64_t docs = 0;
  for (size_t si = 0; si < n_segs; ++si) docs += segs[si]->n_docs;
  if (docs >= (1ull << 28) || rank >= 15) return fail(c, SDBG_EUNSUPPORTED, "rank slot overflow (>= 2^28 docs per rank)");
  shift_keys_kernel<<<256, 256, 0, c->stream>>>(dev.keys, static_cast<unsigned long long*>(d_keys), nq * size_t(k), rank << 28);
  ++c->launches;
  CU(c, cudaGetLastError());
  if (d_totals) CU(c, cudaMemcpyAsync(d_totals, dev.total, nq * 8, cudaMemcpyDeviceToDevice, c->stream));
  if (sync) CU(c, cudaStreamSynchronize(c->stream));
  return SDBG_OK;
}
}  // namespace

extern "C" int sdbg_topk_merge_gathered(sdbg_ctx* c, const void* d_keys_all, uint32_t n_ranks, size_t nq, uint32_t k,
                                        sdbg_hit* out, uint32_t* n_out) {
  if (!c || !d_keys_all || !n_ranks || !nq || !k || (out && !n_out)) return SDBG_EINVAL;   // out == NULL: n_out != NULL asks for a sync
  CU(c, cudaSetDevice(c->device));
  // gathered layout [rank][query][k]; the merge kernel wants [query][list][stride] -> stride trick:
  // treat each rank's block as a list with a rank-major base pointer. Re-pack with a tiny kernel-free
  // copy: n_ranks strided memcpy2D calls.
  DevBuf& b_in = c->scratch[6]; DevBuf& b_keys = c->scratch[7]; DevBuf& b_small = c->scratch[8];
  int rc;
  if ((rc = ensure(c, b_in, nq * size_t(n_ranks) * k * 8))) return rc;
  if ((rc = ensure(c, b_keys, nq * size_t(k) * 8))) return rc;
  if ((rc = ensure(c, b_small, nq * 4))) return rc;
  for (uint32_t r = 0; r < n_ranks; ++r)
    CU(c, cudaMemcpy2DAsync(static_cast<char*>(b_in.p) + size_t(r) * k * 8, size_t(n_ranks) * k * 8,
                            static_cast<const char*>(d_keys_all) + size_t(r) * nq * k * 8, size_t(k) * 8, size_t(k) * 8, nq,
                            cudaMemcpyDeviceToDevice, c->stream));
  const uint32_t cap = std::max(next_pow2(k + 1024), 4096u);
  if (!c->merge_attr_set) {
    CU(c, cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    c->merge_attr_set = true;
  }
  MergeParams M;
  M.cand = static_cast<const unsigned long long*>(b_in.p); M.cand_n = nullptr; M.list_off = nullptr;
  M.G = n_ranks; M.stride = k; M.k = k; M.cap = cap;
  M.keys_out = static_cast<unsigned long long*>(b_keys.p); M.n_out = static_cast<uint32_t*>(b_small.p);
  topk_merge_kernel<<<unsigned(nq), kTopkThreads, size_t(cap) * 8, c->stream>>>(M);
  ++c->launches;
  CU(c, cudaGetLastError());
  if (!out && !n_out) return SDBG_OK;                                       // enqueue only: results stay in HBM, nothing waits
  if (!out) { CU(c, cudaStreamSynchronize(c->stream)); return SDBG_OK; }   // results stay in HBM (scratch of this context)
  const size_t kb = nq * size_t(k) * 8, nb = nq * 4;
  if ((rc = ensure_pinned(c, kb + nb))) return rc;
  char* h = static_cast<char*>(c->h_pinned);
  CU(c, cudaMemcpyAsync(h, M.keys_out, kb, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(h + kb, M.n_out, nb, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  const auto* keys = reinterpret_cast<const unsigned long long*>(h);
  const auto* cnt = reinterpret_cast<const uint32_t*>(h + kb);
  auto convert = [&](size_t q0, size_t q1) {
    for (size_t q = q0; q < q1; ++q) {
      n_out[q] = cnt[q];
      for (uint32_t i = 0; i < cnt[q]; ++i) {
        const unsigned long long key = keys[q * k + i];
        const uint32_t bits = uint32_t(key >> 32), ordinal = ~uint32_t(key);
        sdbg_hit& h2 = out[q * k + i];
        std::memcpy(&h2.score, &bits, 4);
        h2.seg = ordinal >> 28;             // rank slot
        h2.doc = ordinal & ((1u << 28) - 1);  // ordinal within the rank (segment base + doc)
      }
    }
  };
  // a few ns per hit; a large batch (millions of hits) is split over host threads like sdbg_bm25_topk_batch does
  const size_t n_thr = std::min<size_t>(size_t(env_int("SDBG_HOST_THREADS", 8)), (nq * size_t(k)) / 65536);
  if (n_thr <= 1) {
    convert(0, nq);
  } else {
    std::vector<std::thread> pool;
    for (size_t t = 0; t < n_thr; ++t) pool.emplace_back(convert, nq * t / n_thr, nq * (t + 1) / n_thr);
    for (auto& th : pool) th.join();
  }
  return SDBG_OK;
}

extern "C" int sdbg_decode_score_term(sdbg_segment* s, uint32_t term, float c0, float nc, float nl, uint32_t* docs,
                                      uint32_t* freqs, float* scores) {
  if (!s || !docs || !freqs || !scores) return SDBG_EINVAL;
  sdbg_ctx* c = s->ctx;
  if (term + 1 >= s->term_blk_begin.size()) return fail(c, SDBG_EINVAL, "term id out of range");
  CU(c, cudaSetDevice(c->device));
  const uint32_t b0 = s->term_blk_begin[term], nblk = s->term_blk_begin[term + 1] - b0;
  const uint32_t n = s->term_docs[term];
  if (!n) return SDBG_OK;
  const size_t slots = size_t(nblk) * 128;
  int rc;
  if ((rc = ensure(c, c->scratch[9], slots * 12))) return rc;
  auto* d_docs = static_cast<uint32_t*>(c->scratch[9].p);
  auto* d_freqs = d_docs + slots;
  auto* d_scores = reinterpret_cast<float*>(d_freqs + slots);
  const unsigned grid = unsigned(std::min<uint32_t>((nblk + kTopkWarps - 1) / kTopkWarps, uint32_t(c->sm_count) * 8u));
  decode_score_kernel<<<grid, kTopkThreads, 0, c->stream>>>(postings_view(s, 0), b0, nblk, c0, nc, nl, d_docs, d_freqs, d_scores);
  ++c->launches;
  CU(c, cudaGetLastError());
  CU(c, cudaMemcpyAsync(docs, d_docs, size_t(n) * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(freqs, d_freqs, size_t(n) * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(scores, d_scores, size_t(n) * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return SDBG_OK;
}

// ------------------------------------------------------------------------------------------
// columnar
// ------------------------------------------------------------------------------------------
namespace {

int col_view(sdbg_segment* s, uint64_t field, ColDev* out, uint64_t* rows) {
  auto it = s->cols.find(field);
  if (it == s->cols.end()) return fail(s->ctx, SDBG_ENOTFOUND, "column " + std::to_string(field) + " not staged");
  out->values = it->second.d_values; out->validity = it->second.d_validity; out->type = it->second.type; out->pad = 0;
  if (rows) *rows = it->second.rows;
  return SDBG_OK;
}

int pred_set(sdbg_segment* s, const sdbg_col_pred* preds, size_t n, PredSet* ps, uint64_t* rows) {
  if (n > size_t(kMaxPreds)) return fail(s->ctx, SDBG_EUNSUPPORTED, "more than 4 pushed predicates");
  std::memset(ps, 0, sizeof *ps);
  ps->n = int(n);
  for (size_t i = 0; i < n; ++i) {
    uint64_t r = 0;
    const int rc = col_view(s, preds[i].field, &ps->p[i].col, &r);
    if (rc) return rc;
    if (*rows == 0) *rows = r;
    if (r != *rows) return fail(s->ctx, SDBG_EINVAL, "columns of one segment differ in length");
    if (preds[i].op < 0 || preds[i].op > 8) return fail(s->ctx, SDBG_EINVAL, "bad predicate op");
    ps->p[i].op = preds[i].op;
    ps->p[i].lo_i = preds[i].lo_i; ps->p[i].hi_i = preds[i].hi_i;
    ps->p[i].lo_f = preds[i].lo_f; ps->p[i].hi_f = preds[i].hi_f;
  }
  return SDBG_OK;
}

int column_minmax(sdbg_segment* s, uint64_t field, int64_t* mn, int64_t* mx) {
  sdbg_ctx* c = s->ctx;
  auto it = s->cols.find(field);
  if (it == s->cols.end()) return fail(c, SDBG_ENOTFOUND, "column not staged");
  ColumnObj& col = it->second;
  if (col.type == SDBG_F64) return fail(c, SDBG_EINVAL, "min/max statistics are kept for integer columns");
  if (!col.has_minmax) {
    int rc;
    if ((rc = ensure(c, c->scratch[10], 16))) return rc;
    const long long init[2] = {INT64_MAX, INT64_MIN};
    CU(c, cudaMemcpyAsync(c->scratch[10].p, init, 16, cudaMemcpyHostToDevice, c->stream));
    ColDev cd; cd.values = col.d_values; cd.validity = col.d_validity; cd.type = col.type; cd.pad = 0;
    minmax_i64_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(cd, col.rows, static_cast<long long*>(c->scratch[10].p));
    ++c->launches;
    CU(c, cudaGetLastError());
    long long res[2];
    CU(c, cudaMemcpyAsync(res, c->scratch[10].p, 16, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    col.mn = res[0]; col.mx = res[1]; col.has_minmax = true;
  }
  *mn = col.mn; *mx = col.mx;
  return SDBG_OK;
}

// Largest magnitude of a double column (raw bits), computed once per column like the integer min/max.
int column_absmax(sdbg_segment* s, uint64_t field, uint64_t* bits) {
  sdbg_ctx* c = s->ctx;
  auto it = s->cols.find(field);
  if (it == s->cols.end()) return fail(c, SDBG_ENOTFOUND, "column not staged");
  ColumnObj& col = it->second;
  if (col.type != SDBG_F64) return fail(c, SDBG_EINVAL, "absmax statistics are kept for double columns");
  if (!col.has_absmax) {
    int rc;
    if ((rc = ensure(c, c->scratch[10], 16))) return rc;
    CU(c, cudaMemsetAsync(c->scratch[10].p, 0, 16, c->stream));
    ColDev cd; cd.values = col.d_values; cd.validity = col.d_validity; cd.type = col.type; cd.pad = 0;
    absmax_f64_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(cd, col.rows, static_cast<unsigned long long*>(c->scratch[10].p));
    ++c->launches;
    CU(c, cudaGetLastError());
    unsigned long long res = 0;
    CU(c, cudaMemcpyAsync(&res, c->scratch[10].p, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    col.absmax_bits = res; col.has_absmax = true;
  }
  *bits = col.absmax_bits;
  return SDBG_OK;
}

}  // namespace

extern "C" int sdbg_column_minmax_i64(sdbg_segment* s, uint64_t field, int64_t* mn, int64_t* mx) {
  if (!s || !mn || !mx) return SDBG_EINVAL;
  CU(s->ctx, cudaSetDevice(s->ctx->device));
  return column_minmax(s, field, mn, mx);
}

extern "C" int sdbg_filter_bitmap(sdbg_segment* s, const sdbg_col_pred* preds, size_t n_preds, uint64_t* mask_out) {
  if (!s || !mask_out || (!preds && n_preds)) return SDBG_EINVAL;
  sdbg_ctx* c = s->ctx;
  CU(c, cudaSetDevice(c->device));
  PredSet ps; uint64_t rows = 0;
  int rc = pred_set(s, preds, n_preds, &ps, &rows);
  if (rc) return rc;
  if (!rows) rows = s->n_docs;
  if (rows > s->n_docs) return fail(c, SDBG_EINVAL, "filter column longer than the segment: mask_out holds (docs_count + 63) / 64 words");
  const size_t words = (rows + 63) / 64;
  if ((rc = ensure(c, c->scratch[9], words * 8))) return rc;
  const unsigned grid = unsigned(std::min<size_t>((words * 32 + 255) / 256, size_t(c->sm_count) * 8));
  filter_bitmap_kernel<<<std::max(grid, 1u), 256, 0, c->stream>>>(ps, rows, static_cast<unsigned long long*>(c->scratch[9].p));
  ++c->launches;
  CU(c, cudaGetLastError());
  CU(c, cudaMemcpyAsync(mask_out, c->scratch[9].p, words * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return SDBG_OK;
}

extern "C" int sdbg_filter_count_sum(sdbg_segment* const* segs, size_t n_segs, const sdbg_col_pred* preds, size_t n_preds,
                                     uint64_t sum_field, uint64_t* count, int64_t sum_i128[2], double* sum_f64) {
  if (!segs || !n_segs || !count || (!preds && n_preds)) return SDBG_EINVAL;
  sdbg_ctx* c = segs[0]->ctx;
  CU(c, cudaSetDevice(c->device));
  const unsigned grid = unsigned(c->sm_count) * 4u;
  int rc;
  if ((rc = ensure(c, c->scratch[9], (size_t(grid) + 1) * sizeof(CountSumOut) * n_segs + 64))) return rc;
  // Point-query latency: ONE launch per segment and one stream synchronisation. The completion counter is zeroed once
  // (the kernel leaves it at zero), and the final partial is written by the kernel straight into mapped pinned memory.
  if (!c->counter_zeroed || c->scratch[15].cap < 16) {
    if ((rc = ensure(c, c->scratch[15], 16))) return rc;
    CU(c, cudaMemsetAsync(c->scratch[15].p, 0, 16, c->stream));
    c->counter_zeroed = true;
  }
  const size_t seq_off = (n_segs * sizeof(CountSumOut) + 63) & ~size_t(63);   // [results | completion words]
  if (c->h_result_cap < seq_off + n_segs * 8) {
    if (c->h_result) { cudaStreamSynchronize(c->stream); cudaFreeHost(c->h_result); c->h_result = nullptr; }
    const size_t want = std::max<size_t>(seq_off + n_segs * 8, 4096);
    CU(c, cudaHostAlloc(&c->h_result, want, cudaHostAllocMapped));
    CU(c, cudaHostGetDevicePointer(&c->d_result, c->h_result, 0));
    std::memset(c->h_result, 0, want);
    c->h_result_cap = want;
  }
  const unsigned long long seq = ++c->result_seq;
  auto* h_seq = reinterpret_cast<volatile unsigned long long*>(static_cast<char*>(c->h_result) + seq_off);
  auto* d_seq = reinterpret_cast<unsigned long long*>(static_cast<char*>(c->d_result) + seq_off);
  for (size_t si = 0; si < n_segs; ++si) {
    sdbg_segment* s = segs[si];
    PredSet ps; uint64_t rows = 0;
    if ((rc = pred_set(s, preds, n_preds, &ps, &rows))) return rc;
    ColDev sc{}; int has_sum = 0;
    if (sum_field != UINT64_MAX) {
      uint64_t r = 0;
      if ((rc = col_view(s, sum_field, &sc, &r))) return rc;
      if (!rows) rows = r;
      if (r != rows) return fail(c, SDBG_EINVAL, "columns of one segment differ in length");
      has_sum = 1;
    }
    if (!rows) rows = s->n_docs;
    auto* part = static_cast<CountSumOut*>(c->scratch[9].p) + si * (size_t(grid) + 1);
    const unsigned g2 = unsigned(std::max<uint64_t>(1, std::min<uint64_t>(grid, (rows + 2047) / 2048)));   // 8 rows per thread at least
    { ProfScope ps_(c, kProfCountSum);
      filter_count_sum_kernel<<<g2, 256, 0, c->stream>>>(ps, sc, has_sum, rows, part, static_cast<unsigned int*>(c->scratch[15].p),
                                                         static_cast<CountSumOut*>(c->d_result) + si, d_seq + si, seq); }
    ++c->launches;
    CU(c, cudaGetLastError());
  }
  // Wait on the completion words the kernels write into mapped host memory after their result (a PCIe write, ~1 us
  // after the last block finishes) instead of synchronising the stream; the stream is only queried now and then so
  // that a failed launch cannot spin forever.
  for (size_t si = 0; si < n_segs; ++si) {
    uint32_t spins = 0;
    while (__atomic_load_n(const_cast<const unsigned long long*>(&h_seq[si]), __ATOMIC_ACQUIRE) != seq) {
      if ((++spins & 0x3FFFu) == 0u) {
        const cudaError_t q = cudaStreamQuery(c->stream);
        if (q != cudaErrorNotReady && q != cudaSuccess) { CU(c, q); }
        if (q == cudaSuccess && __atomic_load_n(const_cast<const unsigned long long*>(&h_seq[si]), __ATOMIC_ACQUIRE) != seq) {
          CU(c, cudaStreamSynchronize(c->stream));   // finished without the word (cannot happen unless the write was lost)
          break;
        }
      }
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
  }
  unsigned __int128 tot = 0; uint64_t cnt = 0; double sf = 0;
  for (size_t si = 0; si < n_segs; ++si) {
    const CountSumOut& o = static_cast<const CountSumOut*>(c->h_result)[si];
    cnt += o.count; sf += o.sum_f;
    tot += (static_cast<unsigned __int128>(static_cast<uint64_t>(o.sum_hi)) << 64) | o.sum_lo;
  }
  *count = cnt;
  if (sum_i128) { sum_i128[0] = int64_t(uint64_t(tot)); sum_i128[1] = int64_t(uint64_t(tot >> 64)); }
  if (sum_f64) *sum_f64 = sf;
  return SDBG_OK;
}

namespace {

struct GroupPlan { int wide_int = 0; int count_f = 0; int pack_shift = 0; int pack_tables = 0; int64_t pack_bias = 0; int fix_limb = 0; int fix_eunit = 0; int quad = 0; };
constexpr int kGroupByDefaultStages = 3;   // 3 x 20 KB stages -> 3 CTAs (24 consumer warps) per SM; measured best of 2/3/4
constexpr int kGroupByTileRows = 512;   // tile of the default TMA shape; packed accumulators are only planned for it

int groupby_launch(sdbg_segment* const* segs, size_t n_segs, const sdbg_col_pred* preds, size_t n_preds, uint64_t key_field,
                   int64_t key_min, uint64_t span, uint64_t sum_int_field, uint64_t avg_f64_field, void* d_i64, void* d_f64,
                   GroupPlan* plan_out, bool defer_check = false) {
  sdbg_ctx* c = segs[0]->ctx;
  int rc;
  // table | cnt_f | out_of_range
  const size_t table_bytes = span * sizeof(GroupSlot);
  if ((rc = ensure(c, c->scratch[11], table_bytes + span * 8 + 64))) return rc;
  auto* table = static_cast<GroupSlot*>(c->scratch[11].p);
  auto* cnt_f = reinterpret_cast<unsigned long long*>(static_cast<char*>(c->scratch[11].p) + table_bytes);
  auto* oor = cnt_f + span;
  CU(c, cudaMemsetAsync(c->scratch[11].p, 0, table_bytes + span * 8 + 64, c->stream));
  GroupPlan plan;
  uint64_t total_rows = 0;
  int64_t sum_mn = INT64_MAX, sum_mx = INT64_MIN;
  uint64_t absmax_bits = 0;
  bool all_tma = env_int("SDBG_GROUPBY_TMA", 1) != 0 && env_int("SDBG_GROUPBY_TMA_SHAPE", 0) == 0;
  for (size_t si = 0; si < n_segs; ++si) {  // statistics decide the accumulator shape
    sdbg_segment* s = segs[si];
    if (sum_int_field != UINT64_MAX) {
      int64_t mn, mx;
      if ((rc = column_minmax(s, sum_int_field, &mn, &mx))) return rc;
      if (mn < INT32_MIN || mx > INT32_MAX) plan.wide_int = 1;
      sum_mn = std::min(sum_mn, mn); sum_mx = std::max(sum_mx, mx);
      auto sit = s->cols.find(sum_int_field);
      if (sit != s->cols.end() && sit->second.d_validity) all_tma = false;
    }
    for (size_t i = 0; i < n_preds; ++i) {
      auto pit = s->cols.find(preds[i].field);
      if (preds[i].op >= 7 || (pit != s->cols.end() && pit->second.d_validity)) all_tma = false;
    }
    if (avg_f64_field != UINT64_MAX) {
      auto it = s->cols.find(avg_f64_field);
      if (it == s->cols.end()) return fail(c, SDBG_ENOTFOUND, "avg column not staged");
      if (it->second.d_validity) { plan.count_f = 1; all_tma = false; }
      uint64_t ab = 0;
      if (it->second.type == SDBG_F64 && all_tma && env_int("SDBG_GROUPBY_QUAD", 0)) { if ((rc = column_absmax(s, avg_f64_field, &ab))) return rc; }   // statistic only the fixed-point path needs
      absmax_bits = std::max(absmax_bits, ab);
    }
    auto kit = s->cols.find(key_field);
    if (kit == s->cols.end()) return fail(c, SDBG_ENOTFOUND, "key column not staged");
    if (kit->second.d_validity) return fail(c, SDBG_EUNSUPPORTED, "nullable GROUP BY key");
    if (kit->second.type == SDBG_F64) return fail(c, SDBG_EUNSUPPORTED, "float GROUP BY key");
    total_rows += kit->second.rows;
  }
  if (total_rows >= (1ull << 31)) return fail(c, SDBG_EUNSUPPORTED, ">= 2^31 rows per GPU in one GROUP BY (limb overflow guard)");
  // Packed accumulators: one RED carries COUNT and SUM(int) when the column statistics prove that
  // count << shift | sum(v - min) cannot overflow either field (the RED issue rate, not HBM, is what the
  // consumer warps run into). Rows are dealt to 1..3 words of the slot by tile index.
  if (sum_int_field != UINT64_MAX && !plan.wide_int && all_tma && total_rows && sum_mx >= sum_mn && env_int("SDBG_GROUPBY_PACKED", 1)) {
    const unsigned __int128 range = static_cast<unsigned __int128>(static_cast<uint64_t>(sum_mx) - static_cast<uint64_t>(sum_mn));
    const int max_tables = avg_f64_field != UINT64_MAX ? 2 : 3;   // words 2, 3 are kept free for the fixed-point SUM(double) limbs
    for (int nt = std::max(1, env_int("SDBG_GROUPBY_PACK_TABLES_MIN", 1)); nt <= max_tables && !plan.pack_tables; ++nt) {   // env: test hook
      uint64_t cap_rows = 0;   // most rows any one word can receive: its share of every segment's tiles
      for (size_t si = 0; si < n_segs; ++si) {
        const uint64_t tiles = (segs[si]->cols.find(key_field)->second.rows + kGroupByTileRows - 1) / kGroupByTileRows;
        cap_rows += (tiles + nt - 1) / nt * kGroupByTileRows;
      }
      const unsigned __int128 max_sum = range * cap_rows;
      int shift = 1;
      while (shift < 63 && (static_cast<unsigned __int128>(1) << shift) <= max_sum) ++shift;
      if (shift < 63 && cap_rows < (1ull << (64 - shift))) { plan.pack_tables = nt; plan.pack_shift = shift; plan.pack_bias = sum_mn; }
    }
  }
  // Fixed-point SUM(double): two integer limb REDs instead of one floating-point RED (see TmaGroupByParams).
  // Needs words 2 and 3 of the slot (no wide integer sum, at most two packed words) and a finite column.
  if (avg_f64_field != UINT64_MAX && all_tma && !plan.wide_int && total_rows && absmax_bits < 0x7FF0000000000000ull &&
      env_int("SDBG_GROUPBY_FIXED", 1)) {
    int row_bits = 1;
    while ((1ull << row_bits) <= total_rows) ++row_bits;
    plan.fix_limb = std::min(37, 63 - row_bits);                     // |limb sum| <= rows * 2^limb < 2^63
    double mx; std::memcpy(&mx, &absmax_bits, 8);
    int e = 0;
    if (mx > 0) std::frexp(mx, &e);                                  // mx < 2^e
    plan.fix_eunit = e - 2 * plan.fix_limb;                          // |w| / 2^eunit < 2^(2*limb)
  }
  // All accumulators integer => the four words of a slot go out as one RED request per passing row.
  plan.quad = all_tma && (avg_f64_field == UINT64_MAX || plan.fix_limb) && env_int("SDBG_GROUPBY_QUAD", 0);
  if (!plan.quad) plan.fix_limb = 0, plan.fix_eunit = 0;            // separate REDs: one f64 RED beats two integer ones
  for (size_t si = 0; si < n_segs; ++si) {
    sdbg_segment* s = segs[si];
    GroupByParams P;
    std::memset(&P, 0, sizeof P);
    uint64_t rows = 0;
    if ((rc = pred_set(s, preds, n_preds, &P.ps, &rows))) return rc;
    uint64_t r = 0;
    if ((rc = col_view(s, key_field, &P.key, &r))) return rc;
    if (!rows) rows = r;
    if (r != rows) return fail(c, SDBG_EINVAL, "key column length differs");
    if (sum_int_field != UINT64_MAX) {
      if ((rc = col_view(s, sum_int_field, &P.sum_i, &r))) return rc;
      if (r != rows) return fail(c, SDBG_EINVAL, "sum_int column length differs");
      if (P.sum_i.type == SDBG_F64) return fail(c, SDBG_EINVAL, "sum_int_field is a float column");
      P.has_sum_i = 1;
    }
    if (avg_f64_field != UINT64_MAX) {
      if ((rc = col_view(s, avg_f64_field, &P.sum_f, &r))) return rc;
      if (r != rows) return fail(c, SDBG_EINVAL, "avg_f64 column length differs");
      if (P.sum_f.type != SDBG_F64) return fail(c, SDBG_EINVAL, "avg_f64_field is not a float column");
      P.has_sum_f = 1;
    }
    P.wide_int = plan.wide_int; P.count_f = plan.count_f;
    P.key_min = key_min; P.key_span = span; P.rows = rows;
    P.table = table; P.cnt_f = cnt_f; P.out_of_range = oor;
    // NOT NULL columns (the common analytic case) go through the TMA-pipelined kernel; nullable
    // columns need their validity words next to the values and keep the register-staged kernel.
    bool any_nullable = P.key.validity || (P.has_sum_i && P.sum_i.validity) || (P.has_sum_f && P.sum_f.validity);
    for (int i = 0; i < P.ps.n; ++i) any_nullable |= P.ps.p[i].col.validity != nullptr || P.ps.p[i].op >= 7;
    if (plan.pack_tables && any_nullable) return fail(c, SDBG_EINVAL, "internal: packed accumulators planned for a nullable segment");
    if (!any_nullable && env_int("SDBG_GROUPBY_TMA", 1)) {
      TmaGroupByParams T;
      std::memset(&T, 0, sizeof T);
      auto stream_of = [&](const ColDev& col) {
        for (int i = 0; i < T.n_streams; ++i) if (T.src[i] == col.values) return i;
        T.src[T.n_streams] = col.values; T.elem[T.n_streams] = col.type == SDBG_I32 ? 4 : 8;
        return T.n_streams++;
      };
      // Every comparison becomes a closed range [lo, lo + span] in an int64 key space (integers as they
      // are, doubles through fkey()): exact, because integers step by 1 and doubles by one ulp. Predicates
      // that hold for every row are dropped; one that holds for none makes the segment contribute nothing.
      bool never = false;
      int stream_idx[kMaxPreds];
      T.n_preds = 0;
      for (int i = 0; i < P.ps.n; ++i) {
        const PredDev& pd = P.ps.p[i];
        int64_t lo, hi; int neg = 0; bool empty = false;
        if (pd.col.type == SDBG_F64) {
          double lf = -HUGE_VAL, hf = HUGE_VAL;
          const double x = pd.lo_f;
          if (std::isnan(x) || (pd.op == SDBG_OP_BETWEEN && std::isnan(pd.hi_f))) empty = true;   // comparisons with NaN are false
          switch (pd.op) {
            case SDBG_OP_LT: if (x == -HUGE_VAL) empty = true; else hf = std::nextafter(x, -HUGE_VAL); break;
            case SDBG_OP_LE: hf = x; break;
            case SDBG_OP_GT: if (x == HUGE_VAL) empty = true; else lf = std::nextafter(x, HUGE_VAL); break;
            case SDBG_OP_GE: lf = x; break;
            case SDBG_OP_EQ: lf = hf = x; break;
            case SDBG_OP_NE: lf = hf = x; neg = 1; break;
            default: lf = x; hf = pd.hi_f; break;
          }
          if (pd.op == SDBG_OP_NE && std::isnan(x)) continue;      // v <> NaN holds for every row
          if (!empty && lf > hf) empty = true;
          if (!empty) {
            if (lf == 0.0) lf = -0.0;                               // -0.0 == +0.0: the range must cover both keys
            if (hf == 0.0) hf = 0.0;
            int64_t bl, bh; std::memcpy(&bl, &lf, 8); std::memcpy(&bh, &hf, 8);
            lo = fkey(bl); hi = fkey(bh);
          }
        } else {
          int64_t li = INT64_MIN, hi_ = INT64_MAX;
          const int64_t x = pd.lo_i;
          switch (pd.op) {
            case SDBG_OP_LT: if (x == INT64_MIN) empty = true; else hi_ = x - 1; break;
            case SDBG_OP_LE: hi_ = x; break;
            case SDBG_OP_GT: if (x == INT64_MAX) empty = true; else li = x + 1; break;
            case SDBG_OP_GE: li = x; break;
            case SDBG_OP_EQ: li = hi_ = x; break;
            case SDBG_OP_NE: li = hi_ = x; neg = 1; break;
            default: li = x; hi_ = pd.hi_i; break;
          }
          if (li > hi_) empty = true;
          lo = li; hi = hi_;
        }
        if (empty) { if (neg) continue; never = true; break; }
        const int k = T.n_preds++;
        stream_idx[k] = stream_of(pd.col);
        T.pred_type[k] = pd.col.type; T.pred_negate[k] = neg;
        T.pred_lo[k] = lo; T.pred_span[k] = static_cast<uint64_t>(hi) - static_cast<uint64_t>(lo);
      }
      if (never) continue;   // WHERE is false for every row of this segment
      // Zonemap verdicts: blocks whose min / max miss a predicate's range are skipped by producer and consumers alike.
      T.skip = nullptr;
      if (T.n_preds && env_int("SDBG_ZONEMAP", 1) != 0) {
        ZoneVerdictParams Z;
        std::memset(&Z, 0, sizeof Z);
        Z.n_preds = T.n_preds;
        Z.n_blocks = (rows + kZoneRows - 1) / kZoneRows;
        bool any_zone = false;
        for (int k2 = 0; k2 < T.n_preds; ++k2) {
          Z.lo[k2] = T.pred_lo[k2]; Z.span[k2] = T.pred_span[k2]; Z.negate[k2] = T.pred_negate[k2];
          ColumnObj* co = nullptr;
          for (auto& kv : s->cols) if (kv.second.d_values == T.src[stream_idx[k2]]) { co = &kv.second; break; }
          if (!co || co->d_validity) continue;
          if (!co->d_zone) {
            CU(c, cudaMalloc(reinterpret_cast<void**>(&co->d_zone), Z.n_blocks * 16));
            const unsigned zg = unsigned(std::min<uint64_t>((Z.n_blocks + 7) / 8, uint64_t(c->sm_count) * 8));
            const auto* vals = static_cast<const unsigned char*>(co->d_values);
            if (co->type == SDBG_F64) zonemap_kernel<1><<<zg, 256, 0, c->stream>>>(vals, rows, co->d_zone);
            else if (co->type == SDBG_I32) zonemap_kernel<2><<<zg, 256, 0, c->stream>>>(vals, rows, co->d_zone);
            else zonemap_kernel<0><<<zg, 256, 0, c->stream>>>(vals, rows, co->d_zone);
            ++c->launches;
          }
          Z.zone[k2] = co->d_zone;
          any_zone = true;
        }
        if (any_zone) {
          DevBuf& b_skip = c->scratch[13];
          if ((rc = ensure(c, b_skip, Z.n_blocks + 16))) return rc;
          if (!c->d_zone_skipped) CU(c, cudaMalloc(reinterpret_cast<void**>(&c->d_zone_skipped), 8));
          if (si == 0) { CU(c, cudaMemsetAsync(c->d_zone_skipped, 0, 8, c->stream)); c->zone_blocks_total = 0; }
          const unsigned vg = unsigned(std::min<uint64_t>((Z.n_blocks + 255) / 256, uint64_t(c->sm_count) * 4));
          zone_verdict_kernel<<<vg, 256, 0, c->stream>>>(Z, static_cast<uint8_t*>(b_skip.p), c->d_zone_skipped);
          ++c->launches;
          c->zone_blocks_total += Z.n_blocks;
          T.skip = static_cast<const uint8_t*>(b_skip.p);
        }
      }
      const int key_s = stream_of(P.key);
      const int sum_i_s = P.has_sum_i ? stream_of(P.sum_i) : -1;
      const int sum_f_s = P.has_sum_f ? stream_of(P.sum_f) : -1;
      T.key_type = P.key.type; T.sum_i_type = P.has_sum_i ? P.sum_i.type : 0;
      T.has_sum_i = P.has_sum_i; T.has_sum_f = P.has_sum_f;
      T.debug_skip = env_int("SDBG_GROUPBY_DEBUG", 0);
      T.wide_int = plan.wide_int; T.key_min = key_min; T.key_span = span; T.rows = rows; T.table = table; T.out_of_range = oor;
      T.pack_shift = plan.pack_shift; T.pack_tables = plan.pack_tables; T.pack_bias = plan.pack_bias;
      T.fix_limb = plan.fix_limb; T.fix_eunit = plan.fix_eunit;
      auto set_offsets = [&](int tile_rows) {   // stream offsets inside a stage depend on the tile shape
        uint32_t o = 0;
        for (int i = 0; i < T.n_streams; ++i) { T.off[i] = o; o += uint32_t(T.elem[i]) * uint32_t(tile_rows); }
        T.off[T.n_streams] = o;
        for (int i = 0; i < T.n_preds; ++i) T.pred_off[i] = T.off[stream_idx[i]];
        T.key_off = T.off[key_s];
        T.sum_i_off = sum_i_s >= 0 ? T.off[sum_i_s] : 0;
        T.sum_f_off = sum_f_s >= 0 ? T.off[sum_f_s] : 0;
      };
      const int shape = env_int("SDBG_GROUPBY_TMA_SHAPE", 0);
      const bool quad = plan.quad != 0;
      auto launch = [&](auto kern, int stages, int tile_rows, int consumer_warps) -> int {
        set_offsets(tile_rows);
        const size_t smem = size_t(stages) * size_t(T.off[T.n_streams]) + (quad ? size_t(consumer_warps) * (2048 + 256) : 0);
        CU(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        const size_t fit = std::max<size_t>(1, (220 * 1024) / (smem + 2048));
        const size_t by_threads = std::max<size_t>(1, 2048 / (size_t(consumer_warps + 1) * 32));
        const unsigned per_sm = unsigned(std::min<size_t>(std::min(fit, by_threads), size_t(env_int("SDBG_GROUPBY_TMA_CTAS", 8))));
        const unsigned grid = unsigned(c->sm_count) * per_sm;
        ProfScope ps_(c, kProfGroupBy);
        kern<<<grid, (consumer_warps + 1) * 32, smem, c->stream>>>(T);
        return SDBG_OK;
      };
      int lrc;
      switch (shape) {
        case 1: lrc = launch(filter_groupby_tma_kernel<4, 256, 8, false, false>, 4, 256, 8); break;
        case 2: lrc = launch(filter_groupby_tma_kernel<4, 512, 16, false, false>, 4, 512, 16); break;
        case 3: lrc = launch(filter_groupby_tma_kernel<3, 256, 8, false, false>, 3, 256, 8); break;
        case 4: lrc = launch(filter_groupby_tma_kernel<4, 1024, 16, false, false>, 4, 1024, 16); break;
        default: {
          // default shape: 512-row tiles, 8 consumer warps; ring depth by SDBG_GROUPBY_TMA_STAGES (the CTAs per
          // SM follow from the shared-memory footprint, so fewer stages = more resident consumer warps)
          const int stages = env_int("SDBG_GROUPBY_TMA_STAGES", kGroupByDefaultStages);
#define SDBG_GB_LAUNCH(ST) \
          (plan.pack_tables ? (quad ? launch(filter_groupby_tma_kernel<ST, kGroupByTileRows, 8, true, true>, ST, kGroupByTileRows, 8) \
                                    : launch(filter_groupby_tma_kernel<ST, kGroupByTileRows, 8, true, false>, ST, kGroupByTileRows, 8)) \
                            : (quad ? launch(filter_groupby_tma_kernel<ST, kGroupByTileRows, 8, false, true>, ST, kGroupByTileRows, 8) \
                                    : launch(filter_groupby_tma_kernel<ST, kGroupByTileRows, 8, false, false>, ST, kGroupByTileRows, 8)))
          lrc = stages == 2 ? SDBG_GB_LAUNCH(2) : stages == 3 ? SDBG_GB_LAUNCH(3) : SDBG_GB_LAUNCH(4);
#undef SDBG_GB_LAUNCH
          break;
        }
      }
      if (lrc) return lrc;
    } else {
      const unsigned grid = unsigned(c->sm_count) * unsigned(env_int("SDBG_GROUPBY_CTAS_PER_SM", 8));
      ProfScope ps_(c, kProfGroupBy);
      filter_groupby_kernel<2><<<grid, 256, 0, c->stream>>>(P);
    }
    ++c->launches;
    CU(c, cudaGetLastError());
  }
  groupby_pack_kernel<<<c->sm_count * 2, 256, 0, c->stream>>>(table, plan.count_f ? cnt_f : nullptr, span,
                                                               static_cast<long long*>(d_i64), static_cast<double*>(d_f64),
                                                               plan.pack_shift, plan.pack_tables, plan.pack_bias, plan.fix_limb, plan.fix_eunit);
  ++c->launches;
  CU(c, cudaGetLastError());
  if (defer_check) {
    // the partial path stays asynchronous (a collective usually follows on the same stream): the out-of-range count
    // lands in pinned memory and is looked at by the next call that synchronises (finalize / sdbg_sync)
    if (!c->h_oor) CU(c, cudaHostAlloc(reinterpret_cast<void**>(&c->h_oor), 8, cudaHostAllocDefault));
    CU(c, cudaMemcpyAsync(c->h_oor, oor, 8, cudaMemcpyDeviceToHost, c->stream));
    c->oor_pending = true;
  } else {
    unsigned long long h_oor = 0;
    CU(c, cudaMemcpyAsync(&h_oor, oor, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    if (h_oor) return fail(c, SDBG_EINVAL, "GROUP BY key outside [key_min, key_min + span)");
  }
  if (plan_out) *plan_out = plan;
  return SDBG_OK;
}

}  // namespace

namespace {

int groupby_hash(sdbg_segment* const* segs, size_t n_segs, const sdbg_col_pred* preds, size_t n_preds, uint64_t key_field,
                 uint32_t n_groups_hint, uint64_t sum_int_field, uint64_t avg_f64_field, sdbg_group_row* out, uint64_t cap,
                 uint64_t* n_out) {
  sdbg_ctx* c = segs[0]->ctx;
  uint64_t capacity = 1 << 16;
  while (capacity < 2ull * std::max<uint64_t>(n_groups_hint, cap)) capacity <<= 1;
  int rc;
  for (int attempt = 0; attempt < 8; ++attempt, capacity <<= 2) {
    const size_t bytes = (capacity + 1) * sizeof(HashSlot);
    if (bytes > (size_t(8) << 30)) return fail(c, SDBG_ECAPACITY, "hash aggregate table would exceed 8 GiB");
    if ((rc = ensure(c, c->scratch[11], bytes + 64))) return rc;
    auto* table = static_cast<HashSlot*>(c->scratch[11].p);
    auto* overflow = reinterpret_cast<unsigned int*>(static_cast<char*>(c->scratch[11].p) + bytes);
    hash_init_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(table, capacity + 1);
    ++c->launches;
    CU(c, cudaMemsetAsync(overflow, 0, 64, c->stream));
    for (size_t si = 0; si < n_segs; ++si) {
      sdbg_segment* s = segs[si];
      HashGroupByParams P;
      std::memset(&P, 0, sizeof P);
      uint64_t rows = 0, r = 0;
      if ((rc = pred_set(s, preds, n_preds, &P.ps, &rows))) return rc;
      if ((rc = col_view(s, key_field, &P.key, &r))) return rc;
      if (P.key.validity) return fail(c, SDBG_EUNSUPPORTED, "nullable GROUP BY key");
      if (P.key.type == SDBG_F64) return fail(c, SDBG_EUNSUPPORTED, "float GROUP BY key");
      if (!rows) rows = r;
      if (r != rows) return fail(c, SDBG_EINVAL, "key column length differs");
      if (sum_int_field != UINT64_MAX) {
        if ((rc = col_view(s, sum_int_field, &P.sum_i, &r))) return rc;
        if (r != rows) return fail(c, SDBG_EINVAL, "sum_int column length differs");
        if (P.sum_i.type == SDBG_F64) return fail(c, SDBG_EINVAL, "sum_int_field is a float column");
        P.has_sum_i = 1;
      }
      if (avg_f64_field != UINT64_MAX) {
        if ((rc = col_view(s, avg_f64_field, &P.sum_f, &r))) return rc;
        if (r != rows) return fail(c, SDBG_EINVAL, "avg_f64 column length differs");
        if (P.sum_f.type != SDBG_F64) return fail(c, SDBG_EINVAL, "avg_f64_field is not a float column");
        P.has_sum_f = 1;
      }
      P.rows = rows; P.table = table; P.capacity = capacity; P.overflow = overflow;
      { ProfScope ps_(c, kProfGroupBy);
        filter_groupby_hash_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(P); }
      ++c->launches;
      CU(c, cudaGetLastError());
    }
    unsigned int h_over = 0;
    CU(c, cudaMemcpyAsync(&h_over, overflow, 4, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    if (h_over) continue;  // table filled up: retry with 4x the capacity
    std::vector<HashSlot> h(capacity + 1);
    CU(c, cudaMemcpyAsync(h.data(), table, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    std::vector<const HashSlot*> live;
    for (const HashSlot& sl : h) if (sl.count) live.push_back(&sl);
    std::sort(live.begin(), live.end(), [](const HashSlot* a, const HashSlot* b) { return a->key < b->key; });
    *n_out = live.size();
    if (live.size() > cap) return fail(c, SDBG_ECAPACITY, "group output buffer too small");
    for (size_t i = 0; i < live.size(); ++i) {
      const HashSlot& sl = *live[i];
      sdbg_group_row& g = out[i];
      g.key = sl.key; g.count = sl.count;
      const __int128 tot = (static_cast<__int128>(sl.sum_hi) << 32) + static_cast<__int128>(sl.sum_lo);
      g.sum_i128[0] = int64_t(uint64_t(static_cast<unsigned __int128>(tot)));
      g.sum_i128[1] = int64_t(uint64_t(static_cast<unsigned __int128>(tot) >> 64));
      g.sum_f64 = sl.sum_f;
      g.cnt_f64 = avg_f64_field != UINT64_MAX ? sl.cnt_f : sl.count;
    }
    return SDBG_OK;
  }
  return fail(c, SDBG_ECAPACITY, "hash aggregate: too many groups");
}

}  // namespace

extern "C" int sdbg_filter_groupby_partial(sdbg_segment* const* segs, size_t n_segs, const sdbg_col_pred* preds, size_t n_preds,
                                           uint64_t key_field, int64_t key_min, uint64_t key_span, uint64_t sum_int_field,
                                           uint64_t avg_f64_field, void* d_i64, void* d_f64) {
  if (!segs || !n_segs || !d_i64 || !d_f64 || !key_span || (!preds && n_preds)) return SDBG_EINVAL;
  CU(segs[0]->ctx, cudaSetDevice(segs[0]->ctx->device));
  GroupPlan plan;
  // asynchronous: nothing here waits for the GPU (SUM(int) leaves as two limbs, total = hi * 2^32 + lo; a narrow sum
  // keeps the whole value in lo -- sdbg_dist_groupby_merge normalises the limbs before they are all-reduced)
  return groupby_launch(segs, n_segs, preds, n_preds, key_field, key_min, key_span, sum_int_field, avg_f64_field, d_i64, d_f64, &plan, true);
}

extern "C" int sdbg_groupby_finalize(sdbg_ctx* c, int64_t key_min, uint64_t span, const void* d_i64, const void* d_f64,
                                     sdbg_group_row* out, uint64_t cap, uint64_t* n_out) {
  if (!c || !d_i64 || !d_f64 || !out || !n_out || !span) return SDBG_EINVAL;
  CU(c, cudaSetDevice(c->device));
  // The dense partials are exactly as large as the answer (40 B per key), so they cross PCIe once
  // into pinned memory and the ascending-key compaction of non-empty groups is a host loop -- one
  // stream synchronisation per call instead of a compaction kernel plus three.
  int rc;
  if ((rc = ensure_pinned(c, span * 40))) return rc;
  auto* h_i = static_cast<long long*>(c->h_pinned);
  auto* h_f = reinterpret_cast<double*>(h_i + 4 * span);
  CU(c, cudaMemcpyAsync(h_i, d_i64, span * 32, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(h_f, d_f64, span * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  if (c->oor_pending) {
    c->oor_pending = false;
    if (*c->h_oor) return fail(c, SDBG_EINVAL, "GROUP BY key outside [key_min, key_min + span)");
  }
  uint64_t n = 0;
  for (uint64_t i = 0; i < span; ++i) {
    if (h_i[i] == 0) continue;
    if (n < cap) {
      sdbg_group_row& r = out[n];
      r.key = key_min + int64_t(i);
      r.count = uint64_t(h_i[i]);
      // SUM(int) limbs: total = hi * 2^32 + lo (a narrow sum has hi == 0); exact in 128 bits.
      const __int128 tot = (static_cast<__int128>(h_i[2 * span + i]) << 32) + static_cast<__int128>(h_i[span + i]);
      r.sum_i128[0] = int64_t(uint64_t(static_cast<unsigned __int128>(tot)));
      r.sum_i128[1] = int64_t(uint64_t(static_cast<unsigned __int128>(tot) >> 64));
      r.sum_f64 = h_f[i];
      r.cnt_f64 = uint64_t(h_i[3 * span + i]);
    }
    ++n;
  }
  *n_out = n;
  if (n > cap) return fail(c, SDBG_ECAPACITY, "group output buffer too small");
  return SDBG_OK;
}

extern "C" int sdbg_filter_groupby(sdbg_segment* const* segs, size_t n_segs, const sdbg_col_pred* preds, size_t n_preds,
                                   uint64_t key_field, uint32_t n_groups_hint, uint64_t sum_int_field, uint64_t avg_f64_field,
                                   sdbg_group_row* out, uint64_t cap, uint64_t* n_out) {
  if (!segs || !n_segs || !out || !n_out || (!preds && n_preds)) return SDBG_EINVAL;
  sdbg_ctx* c = segs[0]->ctx;
  CU(c, cudaSetDevice(c->device));
  int64_t kmin = INT64_MAX, kmax = INT64_MIN;
  for (size_t si = 0; si < n_segs; ++si) {
    int64_t mn, mx;
    const int rc = column_minmax(segs[si], key_field, &mn, &mx);
    if (rc) return rc;
    kmin = std::min(kmin, mn); kmax = std::max(kmax, mx);
  }
  if (kmin > kmax) { *n_out = 0; return SDBG_OK; }
  const unsigned __int128 span128 = static_cast<unsigned __int128>(static_cast<__int128>(kmax) - kmin) + 1;
  // Dense ("perfect hash") table when statistics bound the key range to something table-sized,
  // otherwise a real hash table sized from the group-count hint.
  const unsigned __int128 dense_limit = std::max<unsigned __int128>(static_cast<unsigned __int128>(1) << 20,
                                                                    static_cast<unsigned __int128>(n_groups_hint) * 8);
  if (span128 > (static_cast<unsigned __int128>(1) << 26) || span128 > dense_limit || env_int("SDBG_GROUPBY_FORCE_HASH", 0))
    return groupby_hash(segs, n_segs, preds, n_preds, key_field, n_groups_hint, sum_int_field, avg_f64_field, out, cap, n_out);
  const uint64_t span = uint64_t(span128);
  int rc;
  if ((rc = ensure(c, c->scratch[8], span * 40 + 64))) return rc;
  void* d_i64 = c->scratch[8].p;
  void* d_f64 = static_cast<char*>(c->scratch[8].p) + span * 32;
  GroupPlan plan;
  if ((rc = groupby_launch(segs, n_segs, preds, n_preds, key_field, kmin, span, sum_int_field, avg_f64_field, d_i64, d_f64, &plan))) return rc;
  return sdbg_groupby_finalize(c, kmin, span, d_i64, d_f64, out, cap, n_out);
}

// ------------------------------------------------------------------------------------------
// Collectives over NVLink for a C++ host (no torch): NCCL resolved lazily with dlopen, so libsdbg.so itself does not
// link it and a process that already carries an NCCL (e.g. PyTorch's bundled one: same SONAME) shares that copy.
// Everything below is enqueued on the context's stream: a partial aggregate or a top-k batch flows kernel ->
// collective -> kernel without a host synchronisation in between.
// ------------------------------------------------------------------------------------------
namespace {
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok() const { return lib && GetUniqueId && CommInitRank && CommDestroy && AllReduce && AllGather && GetErrorString; }
};
NcclApi& nccl_api() {
  static NcclApi api = [] {
    NcclApi a;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      a.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (a.lib) break;
    }
    if (a.lib) {
      a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(a.lib, "ncclGetUniqueId"));
      a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(a.lib, "ncclCommInitRank"));
      a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(a.lib, "ncclCommDestroy"));
      a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(a.lib, "ncclAllReduce"));
      a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(a.lib, "ncclAllGather"));
      a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(a.lib, "ncclGetErrorString"));
    }
    return a;
  }();
  return api;
}
#define NC(c, call)                                                                          \
  do {                                                                                       \
    const ncclResult_t r_ = (call);                                                          \
    if (r_ != ncclSuccess) return fail((c), SDBG_ECUDA, std::string("NCCL: ") + nccl_api().GetErrorString(r_)); \
  } while (0)

// SUM(double) partials as fixed point: x / 2^eunit rounded to a 120-bit integer, two signed 60-bit limbs in int64 --
// sums of up to 8 ranks cannot overflow a limb, the integer all-reduce is exact and independent of the rank order,
// and the only rounding left is the final conversion back to double.
__global__ void __launch_bounds__(256)
dist_pack_kernel(const long long* __restrict__ part_i64, const double* __restrict__ part_f64, uint64_t span, int eunit,
                 long long* __restrict__ wire /* [6 * span] */) {
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < span; i += uint64_t(gridDim.x) * blockDim.x) {
    wire[i] = part_i64[i];
    // SUM(int) limbs normalised so that lo is in [0, 2^32): the all-reduce of up to 2^31 ranks' lo limbs cannot wrap
    const long long lo = part_i64[span + i], hi = part_i64[2 * span + i];
    const long long carry = lo >> 32;                         // arithmetic shift: floor(lo / 2^32)
    wire[span + i] = lo - (carry << 32);
    wire[2 * span + i] = hi + carry;
    wire[3 * span + i] = part_i64[3 * span + i];
    long long l0, l1;
    fix_limbs(part_f64[i], 60, eunit, l0, l1);
    wire[4 * span + i] = l0;
    wire[5 * span + i] = l1;
  }
}
__global__ void __launch_bounds__(256)
dist_unpack_kernel(const long long* __restrict__ wire, uint64_t span, int eunit, long long* __restrict__ part_i64,
                   double* __restrict__ part_f64) {
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < span; i += uint64_t(gridDim.x) * blockDim.x) {
    part_i64[i] = wire[i];
    part_i64[span + i] = wire[span + i];
    part_i64[2 * span + i] = wire[2 * span + i];
    part_i64[3 * span + i] = wire[3 * span + i];
    part_f64[i] = fix_total(wire[4 * span + i], wire[5 * span + i], 60, eunit);
  }
}
}  // namespace

extern "C" int sdbg_dist_unique_id(uint8_t* id128) {
  if (!id128) return SDBG_EINVAL;
  if (!nccl_api().ok()) return SDBG_EUNSUPPORTED;
  ncclUniqueId id;
  if (nccl_api().GetUniqueId(&id) != ncclSuccess) return SDBG_ECUDA;
  std::memcpy(id128, id.internal, NCCL_UNIQUE_ID_BYTES);
  return SDBG_OK;
}

extern "C" int sdbg_dist_init(sdbg_ctx* c, const uint8_t* id128, int rank, int world) {
  if (!c || !id128 || world < 1 || rank < 0 || rank >= world) return SDBG_EINVAL;
  if (!nccl_api().ok()) return fail(c, SDBG_EUNSUPPORTED, "libnccl.so.2 not found");
  if (c->nccl_comm) return fail(c, SDBG_EINVAL, "context already has a communicator");
  CU(c, cudaSetDevice(c->device));
  ncclUniqueId id;
  std::memcpy(id.internal, id128, NCCL_UNIQUE_ID_BYTES);
  ncclComm_t comm = nullptr;
  NC(c, nccl_api().CommInitRank(&comm, world, id, rank));
  c->nccl_comm = comm; c->dist_rank = rank; c->dist_world = world;
  return SDBG_OK;
}

extern "C" int sdbg_dist_destroy(sdbg_ctx* c) {
  if (!c) return SDBG_EINVAL;
  if (c->nccl_comm) { nccl_api().CommDestroy(static_cast<ncclComm_t>(c->nccl_comm)); c->nccl_comm = nullptr; }
  c->dist_rank = 0; c->dist_world = 1;
  return SDBG_OK;
}

extern "C" int sdbg_dist_allreduce_i64(sdbg_ctx* c, void* d_buf, size_t n) {
  if (!c || !d_buf) return SDBG_EINVAL;
  if (!c->nccl_comm) return c->dist_world == 1 ? SDBG_OK : fail(c, SDBG_EINVAL, "sdbg_dist_init has not run");
  NC(c, nccl_api().AllReduce(d_buf, d_buf, n, ncclInt64, ncclSum, static_cast<ncclComm_t>(c->nccl_comm), c->stream));
  return SDBG_OK;
}

extern "C" int sdbg_dist_allgather(sdbg_ctx* c, const void* d_send, void* d_recv, size_t bytes_per_rank) {
  if (!c || !d_send || !d_recv) return SDBG_EINVAL;
  if (!c->nccl_comm) {
    if (c->dist_world != 1) return fail(c, SDBG_EINVAL, "sdbg_dist_init has not run");
    CU(c, cudaMemcpyAsync(d_recv, d_send, bytes_per_rank, cudaMemcpyDeviceToDevice, c->stream));
    return SDBG_OK;
  }
  NC(c, nccl_api().AllGather(d_send, d_recv, bytes_per_rank, ncclInt8, static_cast<ncclComm_t>(c->nccl_comm), c->stream));
  return SDBG_OK;
}

// Dense GROUP BY partials of every rank -> the global partials on every rank, in ONE all-reduce: counts, the SUM(int)
// limbs and SUM(double) as fixed-point limbs travel in one int64 buffer. abs_bound >= |sum of the double column over
// all ranks' passing rows| fixes the fixed-point unit (identical on every rank: derive it from the column statistics
// and the total row count, which are known when the shards are built).
// Distributed top-k in one call: local scan of this rank's segments, ONE all-gather of every rank's k best keys per query
// over NVLink, local selection of the global top-k -- enqueued back to back on the context's stream, one host
// synchronisation at the very end (none when out == NULL: the keys stay in HBM, see sdbg_topk_merge_gathered).
extern "C" int sdbg_dist_bm25_topk_batch(sdbg_segment* const* segs, size_t n_segs, int kind, const sdbg_bm25_term* terms,
                                         const uint32_t* term_off, size_t nq, float k1, float b, const sdbg_col_pred* filt, uint32_t k,
                                         float threshold_in, sdbg_hit* out, uint32_t* n_out) {
  if (!segs || !n_segs) return SDBG_EINVAL;
  sdbg_ctx* c = segs[0]->ctx;
  const uint32_t world = uint32_t(c->dist_world), rank = uint32_t(c->dist_rank);
  if (world > 1 && !c->nccl_comm) return fail(c, SDBG_EINVAL, "sdbg_dist_init has not run");
  DevBuf& mine = c->scratch[6 + 6];  // scratch[12..]: see DevBuf scratch[] size
  DevBuf& all = c->scratch[6 + 7];
  int rc;
  const size_t bytes = nq * size_t(k) * 8;
  if ((rc = ensure(c, mine, bytes))) return rc;
  if ((rc = ensure(c, all, bytes * world))) return rc;
  if ((rc = topk_batch_device_impl(segs, n_segs, kind, terms, term_off, nq, k1, b, filt, k, threshold_in, rank, mine.p, nullptr, false))) return rc;
  if ((rc = sdbg_dist_allgather(c, mine.p, all.p, bytes))) return rc;
  if (!out) return sdbg_topk_merge_gathered(c, all.p, world, nq, k, nullptr, nullptr);
  return sdbg_topk_merge_gathered(c, all.p, world, nq, k, out, n_out);
}

extern "C" int sdbg_dist_groupby_merge(sdbg_ctx* c, void* d_i64, void* d_f64, uint64_t span, double abs_bound) {
  if (!c || !d_i64 || !d_f64 || !span || !(abs_bound >= 0.0)) return SDBG_EINVAL;
  if (!c->nccl_comm) return c->dist_world == 1 ? SDBG_OK : fail(c, SDBG_EINVAL, "sdbg_dist_init has not run");
  if (c->dist_world > 8) return fail(c, SDBG_EUNSUPPORTED, "fixed-point limbs are sized for up to 8 ranks");
  CU(c, cudaSetDevice(c->device));
  int ex = 0;
  std::frexp(abs_bound > 0.0 ? abs_bound : 1.0, &ex);        // abs_bound < 2^ex
  const int eunit = ex + 1 - 117;                              // 120-bit fixed point with 3 bits of headroom for 8 ranks
  DevBuf& wire = c->scratch[14];
  int rc;
  if ((rc = ensure(c, wire, span * 6 * sizeof(long long)))) return rc;
  const unsigned grid = unsigned(std::min<uint64_t>((span + 255) / 256, uint64_t(c->sm_count) * 8));
  dist_pack_kernel<<<grid, 256, 0, c->stream>>>(static_cast<const long long*>(d_i64), static_cast<const double*>(d_f64), span, eunit,
                                               static_cast<long long*>(wire.p));
  ++c->launches;
  NC(c, nccl_api().AllReduce(wire.p, wire.p, span * 6, ncclInt64, ncclSum, static_cast<ncclComm_t>(c->nccl_comm), c->stream));
  dist_unpack_kernel<<<grid, 256, 0, c->stream>>>(static_cast<const long long*>(wire.p), span, eunit, static_cast<long long*>(d_i64),
                                                 static_cast<double*>(d_f64));
  ++c->launches;
  CU(c, cudaGetLastError());
  return SDBG_OK;
}

// ------------------------------------------------------------------------------------------
// host-side writer mirror + synthetic inputs
// ------------------------------------------------------------------------------------------
extern "C" int sdbg_writer_create(uint32_t segment_docs, int has_wand, float wand_b, const uint32_t* norms, sdbg_writer** out) {
  if (!out) return SDBG_EINVAL;
  auto* w = new sdbg_writer;
  w->w.reset(new PostingWriter(segment_docs, has_wand != 0, wand_b, norms));
  *out = w;
  return SDBG_OK;
}
extern "C" void sdbg_writer_destroy(sdbg_writer* w) { delete w; }
extern "C" int sdbg_writer_add_term(sdbg_writer* w, const uint32_t* docs, const uint32_t* freqs, uint32_t n) {
  if (!w || (n && (!docs || !freqs))) return SDBG_EINVAL;
  for (uint32_t i = 1; i < n; ++i) if (docs[i] <= docs[i - 1]) return SDBG_EINVAL;
  if (n && docs[0] == 0) return SDBG_EINVAL;
  w->w->add_term(docs, freqs, n);
  return SDBG_OK;
}
extern "C" int sdbg_writer_finish(sdbg_writer* w, const uint8_t** doc_file, size_t* n, const sdbg_term_meta** terms, size_t* n_terms) {
  if (!w) return SDBG_EINVAL;
  w->metas.clear();
  for (const TermMeta& m : w->w->terms()) w->metas.push_back(sdbg_term_meta{m.docs_count, m.freq, m.doc_start, m.e_skip_start});
  if (doc_file) *doc_file = w->w->bytes().data();
  if (n) *n = w->w->bytes().size();
  if (terms) *terms = w->metas.data();
  if (n_terms) *n_terms = w->metas.size();
  return SDBG_OK;
}

extern "C" uint64_t sdbg_synth_hash(uint64_t stream, uint64_t index) { return synth_hash(stream, index); }

extern "C" int sdbg_synth_corpus_ex(sdbg_segment* seg, uint64_t doc0, uint32_t n_docs, uint32_t t0, uint32_t nt, int threads,
                                    double p_floor, uint32_t* docs_count_out, uint64_t* sum_dl_out);
extern "C" int sdbg_synth_corpus(sdbg_segment* seg, uint64_t doc0, uint32_t n_docs, uint32_t t0, uint32_t nt, int threads,
                                 uint32_t* docs_count_out, uint64_t* sum_dl_out) {
  return sdbg_synth_corpus_ex(seg, doc0, n_docs, t0, nt, threads, 0.0, docs_count_out, sum_dl_out);
}
// p_floor > 0: inclusion probability max(p_floor, min(0.5, 0.6 / (t + 1))) -- a flat tail of equally sized lists, used to
// build an index much larger than L2 (the HBM-resident bench workload).
extern "C" int sdbg_synth_corpus_ex(sdbg_segment* seg, uint64_t doc0, uint32_t n_docs, uint32_t t0, uint32_t nt, int threads,
                                    double p_floor, uint32_t* docs_count_out, uint64_t* sum_dl_out) {
  if (!seg || !n_docs || !nt || n_docs != seg->n_docs || !(p_floor >= 0.0 && p_floor <= 0.5)) return SDBG_EINVAL;
  threads = std::max(1, threads);
  std::vector<uint32_t> dl(n_docs);
  uint64_t sum_dl = 0;
  for (uint32_t i = 0; i < n_docs; ++i) { dl[i] = 16 + uint32_t(synth_hash(1, doc0 + 1 + i) % 240); sum_dl += dl[i]; }
  // one writer per term (terms are independent streams), built by a pool of threads, then
  // concatenated in term order -- byte-identical to a single sequential writer.
  std::vector<std::unique_ptr<PostingWriter>> per_term(nt);
  std::atomic<uint32_t> next{0};
  const float avg = float(double(sum_dl) / double(n_docs));
  auto work = [&]() {
    std::vector<uint32_t> docs, freqs;
    for (;;) {
      const uint32_t i = next.fetch_add(1);
      if (i >= nt) break;
      const uint32_t t = t0 + i;
      // terms 1000000 .. 1000004: BASELINE configs[3]'s conjunction terms, p = 0.50, 0.40, 0.30, 0.25, 0.20
      static const double kCfg4P[5] = {0.50, 0.40, 0.30, 0.25, 0.20};
      const double p = (t >= 1000000u && t < 1000005u) ? kCfg4P[t - 1000000u] : std::max(p_floor, std::min(0.5, 0.6 / double(t + 1)));
      const uint64_t thr = uint64_t(std::ldexp(p, 64));
      docs.clear(); freqs.clear();
      for (uint32_t d = 0; d < n_docs; ++d) {
        const uint64_t g = doc0 + 1 + d;
        if (synth_hash(100 + t, g) >= thr) continue;
        const uint64_t h2 = synth_hash(1000 + t, g);
        uint32_t f = 1 + (h2 ? uint32_t(__builtin_ctzll(h2)) : 64);
        f = std::min(f, dl[d]);
        docs.push_back(d + 1); freqs.push_back(f);
      }
      per_term[i].reset(new PostingWriter(n_docs, true, 0.75f, dl.data(), avg));
      per_term[i]->add_term(docs.data(), freqs.data(), uint32_t(docs.size()));
    }
  };
  std::vector<std::thread> pool;
  for (int i = 0; i < threads; ++i) pool.emplace_back(work);
  for (auto& th : pool) th.join();
  PostingWriter all(n_docs, true, 0.75f, dl.data(), avg);
  for (uint32_t i = 0; i < nt; ++i) {
    all.append(*per_term[i]);
    if (docs_count_out) docs_count_out[i] = per_term[i]->terms()[0].docs_count;
    per_term[i].reset();
  }
  if (sum_dl_out) *sum_dl_out = sum_dl;
  int rc = sdbg_stage_postings(seg, all.bytes().data(), all.bytes().size(),
                               reinterpret_cast<const sdbg_term_meta*>(all.terms().data()), all.terms().size(), 1);
  if (rc) return rc;
  std::vector<uint8_t> nb(n_docs);
  for (uint32_t i = 0; i < n_docs; ++i) nb[i] = uint8_t(dl[i]);
  const sdbg_norm_rg rg{1, n_docs, 0};
  return sdbg_stage_norms(seg, nb.data(), nb.size(), &rg, 1);
}

extern "C" int sdbg_synth_column(sdbg_segment* seg, uint64_t field, uint64_t stream, int kind, uint64_t row0, uint64_t rows) {
  if (!seg || !rows) return SDBG_EINVAL;
  sdbg_ctx* c = seg->ctx;
  CU(c, cudaSetDevice(c->device));
  const int type = (kind == 2 || kind == 4) ? SDBG_F64 : (kind == 6 ? SDBG_I32 : SDBG_I64);
  ColumnObj& col = seg->cols[field];
  free_column(col);
  col = ColumnObj{};
  const size_t bytes = rows * type_width(type);
  CU(c, cudaMalloc(&col.d_values, bytes + 64));
  CU(c, cudaMemsetAsync(static_cast<char*>(col.d_values) + bytes, 0, 64, c->stream));
  col.type = type; col.rows = rows; col.owned = true;
  synth_column_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(stream, kind, row0, rows, col.d_values);
  ++c->launches;
  CU(c, cudaGetLastError());
  return SDBG_OK;
}

// ------------------------------------------------------------------------------------------
// Host-only probe of the staging parser (no device needed): lets CPU tests compare the block table
// built from a ".doc" stream with the oracle's reading of the same bytes.
// ------------------------------------------------------------------------------------------
extern "C" int sdbg_debug_stage_host(const uint8_t* doc_file, size_t n, const sdbg_term_meta* terms, size_t n_terms, int has_wand,
                                     uint32_t cap, uint32_t* n_blocks, uint32_t* term_blk_begin /* n_terms+1 */,
                                     uint32_t* last_doc, uint32_t* prev_last, uint32_t* packed, uint32_t* max_freq,
                                     uint32_t* max_norm, uint64_t* arena_bytes) {
  StagedPostings sp;
  const std::string e = stage_postings(doc_file, n, reinterpret_cast<const TermMeta*>(terms), n_terms, has_wand != 0, &sp);
  if (!e.empty()) { std::fprintf(stderr, "sdbg_debug_stage_host: %s\n", e.c_str()); return SDBG_EFORMAT; }
  if (n_blocks) *n_blocks = uint32_t(sp.blocks.size());
  if (arena_bytes) *arena_bytes = sp.arena.size();
  if (term_blk_begin) std::copy(sp.term_blk_begin.begin(), sp.term_blk_begin.end(), term_blk_begin);
  if (sp.blocks.size() > cap) return SDBG_ECAPACITY;
  for (size_t i = 0; i < sp.blocks.size(); ++i) {
    if (last_doc) last_doc[i] = sp.blocks[i].last_doc;
    if (prev_last) prev_last[i] = sp.blocks[i].prev_last;
    if (packed) packed[i] = sp.blocks[i].packed;
    if (max_freq) max_freq[i] = sp.blk_max[i].freq;
    if (max_norm) max_norm[i] = sp.blk_max[i].norm;
  }
  return SDBG_OK;
}
