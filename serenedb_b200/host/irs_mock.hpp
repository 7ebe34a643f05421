// irs_mock.hpp -- MOCK of the reference declarations the adapters are written against.
//
// The real headers need Abseil, DuckDB and clang-21/C++26, none of which exist in this build
// environment, so the adapters compile against these minimal stand-ins. Each declaration copies the
// SIGNATURE (names, argument order and meaning) of the cited reference declaration and nothing else;
// inside the real tree this header is replaced by the includes named next to each block.
#pragma once

#include <cstddef>
#include <cstdint>
#include <limits>
#include <span>
#include <vector>

namespace irs {

using doc_id_t = uint32_t;  // irs/types.hpp
using score_t = float;
namespace doc_limits {      // irs/utils/type_limits.hpp:39-51
constexpr doc_id_t eof() noexcept { return std::numeric_limits<doc_id_t>::max(); }
constexpr doc_id_t invalid() noexcept { return 0; }
constexpr doc_id_t(min)() noexcept { return 1; }
}  // namespace doc_limits

struct ScoreFunction {};        // irs/search/score_function.hpp:77 (opaque here: bulk scoring stays on the GPU)
class ColumnArgsFetcher {};     // irs/search/column_collector.hpp:30
struct PrepareScoreContext {};  // irs/index/iterators.hpp:49-53
struct FillBlockScoreContext {};
struct FillBlockMatchContext {};

// irs/search/scorer.hpp:49-55
struct ScoreThresholdAttr { score_t value = std::numeric_limits<score_t>::lowest(); };
// irs/search/cost.hpp
struct CostAttr { uint64_t value = 0; };

// irs/index/iterators.hpp:67-91
class ScoreCollector {
 public:
  virtual void Add(score_t score, doc_id_t doc) = 0;
  virtual void AddWindow(const score_t* scores, const uint64_t* mask, doc_id_t min, size_t num_blocks, bool clear_score) = 0;
  virtual void AddDocs(const doc_id_t* docs, size_t count, const score_t* scores) = 0;
 protected:
  ~ScoreCollector() = default;
};

// irs/index/iterators.hpp:93-101
struct ScoreDoc {
  score_t score = 0.0f;
  doc_id_t doc = doc_limits::eof();
  uint32_t segment_idx = 0;
};

// irs/index/iterators.hpp:279-356 (the members the adapters override)
struct DocIterator {
  virtual ~DocIterator() = default;
  const doc_id_t& value() const noexcept { return _doc; }
  virtual doc_id_t advance() = 0;
  virtual doc_id_t seek(doc_id_t target) = 0;
  virtual doc_id_t LazySeek(doc_id_t target) { return seek(target); }
  virtual void Collect(const ScoreFunction& scorer, ColumnArgsFetcher& fetcher, ScoreCollector& collector) = 0;
  virtual ScoreFunction PrepareScore(const PrepareScoreContext&) { return {}; }
  virtual uint32_t count() = 0;
  virtual uint32_t EmitDocs(doc_id_t* out, doc_id_t min, doc_id_t max) = 0;
  virtual uint32_t EmitScoredDocs(doc_id_t* out, score_t* scores, doc_id_t max, const ScoreFunction& scorer,
                                  ColumnArgsFetcher* fetcher, doc_id_t min) = 0;
 protected:
  mutable doc_id_t _doc = doc_limits::invalid();
};

}  // namespace irs

namespace duckdb {  // third_party/duckdb (absent): only what the scan adapter touches
constexpr uint64_t STANDARD_VECTOR_SIZE = 2048;
struct DataChunkMock {          // stands in for duckdb::DataChunk with flat vectors
  std::vector<int64_t> key;     // k
  std::vector<int64_t> count;   // COUNT(*)
  std::vector<int64_t> sum_lo;  // SUM(v) as HUGEINT: lower / upper
  std::vector<int64_t> sum_hi;
  std::vector<double> avg;      // AVG(w)
  uint64_t size = 0;
  void Reset() { key.clear(); count.clear(); sum_lo.clear(); sum_hi.clear(); avg.clear(); size = 0; }
};
}  // namespace duckdb
