// gpu_adapters.hpp -- the two C++ adapters a SereneDB maintainer would add:
//
//  * GpuTopKIterator : irs::DocIterator -- returned from PostingsReaderImpl::WandIterator
//    (irs/formats/posting/reader.hpp:457-501) instead of MaxScoreIterator / SingleWandIterator when a
//    segment is staged on a GPU. Collect() runs the fused scan+score+top-k kernel for the segment and
//    feeds <= k (doc, score) pairs to the caller's ScoreCollector, honouring / publishing the
//    ScoreThresholdAttr exactly as CollectSegmentTopK expects
//    (server/connector/duckdb_search_full_scan.cpp:1898-1920).
//  * GpuAggScan -- the body of a new ScanMode::GpuAgg branch of IResearchScanFunction
//    (server/connector/duckdb_search_full_scan.cpp:1645-1709, modes :56-77 of the .hpp): the pushed
//    TableFilterSet + GROUP BY + SUM/AVG/COUNT run in one kernel and the scan emits already-aggregated
//    rows, <= STANDARD_VECTOR_SIZE per call, cardinality 0 = end of scan (like RunCountScan :2201-2239).
//
// Errors: the ABI returns codes; the adapters turn them into C++ exceptions like the reference's
// IoError / THROW_SQL_ERROR call sites.
#pragma once

#include <atomic>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/sdbg.h"
#include "irs_mock.hpp"

namespace sdbg_host {

struct GpuError : std::runtime_error {
  int code;
  GpuError(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

// One scored query over one staged segment.
class GpuTopKIterator final : public irs::DocIterator {
 public:
  GpuTopKIterator(sdbg_segment* segment, int kind /* SDBG_QUERY_OR | SDBG_QUERY_AND */,
                  std::vector<sdbg_bm25_term> terms /* BM25Stats per term + boost */, float k1 /* BM25::k() */,
                  float b /* BM25::b() */, uint32_t k /* 0 = streaming mode: every match, see EmitScoredDocs */,
                  const sdbg_col_pred* table_filter /* nullable: the ColFilter wrap */);

  // Scored top-k: the hot path.
  void Collect(const irs::ScoreFunction&, irs::ColumnArgsFetcher&, irs::ScoreCollector& collector) override;
  // Windowed variant used by TableFilterDocIterator / streaming callers (RunStreamingScan,
  // duckdb_search_full_scan.cpp:2370-2403): hits with doc in [min, max). With k = 0 the iterator holds every match
  // of the segment (sdbg_bm25_scan), so draining it window by window is the reference's streaming scan.
  uint32_t EmitScoredDocs(irs::doc_id_t* out, irs::score_t* scores, irs::doc_id_t max, const irs::ScoreFunction&,
                          irs::ColumnArgsFetcher*, irs::doc_id_t min) override;
  uint32_t EmitDocs(irs::doc_id_t* out, irs::doc_id_t min, irs::doc_id_t max) override;
  uint32_t count() override;  // total matches (exhaustive, like DocIterator::count)
  irs::doc_id_t advance() override;
  irs::doc_id_t seek(irs::doc_id_t target) override;
  // Bitmap + score window of [min, max) for callers that merge iterators (Conjunction / MaxScore windows,
  // iterators.hpp:322-337): bit (doc - min) per hit, score accumulated into score.score_window per merge_type,
  // match counts when match.matches is given. The hits are this iterator's top-k (its whole result set).
  std::pair<irs::doc_id_t, bool> FillBlock(irs::doc_id_t min, irs::doc_id_t max, uint64_t* mask, irs::FillBlockScoreContext score,
                                           irs::FillBlockMatchContext match) override;
  void FetchScoreArgs(uint16_t) override {}   // score arguments never leave the GPU
  // AttributeProvider: ScoreThresholdAttr (the caller seeds / reads the running threshold through it,
  // doc_collector.hpp:124-130, duckdb_search_full_scan.cpp:1910-1914) and CostAttr (conjunction ordering).
  irs::Attribute* GetMutable(irs::TypeInfo::type_id type) noexcept override;

  irs::ScoreThresholdAttr& threshold() noexcept { return threshold_; }
  const irs::CostAttr& cost() const noexcept { return cost_; }
  uint64_t total_matches() const noexcept { return total_; }

 private:
  void run();
  sdbg_segment* seg_;
  int kind_;
  std::vector<sdbg_bm25_term> terms_;
  float k1_, b_;
  uint32_t k_;
  bool has_filter_;
  sdbg_col_pred filter_{};
  irs::ScoreThresholdAttr threshold_;
  irs::CostAttr cost_;
  std::vector<sdbg_hit> hits_;   // sorted by (score desc, doc asc) after run()
  std::vector<sdbg_hit> by_doc_; // same hits ordered by doc for advance()/seek()/Emit*
  size_t pos_ = 0;
  uint64_t total_ = 0;
  bool ran_ = false;
};

// SELECT key, COUNT(*), SUM(sum_int), AVG(avg_f64) FROM t WHERE preds GROUP BY key -- emitted in chunks.
class GpuAggScan {
 public:
  GpuAggScan(std::vector<sdbg_segment*> segments, std::vector<sdbg_col_pred> pushed_filters, uint64_t key_field,
             uint64_t sum_int_field, uint64_t avg_f64_field, uint32_t n_groups_hint);
  // IResearchScanFunction body for ScanMode::GpuAgg: fills `output`; output.size == 0 => exhausted.
  void Scan(duckdb::DataChunkMock& output);
  uint64_t rows_scanned() const noexcept { return rows_scanned_; }  // get_metrics hook (:860-864)

 private:
  std::vector<sdbg_segment*> segs_;
  std::vector<sdbg_col_pred> preds_;
  uint64_t key_, sum_i_, avg_f_;
  uint32_t hint_;
  std::vector<sdbg_group_row> groups_;
  size_t cursor_ = 0;
  bool ran_ = false;
  uint64_t rows_scanned_ = 0;
};

// The same scan mode under DuckDB's threading contract (duckdb_search_full_scan.hpp:85-255, .cpp:99-268): ONE global
// state shared by all workers of the query -- touched through atomics only, like next_segment / next_unit there -- and
// one local state per worker. The first worker to arrive runs the aggregation on the GPU (the others wait on the
// once-flag, as they would wait for rows anyway); afterwards every worker claims chunks of <= STANDARD_VECTOR_SIZE
// result rows from an atomic cursor and fills its own DataChunk. Cardinality 0 = this worker is done.
struct GpuAggGlobalState {
  GpuAggGlobalState(std::vector<sdbg_segment*> segments, std::vector<sdbg_col_pred> pushed_filters, uint64_t key_field,
                    uint64_t sum_int_field, uint64_t avg_f64_field, uint32_t n_groups_hint);
  std::vector<sdbg_segment*> segs;
  std::vector<sdbg_col_pred> preds;
  uint64_t key, sum_i, avg_f;
  uint32_t hint;
  std::once_flag ran;
  std::vector<sdbg_group_row> groups;          // written once (under `ran`), read-only afterwards
  std::atomic<size_t> next_chunk{0};           // claim cursor, in units of STANDARD_VECTOR_SIZE rows
  std::atomic<uint64_t> rows_emitted{0};       // get_metrics hook
};
struct GpuAggLocalState {
  uint64_t chunks_claimed = 0;
};
// Body of IResearchScanFunction for ScanMode::GpuAgg (:1645-1709): safe to call from many threads at once.
void GpuAggScanFunction(GpuAggGlobalState& g, GpuAggLocalState& l, duckdb::DataChunkMock& output);

}  // namespace sdbg_host
