#include "gpu_adapters.hpp"

#include <algorithm>
#include <cfloat>

namespace sdbg_host {

namespace {
void check(int rc, const char* what) {
  if (rc != SDBG_OK) throw GpuError(rc, std::string(what) + " failed with code " + std::to_string(rc));
}
}  // namespace

GpuTopKIterator::GpuTopKIterator(sdbg_segment* segment, int kind, std::vector<sdbg_bm25_term> terms, float k1, float b,
                                 uint32_t k, const sdbg_col_pred* table_filter)
    : seg_(segment), kind_(kind), terms_(std::move(terms)), k1_(k1), b_(b), k_(k), has_filter_(table_filter != nullptr) {
  if (table_filter) filter_ = *table_filter;
  threshold_.value = FLT_MIN;  // doc_collector.hpp:102
}

void GpuTopKIterator::run() {
  if (ran_) return;
  if (k_ == 0) {                   // streaming mode: every match, already in doc order
    uint64_t n = 0, cap = 0;
    std::vector<uint32_t> docs;
    std::vector<float> scores;
    for (;;) {
      const int rc = sdbg_bm25_scan(seg_, kind_, terms_.data(), terms_.size(), k1_, b_, has_filter_ ? &filter_ : nullptr, 1, UINT32_MAX,
                                    docs.data(), scores.data(), cap, &n);
      if (rc == SDBG_ECAPACITY && n > cap) { cap = n; docs.resize(n); scores.resize(n); continue; }   // count-only call, then one with room
      check(rc, "sdbg_bm25_scan");
      break;
    }
    by_doc_.resize(n);
    for (uint64_t i = 0; i < n; ++i) by_doc_[i] = sdbg_hit{scores[i], docs[i], 0};
    total_ = n;
    cost_.reset(total_);
    ran_ = true;
    return;
  }
  hits_.assign(k_, sdbg_hit{});
  uint32_t n = 0;
  float thr_out = 0;
  sdbg_segment* segs[1] = {seg_};
  check(sdbg_bm25_topk(segs, 1, kind_, terms_.data(), terms_.size(), k1_, b_, has_filter_ ? &filter_ : nullptr, k_,
                       threshold_.value, hits_.data(), &n, &total_, &thr_out),
        "sdbg_bm25_topk");
  hits_.resize(n);
  by_doc_ = hits_;
  std::sort(by_doc_.begin(), by_doc_.end(), [](const sdbg_hit& a, const sdbg_hit& b) { return a.doc < b.doc; });
  if (n == k_ && thr_out > threshold_.value) threshold_.value = thr_out;  // raise the caller-visible threshold
  cost_.reset(total_);
  ran_ = true;
}

void GpuTopKIterator::Collect(const irs::ScoreFunction&, irs::ColumnArgsFetcher&, irs::ScoreCollector& collector) {
  run();
  if (k_ == 0) hits_ = by_doc_;    // a streaming iterator asked to Collect feeds everything it has
  if (hits_.empty()) { _doc = irs::doc_limits::eof(); return; }
  std::vector<irs::doc_id_t> docs(hits_.size());
  std::vector<irs::score_t> scores(hits_.size());
  for (size_t i = 0; i < hits_.size(); ++i) { docs[i] = hits_[i].doc; scores[i] = hits_[i].score; }
  collector.AddDocs(docs.data(), docs.size(), scores.data());  // iterators.hpp:176-207
  _doc = irs::doc_limits::eof();
}

uint32_t GpuTopKIterator::EmitScoredDocs(irs::doc_id_t* out, irs::score_t* scores, irs::doc_id_t max, const irs::ScoreFunction&,
                                         irs::ColumnArgsFetcher*, irs::doc_id_t min) {
  run();
  uint32_t n = 0;
  while (pos_ < by_doc_.size() && by_doc_[pos_].doc < min) ++pos_;
  while (pos_ < by_doc_.size() && by_doc_[pos_].doc < max) { out[n] = by_doc_[pos_].doc; scores[n] = by_doc_[pos_].score; ++n; ++pos_; }
  _doc = pos_ < by_doc_.size() ? by_doc_[pos_].doc : irs::doc_limits::eof();
  return n;
}

uint32_t GpuTopKIterator::EmitDocs(irs::doc_id_t* out, irs::doc_id_t min, irs::doc_id_t max) {
  run();
  uint32_t n = 0;
  while (pos_ < by_doc_.size() && by_doc_[pos_].doc < min) ++pos_;
  while (pos_ < by_doc_.size() && by_doc_[pos_].doc < max) out[n++] = by_doc_[pos_++].doc;
  _doc = pos_ < by_doc_.size() ? by_doc_[pos_].doc : irs::doc_limits::eof();
  return n;
}

uint32_t GpuTopKIterator::count() { run(); return uint32_t(total_); }

irs::Attribute* GpuTopKIterator::GetMutable(irs::TypeInfo::type_id type) noexcept {
  if (type == irs::Type<irs::ScoreThresholdAttr>::id()) return &threshold_;
  if (type == irs::Type<irs::CostAttr>::id()) return &cost_;
  return nullptr;
}

std::pair<irs::doc_id_t, bool> GpuTopKIterator::FillBlock(irs::doc_id_t min, irs::doc_id_t max, uint64_t* mask,
                                                          irs::FillBlockScoreContext score, irs::FillBlockMatchContext match) {
  run();
  bool empty = true;
  while (pos_ < by_doc_.size() && by_doc_[pos_].doc < min) ++pos_;
  for (; pos_ < by_doc_.size() && by_doc_[pos_].doc < max; ++pos_) {
    const uint32_t off = by_doc_[pos_].doc - min;
    bool set = true;
    if (match.matches) set = ++match.matches[off] >= match.min_match_count;   // TrackMatch: bit only when the threshold is met
    if (set) { mask[off >> 6] |= uint64_t(1) << (off & 63); empty = false; }
    if (score.score_window) {
      irs::score_t& w = score.score_window[off];
      switch (score.merge_type) {
        case irs::ScoreMergeType::Sum: w += by_doc_[pos_].score; break;
        case irs::ScoreMergeType::Max: w = std::max(w, by_doc_[pos_].score); break;
        default: w = by_doc_[pos_].score; break;
      }
    }
  }
  _doc = pos_ < by_doc_.size() ? by_doc_[pos_].doc : irs::doc_limits::eof();
  return {_doc, match.matches ? empty : false};
}

irs::doc_id_t GpuTopKIterator::advance() {
  run();
  if (_doc != irs::doc_limits::invalid() && pos_ < by_doc_.size() && by_doc_[pos_].doc == _doc) ++pos_;
  return _doc = pos_ < by_doc_.size() ? by_doc_[pos_].doc : irs::doc_limits::eof();
}

irs::doc_id_t GpuTopKIterator::seek(irs::doc_id_t target) {
  run();
  while (pos_ < by_doc_.size() && by_doc_[pos_].doc < target) ++pos_;
  return _doc = pos_ < by_doc_.size() ? by_doc_[pos_].doc : irs::doc_limits::eof();
}

GpuAggScan::GpuAggScan(std::vector<sdbg_segment*> segments, std::vector<sdbg_col_pred> pushed_filters, uint64_t key_field,
                       uint64_t sum_int_field, uint64_t avg_f64_field, uint32_t n_groups_hint)
    : segs_(std::move(segments)), preds_(std::move(pushed_filters)), key_(key_field), sum_i_(sum_int_field),
      avg_f_(avg_f64_field), hint_(n_groups_hint) {}

void GpuAggScan::Scan(duckdb::DataChunkMock& output) {
  output.Reset();
  if (!ran_) {
    uint64_t cap = std::max<uint64_t>(hint_, 1024), n = 0;
    for (;;) {
      groups_.resize(cap);
      const int rc = sdbg_filter_groupby(segs_.data(), segs_.size(), preds_.data(), preds_.size(), key_, hint_, sum_i_, avg_f_,
                                         groups_.data(), cap, &n);
      if (rc == SDBG_ECAPACITY && n > cap) { cap = n; continue; }  // the call reports how many groups exist: one retry with room for them
      if (rc != SDBG_OK) throw GpuError(rc, std::string("sdbg_filter_groupby: ") + sdbg_last_error(sdbg_segment_context(segs_[0])));
      break;
    }
    groups_.resize(n);
    for (const auto& g : groups_) rows_scanned_ += g.count;
    ran_ = true;
  }
  const size_t take = std::min<size_t>(duckdb::STANDARD_VECTOR_SIZE, groups_.size() - cursor_);
  for (size_t i = 0; i < take; ++i) {
    const sdbg_group_row& g = groups_[cursor_ + i];
    output.key.push_back(g.key);
    output.count.push_back(int64_t(g.count));
    output.sum_lo.push_back(g.sum_i128[0]);
    output.sum_hi.push_back(g.sum_i128[1]);
    output.avg.push_back(g.cnt_f64 ? g.sum_f64 / double(g.cnt_f64) : 0.0);
  }
  output.size = take;
  cursor_ += take;
}

GpuAggGlobalState::GpuAggGlobalState(std::vector<sdbg_segment*> segments, std::vector<sdbg_col_pred> pushed_filters, uint64_t key_field,
                                     uint64_t sum_int_field, uint64_t avg_f64_field, uint32_t n_groups_hint)
    : segs(std::move(segments)), preds(std::move(pushed_filters)), key(key_field), sum_i(sum_int_field), avg_f(avg_f64_field),
      hint(n_groups_hint) {}

void GpuAggScanFunction(GpuAggGlobalState& g, GpuAggLocalState& l, duckdb::DataChunkMock& output) {
  output.Reset();
  std::call_once(g.ran, [&] {                 // an exception here leaves the flag unset: the next worker retries and rethrows
    uint64_t cap = std::max<uint64_t>(g.hint, 1024), n = 0;
    std::vector<sdbg_group_row> rows;
    for (;;) {
      rows.resize(cap);
      const int rc = sdbg_filter_groupby(g.segs.data(), g.segs.size(), g.preds.data(), g.preds.size(), g.key, g.hint, g.sum_i, g.avg_f,
                                         rows.data(), cap, &n);
      if (rc == SDBG_ECAPACITY && n > cap) { cap = n; continue; }
      if (rc != SDBG_OK) throw GpuError(rc, std::string("sdbg_filter_groupby: ") + sdbg_last_error(sdbg_segment_context(g.segs[0])));
      break;
    }
    rows.resize(n);
    g.groups = std::move(rows);
  });
  const size_t chunk = g.next_chunk.fetch_add(1, std::memory_order_relaxed);
  const size_t first = chunk * size_t(duckdb::STANDARD_VECTOR_SIZE);
  if (first >= g.groups.size()) return;                      // cardinality 0
  const size_t take = std::min<size_t>(duckdb::STANDARD_VECTOR_SIZE, g.groups.size() - first);
  for (size_t i = 0; i < take; ++i) {
    const sdbg_group_row& r = g.groups[first + i];
    output.key.push_back(r.key);
    output.count.push_back(int64_t(r.count));
    output.sum_lo.push_back(r.sum_i128[0]);
    output.sum_hi.push_back(r.sum_i128[1]);
    output.avg.push_back(r.cnt_f64 ? r.sum_f64 / double(r.cnt_f64) : 0.0);
  }
  output.size = take;
  ++l.chunks_claimed;
  g.rows_emitted.fetch_add(take, std::memory_order_relaxed);
}

}  // namespace sdbg_host
