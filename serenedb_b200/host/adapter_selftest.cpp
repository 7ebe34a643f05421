// adapter_selftest.cpp -- drives the two adapters the way the reference's callers would
// (ExecuteTopK's collector loop, IResearchScanFunction's chunk loop) and prints results as JSON lines
// for tests/test_gpu_adapters.py to compare with the oracle. Needs a GPU at run time.
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "gpu_adapters.hpp"

namespace {
struct ListCollector final : irs::ScoreCollector {  // a trivial ScoreCollector: keeps what it is fed
  ListCollector() : irs::ScoreCollector(Tag::Generic) {}
  std::vector<irs::ScoreDoc> docs;
  void Add(irs::score_t s, irs::doc_id_t d) override { docs.push_back({s, d, 0}); }
  void AddWindow(const irs::score_t*, const uint64_t*, irs::doc_id_t, size_t, bool) override {}
  void AddDocs(const irs::doc_id_t* d, size_t n, const irs::score_t* s) override { for (size_t i = 0; i < n; ++i) docs.push_back({s[i], d[i], 0}); }
};
}  // namespace

int main(int argc, char** argv) {
  const uint32_t n_docs = argc > 1 ? uint32_t(std::atoi(argv[1])) : 200000;
  sdbg_ctx* ctx = nullptr;
  int rc = sdbg_init(0, &ctx);
  if (rc != SDBG_OK) { std::printf("{\"error\": %d}\n", rc); return rc == SDBG_ENODEVICE ? 3 : 1; }
  sdbg_segment* seg = nullptr;
  sdbg_segment_create(ctx, n_docs, &seg);
  std::vector<uint32_t> dc(8);
  uint64_t sum_dl = 0;
  rc = sdbg_synth_corpus(seg, 0, n_docs, 0, 8, 4, dc.data(), &sum_dl);
  if (rc) { std::printf("{\"error\": %d}\n", rc); return 1; }
  sdbg_synth_column(seg, 9, 2, 6, 1, n_docs);
  // --- top-k through the DocIterator adapter ---
  std::vector<sdbg_bm25_term> terms(2);
  const uint32_t ids[2] = {2, 5};
  for (int i = 0; i < 2; ++i) { sdbg_bm25_collect(n_docs, sum_dl, dc[ids[i]], 1.2f, 0.75f, &terms[size_t(i)]); terms[size_t(i)].term = ids[i]; }
  sdbg_col_pred filt{}; filt.field = 9; filt.op = SDBG_OP_BETWEEN; filt.lo_i = 250000; filt.hi_i = 749999;
  sdbg_host::GpuTopKIterator it(seg, SDBG_QUERY_OR, terms, 1.2f, 0.75f, 100, &filt);
  ListCollector col;
  irs::ScoreFunction sf; irs::ColumnArgsFetcher fetcher;
  it.Collect(sf, fetcher, col);
  std::printf("{\"topk\": [");
  for (size_t i = 0; i < col.docs.size(); ++i) std::printf("%s[%u, %.9g]", i ? ", " : "", col.docs[i].doc, double(col.docs[i].score));
  // attributes through the reference's accessor, and the bitmap window of the first 4096 docs
  const auto* thr = irs::get<irs::ScoreThresholdAttr>(it);
  const auto* cost = irs::get<irs::CostAttr>(it);
  sdbg_host::GpuTopKIterator it2(seg, SDBG_QUERY_OR, terms, 1.2f, 0.75f, 100, &filt);
  std::vector<uint64_t> mask(64, 0);
  std::vector<irs::score_t> window(4096, 0.f);
  irs::FillBlockScoreContext sc; sc.score_window = window.data(); sc.merge_type = irs::ScoreMergeType::Sum;
  const auto fb = it2.FillBlock(1, 4097, mask.data(), sc, irs::FillBlockMatchContext{});
  unsigned bits = 0; double wsum = 0;
  for (uint64_t w : mask) bits += unsigned(__builtin_popcountll(w));
  for (float w : window) wsum += w;
  std::printf("], \"total\": %llu, \"threshold\": %.9g, \"attr_threshold\": %.9g, \"attr_cost\": %llu, \"fill_bits\": %u, \"fill_sum\": %.9g, \"fill_next\": %u}\n",
              static_cast<unsigned long long>(it.total_matches()), double(it.threshold().value), double(thr ? thr->value : -1.f),
              static_cast<unsigned long long>(cost ? cost->estimate() : 0), bits, wsum, fb.first);
  // --- streaming mode: drain EmitScoredDocs in STANDARD_VECTOR_SIZE-bounded windows like StreamScanLocalState::EmitChunk ---
  {
    sdbg_host::GpuTopKIterator st(seg, SDBG_QUERY_OR, terms, 1.2f, 0.75f, 0, &filt);
    std::vector<irs::doc_id_t> d(2048);
    std::vector<irs::score_t> sc(2048);
    uint64_t n = 0, chunks = 0, doc_sum = 0; double score_sum = 0; bool ordered = true; irs::doc_id_t last = 0;
    for (irs::doc_id_t lo = 1; lo <= n_docs; lo += 2048) {      // a 2048-doc window can hold at most 2048 matches
      const uint32_t got = st.EmitScoredDocs(d.data(), sc.data(), lo + 2048, sf, &fetcher, lo);
      if (got) ++chunks;
      for (uint32_t i = 0; i < got; ++i) { ordered = ordered && d[i] > last; last = d[i]; doc_sum += d[i]; score_sum += sc[i]; }
      n += got;
    }
    std::printf("{\"stream_n\": %llu, \"stream_chunks\": %llu, \"stream_doc_sum\": %llu, \"stream_score_sum\": %.12g, \"stream_ordered\": %d, \"stream_count\": %u}\n",
                static_cast<unsigned long long>(n), static_cast<unsigned long long>(chunks), static_cast<unsigned long long>(doc_sum), score_sum,
                ordered ? 1 : 0, st.count());
  }
  // --- aggregate scan through the table-function adapter ---
  for (uint64_t f = 10; f <= 14; ++f) sdbg_synth_column(seg, f, f, int(f - 10), 0, n_docs);
  std::vector<sdbg_col_pred> preds(2);
  preds[0].field = 11; preds[0].op = SDBG_OP_LT; preds[0].lo_i = 500000;
  preds[1].field = 12; preds[1].op = SDBG_OP_GE; preds[1].is_float = 1; preds[1].lo_f = 0.25;
  sdbg_host::GpuAggScan scan({seg}, preds, 10, 13, 14, 100000);
  duckdb::DataChunkMock chunk;
  uint64_t groups = 0, rows = 0, chunks = 0; __int128 sum = 0; double avg_sum = 0;
  for (;;) {
    scan.Scan(chunk);
    if (chunk.size == 0) break;
    ++chunks;
    for (size_t i = 0; i < chunk.size; ++i) { ++groups; rows += uint64_t(chunk.count[i]); sum += (static_cast<__int128>(chunk.sum_hi[i]) << 64) + static_cast<unsigned long long>(chunk.sum_lo[i]); avg_sum += chunk.avg[i]; }
    if (chunk.size > duckdb::STANDARD_VECTOR_SIZE) return 2;
  }
  std::printf("{\"groups\": %llu, \"rows\": %llu, \"chunks\": %llu, \"sum_v\": %lld, \"avg_sum\": %.12g}\n", static_cast<unsigned long long>(groups),
              static_cast<unsigned long long>(rows), static_cast<unsigned long long>(chunks), static_cast<long long>(sum), avg_sum);
  // --- the same scan mode driven by four concurrent workers (one global state, a local state each) ---
  {
    sdbg_host::GpuAggGlobalState gs({seg}, preds, 10, 13, 14, 100000);
    constexpr int kWorkers = 4;
    uint64_t w_groups[kWorkers] = {}, w_rows[kWorkers] = {}, w_chunks[kWorkers] = {};
    long long w_sum[kWorkers] = {};
    bool w_ok[kWorkers] = {};
    std::vector<std::thread> pool;
    for (int w = 0; w < kWorkers; ++w)
      pool.emplace_back([&, w] {
        sdbg_host::GpuAggLocalState ls;
        duckdb::DataChunkMock out;
        bool ok = true;
        for (;;) {
          sdbg_host::GpuAggScanFunction(gs, ls, out);
          if (out.size == 0) break;
          ok = ok && out.size <= duckdb::STANDARD_VECTOR_SIZE;
          for (size_t i = 0; i < out.size; ++i) { ++w_groups[w]; w_rows[w] += uint64_t(out.count[i]); w_sum[w] += out.sum_lo[i]; }
        }
        w_chunks[w] = ls.chunks_claimed;
        w_ok[w] = ok;
      });
    for (auto& th : pool) th.join();
    uint64_t tg = 0, tr = 0, tc = 0; long long ts = 0; bool ok = true;
    for (int w = 0; w < kWorkers; ++w) { tg += w_groups[w]; tr += w_rows[w]; tc += w_chunks[w]; ts += w_sum[w]; ok = ok && w_ok[w]; }
    std::printf("{\"mt_groups\": %llu, \"mt_rows\": %llu, \"mt_chunks\": %llu, \"mt_sum_v\": %lld, \"mt_ok\": %d, \"mt_emitted\": %llu}\n",
                static_cast<unsigned long long>(tg), static_cast<unsigned long long>(tr), static_cast<unsigned long long>(tc), ts, ok ? 1 : 0,
                static_cast<unsigned long long>(gs.rows_emitted.load()));
  }
  sdbg_segment_destroy(seg);
  sdbg_destroy(ctx);
  return 0;
}
