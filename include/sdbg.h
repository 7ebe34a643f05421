/* include/sdbg.h -- C ABI of libsdbg.so: the B200 (sm_100a) implementation of SereneDB's
 * query-time hot path. Plain pointers and sizes only; no C++/torch types cross this boundary.
 *
 * What each entry point replaces in the reference (paths relative to /root/reference, "irs/" =
 * libs/iresearch/include/iresearch/):
 *
 *   sdbg_stage_postings   the per-segment open of the ".doc" stream + term metas that
 *                         PostingsReaderBase::prepare / ::decode perform
 *                         (irs/formats/posting/reader.hpp:100-226) -- index-load time.
 *   sdbg_stage_norms      NormColumnReader (irs/formats/column/norm_column_reader.hpp:43-108).
 *   sdbg_stage_column     ColumnReader open of a `.col` column (irs/formats/column/column_reader.hpp:90-256).
 *   sdbg_bm25_topk        the body of DocIterator::Collect for the WAND iterators built in
 *                         PostingsReaderImpl::WandIterator (irs/formats/posting/reader.hpp:457-501),
 *                         driven by irs::ExecuteTopK (irs/search/doc_collector.hpp:88-136) and by
 *                         CollectSegmentTopK (server/connector/duckdb_search_full_scan.cpp:1868-1921);
 *                         with `filt` it is TableFilterDocIterator::Collect
 *                         (irs/index/table_filter_iterator.cpp:450-475).
 *   sdbg_filter_bitmap    ColFilterChain::FilterWindow (irs/index/table_filter_iterator.cpp:147-264).
 *   sdbg_filter_count_sum RunCountScan / UNGROUPED_AGGREGATE over iresearch_scan
 *                         (server/connector/duckdb_search_full_scan.cpp:2201-2239).
 *   sdbg_filter_groupby   RunColScan + FullScanner::Scan feeding DuckDB's HASH_GROUP_BY
 *                         (duckdb_search_full_scan.cpp:2405-2433, server/connector/full_scanner.cpp:81-147).
 *
 * Conventions: every function returns 0 on success and a negative SDBG_E* code otherwise; it never
 * throws. The caller owns all host buffers. Handles are opaque. A context owns one CUDA device and
 * one stream; calls on one context are serialised by the caller (one context per worker thread,
 * like one DocIterator per (segment, query, worker) in the reference). There is NO CPU fallback:
 * without a CUDA device sdbg_init fails with SDBG_ENODEVICE.
 */
#ifndef SDBG_H_
#define SDBG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDBG_OK 0
#define SDBG_EINVAL (-1)    /* bad argument */
#define SDBG_ENODEVICE (-2) /* no CUDA device / wrong architecture */
#define SDBG_ECUDA (-3)     /* CUDA runtime error (see sdbg_last_error) */
#define SDBG_EFORMAT (-4)   /* corrupt posting stream */
#define SDBG_ENOTFOUND (-5) /* unknown field */
#define SDBG_ECAPACITY (-6) /* output buffer too small */
#define SDBG_EUNSUPPORTED (-7)

typedef struct sdbg_ctx sdbg_ctx;
typedef struct sdbg_segment sdbg_segment;

/* ---- lifecycle ---- */
int sdbg_init(int device, sdbg_ctx** out);
void sdbg_destroy(sdbg_ctx*);
const char* sdbg_last_error(const sdbg_ctx*);
const char* sdbg_version(void);
/* Device-side timing on the context's stream (CUDA events), so callers never guess the stream. */
int sdbg_timer_start(sdbg_ctx*);
int sdbg_timer_stop(sdbg_ctx*, float* ms);
int sdbg_sync(sdbg_ctx*);
/* Number of kernels this context has launched since creation (bench's gpu_launches claim). */
uint64_t sdbg_launch_count(const sdbg_ctx*);
/* Block-max (WAND / MaxScore) pruning, the `WandContext` of irs::ExecuteTopK (doc_collector.hpp:88). On by
 * default when the segment carries block-max data; results (hits) are identical either way, but with
 * pruning total_matches is a lower bound (wand_scoring_test.cpp:382-384). Level 1: blocks / windows whose
 * block-max bound cannot beat the threshold are dropped while planning (single-term queries:
 * SingleWandIterator's skip). Level 2 (default) adds MaxScore's essential / non-essential split for
 * disjunctions (max_score_iterator.hpp:406-508): once the threshold exceeds the global bound of the
 * largest list, that list is no longer scanned but probed per surviving candidate -- used for queries whose
 * largest list is bitset-encoded (a probe is then a bit test + popcount rank, no block decode). 0 = off. */
int sdbg_set_wand(sdbg_ctx*, int level);
/* Per-kernel device timing for roofline reports: when enabled, a CUDA event pair is recorded on the
 * context's stream around each hot kernel launch. kernel_id: 0 filter_groupby, 1 bm25_topk,
 * 2 topk_merge, 3 filter_count_sum. read() synchronises and returns the summed duration and the
 * number of launches since enable(1); enable() resets the counters. */
int sdbg_profile_enable(sdbg_ctx*, int on);
int sdbg_profile_read(sdbg_ctx*, int kernel_id, double* total_ms, uint64_t* launches);
/* Writes `bytes` of device memory (> L2) to evict cached inputs between timed iterations. */
int sdbg_flush_l2(sdbg_ctx*);

/* ---- staging (index-load time): host bytes are copied to HBM; the caller may free them ---- */
typedef struct {
  uint32_t docs_count;   /* TermMetaImpl::docs_count */
  uint32_t freq;         /* total term frequency */
  uint64_t doc_start;    /* offset of the term's stream in the .doc bytes handed in */
  uint64_t e_skip_start; /* e_single_doc when docs_count == 1 (same union as reader.hpp:213-217) */
} sdbg_term_meta;

int sdbg_segment_create(sdbg_ctx*, uint32_t docs_count, sdbg_segment** out);
void sdbg_segment_destroy(sdbg_segment*);
/* doc_file: the ".doc" stream (posting blocks + skip data, format "1_5simd"); has_wand != 0 when
 * the field was indexed with block-max data (optimize_top_k) in the (freq, norm) pair layout
 * (WandType::MinNorm / DivNorm, wand_writer.hpp:302-381). The pairs only bound the scores of the scorer
 * they were written for: as in the reference (the field's wand scorers are matched with Scorer::equals),
 * the caller enables pruning (sdbg_set_wand) only for queries of that scorer; pass 0 for a field whose wand
 * payload has another layout (BM15's freq-only entries). Builds the per-block offset table and copies
 * 16-byte-aligned block payloads to HBM. */
int sdbg_stage_postings(sdbg_segment*, const uint8_t* doc_file, size_t n, const sdbg_term_meta* terms,
                        size_t n_terms, int has_wand);
/* The segment's DocumentMask (index_meta.hpp:39-43: the set of deleted doc ids). Masked docs are neither
 * scored, collected nor counted by the BM25 calls, like SegmentReaderImpl::mask wrapping the query iterator
 * (segment_reader_impl.cpp:95-157,318-326; duckdb_search_full_scan.cpp:1898). n == 0 clears the mask. */
int sdbg_stage_docs_mask(sdbg_segment*, const uint32_t* deleted_docs, size_t n);
/* The b of the BM25 scorer the segment's block-max (wand) entries were written for (wand_writer.hpp:142-175; default
   0.75). Block-max pruning is used only for queries whose scorer has the same b -- the check Scorer::equals makes in
   PostingsReaderImpl::WandIterator (formats/posting/reader.hpp:457-501); any other scorer is evaluated exhaustively. */
int sdbg_segment_set_wand_b(sdbg_segment*, float wand_b);
/* Zonemap effect of the last GROUP BY scan: 2048-row blocks judged / proven dead from their min-max (never read). */
int sdbg_scan_stats(sdbg_ctx*, uint64_t* blocks_total, uint64_t* blocks_skipped);
/* The context a segment was created in (for sdbg_last_error after a failed call that only has segments at hand). */
sdbg_ctx* sdbg_segment_context(const sdbg_segment*);
typedef struct { uint8_t byte_size; uint32_t row_count; uint64_t file_offset; } sdbg_norm_rg; /* norm_writer.hpp:41-48 */
/* Row groups of fixed-width (1/2/4 B) little-endian field lengths; row = doc - 1. */
int sdbg_stage_norms(sdbg_segment*, const uint8_t* bytes, size_t n, const sdbg_norm_rg* rgs, size_t n_rg);
typedef enum { SDBG_I64 = 0, SDBG_F64 = 1, SDBG_I32 = 2 } sdbg_type;
/* validity may be NULL (NOT NULL column); otherwise bit r of validity[r/64] set => row r is valid. */
int sdbg_stage_column(sdbg_segment*, uint64_t field, sdbg_type t, const void* values,
                      const uint64_t* validity, uint64_t rows);
/* Same, but `d_values` already lives in device memory (borrowed, not copied, not freed). */
int sdbg_stage_column_device(sdbg_segment*, uint64_t field, sdbg_type t, const void* d_values, uint64_t rows);
/* Device address of a staged column (for callers that generate data in place). */
int sdbg_column_device_ptr(sdbg_segment*, uint64_t field, void** d_values, uint64_t* rows);
/* Copies the first `rows` values of a staged column back to host memory (tests / bench set-up). */
int sdbg_column_to_host(sdbg_segment*, uint64_t field, void* host_dst, uint64_t rows);
/* Bytes of HBM held by the segment's postings (payload + tables) and how many blocks were staged. */
int sdbg_segment_posting_stats(const sdbg_segment*, uint64_t* payload_bytes, uint64_t* table_bytes,
                               uint64_t* n_blocks, uint64_t* n_postings);

/* Encoded bytes (block headers + payloads, as in the .doc stream) of the first n_terms terms. */
int sdbg_segment_term_bytes(const sdbg_segment*, uint64_t* bytes_out, size_t n_terms);

/* ---- predicates (pushed TableFilterSet entries; NULL never passes) ---- */
enum { SDBG_OP_LT = 0, SDBG_OP_LE, SDBG_OP_GT, SDBG_OP_GE, SDBG_OP_EQ, SDBG_OP_NE, SDBG_OP_BETWEEN,
       SDBG_OP_IS_NULL, SDBG_OP_IS_NOT_NULL };
typedef struct {
  uint64_t field;
  int32_t op;
  int32_t is_float;
  int64_t lo_i, hi_i;
  double lo_f, hi_f;
} sdbg_col_pred;

/* ---- BM25 top-k (boundary B2, irs::DocIterator::Collect) ---- */
enum { SDBG_QUERY_OR = 0, SDBG_QUERY_AND = 1 };
typedef struct { float idf, norm_const, norm_length, boost; uint32_t term; } sdbg_bm25_term; /* BM25Stats (bm25.hpp:49-56) + boost */
typedef struct { float score; uint32_t doc; uint32_t seg; } sdbg_hit;                          /* irs::ScoreDoc (iterators.hpp:93-101) */

/* BM25::collect mirror (irs/search/bm25.cpp:279-310): corpus-wide statistics -> BM25Stats. idf is
 * computed in double and narrowed, avg_dl divides two floats, norm_const = k - k*b. boost is set to 1. */
int sdbg_bm25_collect(uint64_t docs_with_field, uint64_t total_term_freq, uint64_t docs_with_term, float k,
                      float b, sdbg_bm25_term* out);

/* (k1, b) are the scorer's parameters (BM25::k(), BM25::b()): like BM25::PrepareScorer (bm25.cpp:312-365) they
 * select the scoring form -- k1 == 0: BM1 (every score 0 without a filter boost, :112-126), b == 0: BM15
 * (c0 - c0 / (1 + freq / k1), no norms, :70-87), otherwise BM25 (:90-107). Block-max pruning is applied only to
 * the BM25 form: the staged (freq, norm) pairs were chosen for it (wand_type() differs per form, bm25.cpp:407-418).
 * One query over the segments of this GPU. Accepts docs with score > threshold_in (seed it with
 * FLT_MIN like doc_collector.hpp:102, or with the cross-worker threshold). out has room for k hits,
 * returned sorted by (score desc, seg asc, doc asc); *threshold_out = k-th score if k hits exist. */
int sdbg_bm25_topk(sdbg_segment* const* segs, size_t n_segs, int kind, const sdbg_bm25_term* terms,
                   size_t n_terms, float k1, float b, const sdbg_col_pred* filt, uint32_t k, float threshold_in,
                   sdbg_hit* out, uint32_t* n_out, uint64_t* total_matches, float* threshold_out);
/* A batch of independent queries in one launch set (the benchmark-game / many-workers shape).
 * Query q uses terms[term_off[q] .. term_off[q+1]); out holds n_queries*k hits, n_out/total per query. */
/* k1 = -1 is reserved: it selects the TFIDF scorer (b != 0: normalised) -- sdbg_tfidf_topk_batch is the named entry. */
int sdbg_bm25_topk_batch(sdbg_segment* const* segs, size_t n_segs, int kind, const sdbg_bm25_term* terms,
                         const uint32_t* term_off, size_t n_queries, float k1, float b, const sdbg_col_pred* filt,
                         uint32_t k, float threshold_in, sdbg_hit* out, uint32_t* n_out,
                         uint64_t* total_matches);
/* TFIDF (irs::TFIDF, search/tfidf.cpp): statistics (:149-150; only .idf and .boost of the term are used) and the same
 * batched scan scored with sqrt(freq) * boost * idf [/ sqrt(doc length) when normalize] (:59-80). Always exhaustive. */
int sdbg_tfidf_collect(uint64_t docs_with_field, uint64_t docs_with_term, sdbg_bm25_term* out);
int sdbg_tfidf_topk_batch(sdbg_segment* const* segs, size_t n_segs, int kind, const sdbg_bm25_term* terms,
                          const uint32_t* term_off, size_t n_queries, int normalize, const sdbg_col_pred* filt, uint32_t k,
                          float threshold_in, sdbg_hit* out, uint32_t* n_out, uint64_t* total_matches);
/* Streaming mode of the search scan (server/connector/duckdb_search_full_scan.cpp:2370-2403 RunStreamingScan, which
 * drains DocIterator::EmitScoredDocs, iterators.hpp:202-204): EVERY match of one query in docs [doc_min, doc_max) of
 * one segment with its BM25 score, ascending by doc id. Disjunctions of 1..4 terms, conjunctions of 1..16, hybrid
 * filter and deleted-doc mask honoured. *n_out = number of matches; SDBG_ECAPACITY when cap is too small (*n_out then
 * says how much room is needed; cap = 0 with NULL outputs is the count-only form). */
int sdbg_bm25_scan(sdbg_segment*, int kind, const sdbg_bm25_term* terms, size_t n_terms, float k1, float b,
                   const sdbg_col_pred* filt, uint32_t doc_min, uint32_t doc_max, uint32_t* out_docs, float* out_scores,
                   uint64_t cap, uint64_t* n_out);
/* Multi-GPU: leave each query's top-k on the device as sortable 64-bit keys + a base ordinal so a
 * collective can gather them; merge gathered keys from `n_ranks` ranks. */
int sdbg_bm25_topk_batch_device(sdbg_segment* const* segs, size_t n_segs, int kind,
                                const sdbg_bm25_term* terms, const uint32_t* term_off, size_t n_queries,
                                float k1, float b, const sdbg_col_pred* filt, uint32_t k, float threshold_in,
                                uint32_t rank, void* d_keys /* n_queries*k u64 */,
                                void* d_totals /* n_queries u64 */);
/* out may be NULL: the merged keys then stay in HBM (device-resident pipelines / timing). */
int sdbg_topk_merge_gathered(sdbg_ctx*, const void* d_keys_all /* n_ranks*n_queries*k u64 */,
                             uint32_t n_ranks, size_t n_queries, uint32_t k, sdbg_hit* out, uint32_t* n_out);
/* Test probe: decode+score one whole posting list (exhaustive, no top-k). Buffers sized docs_count. */
int sdbg_decode_score_term(sdbg_segment*, uint32_t term, float c0, float norm_const, float norm_length,
                           uint32_t* docs, uint32_t* freqs, float* scores);

/* Encoded int64 columns: frame-of-reference bit-packing in groups of 2048 rows -- per group a base (the minimum) and a
 * bit width, values stored as (v - base) in `bits` bits, little-endian, groups 8-byte aligned; bits = 0 is a constant
 * group. This is the algorithm of DuckDB's `bitpacking` codec in FOR mode, which the reference's column blocks name in
 * ColumnBlockMeta::codec (irs/formats/column/column_reader.hpp:90-96); DuckDB is not vendored in the reference tree, so
 * the byte layout is this library's. sdbg_pack_for is the host-side writer (returns SDBG_ECAPACITY with *n_words = the
 * room needed); sdbg_stage_column_for copies the packed stream to the GPU and decodes it there into a staged SDBG_I64
 * column, so only the packed bytes cross PCIe. */
typedef struct { int64_t base; uint32_t bits; uint32_t off8; } sdbg_for_block;
int sdbg_pack_for(const int64_t* values, uint64_t rows, sdbg_for_block* headers /* (rows + 2047) / 2048 */, uint64_t* words,
                  uint64_t cap_words, uint64_t* n_words);
int sdbg_stage_column_for(sdbg_segment*, uint64_t field, const sdbg_for_block* headers, const uint64_t* words, uint64_t n_words,
                          uint64_t rows);

/* Late materialisation (HitBatcher::MaterializeColumn, irs/index/hit_batcher.hpp:39; FinalizeBatch of the search scan):
 * out_values[i] = column[docs[i] - 1] for n hit docs of the segment (element width = the staged type's), out_valid[i]
 * (nullable) = 0 for NULL or out-of-range rows, whose value is written as 0. */
int sdbg_gather_column(sdbg_segment*, uint64_t field, const uint32_t* docs, size_t n, void* out_values, uint8_t* out_valid);

/* ---- columnar filter / aggregate (boundary B3, iresearch_scan) ---- */
int sdbg_filter_bitmap(sdbg_segment*, const sdbg_col_pred* preds, size_t n_preds, uint64_t* mask_out);
int sdbg_filter_count_sum(sdbg_segment* const* segs, size_t n_segs, const sdbg_col_pred* preds,
                          size_t n_preds, uint64_t sum_field, uint64_t* count, int64_t sum_i128[2],
                          double* sum_f64);
typedef struct { int64_t key; uint64_t count; int64_t sum_i128[2]; double sum_f64; uint64_t cnt_f64; } sdbg_group_row;
/* SELECT key, COUNT(*), SUM(sum_int_field), SUM(avg_f64_field)/cnt_f64 ... GROUP BY key.
 * Rows come back sorted by key. Pass UINT64_MAX for an aggregate field that is not wanted. */
int sdbg_filter_groupby(sdbg_segment* const* segs, size_t n_segs, const sdbg_col_pred* preds,
                        size_t n_preds, uint64_t key_field, uint32_t n_groups_hint,
                        uint64_t sum_int_field, uint64_t avg_f64_field, sdbg_group_row* out, uint64_t cap,
                        uint64_t* n_out);
/* Multi-GPU split of the same: partial dense aggregates stay on the device in two flat buffers that
 * a SUM all-reduce can merge (int64 limbs + counts, and float64 sums), then finalize on any rank.
 * d_i64: 4*span int64 = [count | sum_lo | sum_hi | cnt_f64]; d_f64: span float64. */
int sdbg_filter_groupby_partial(sdbg_segment* const* segs, size_t n_segs, const sdbg_col_pred* preds,
                                size_t n_preds, uint64_t key_field, int64_t key_min, uint64_t key_span,
                                uint64_t sum_int_field, uint64_t avg_f64_field, void* d_i64, void* d_f64);
int sdbg_groupby_finalize(sdbg_ctx*, int64_t key_min, uint64_t key_span, const void* d_i64,
                          const void* d_f64, sdbg_group_row* out, uint64_t cap, uint64_t* n_out);
/* min/max of a staged int column (zonemap-style statistics gathered at staging). */
int sdbg_column_minmax_i64(sdbg_segment*, uint64_t field, int64_t* mn, int64_t* mx);

/* ---- host-side writer mirror + deterministic synthetic inputs (index-build side; not timed) ---- */
/* PostingsWriter mirror (irs/formats/posting/writer.hpp): builds a ".doc" stream on the host. */
typedef struct sdbg_writer sdbg_writer;
int sdbg_writer_create(uint32_t segment_docs, int has_wand, float wand_b, const uint32_t* norms /* per doc, may be NULL */,
                       sdbg_writer** out);
void sdbg_writer_destroy(sdbg_writer*);
int sdbg_writer_add_term(sdbg_writer*, const uint32_t* docs, const uint32_t* freqs, uint32_t n);
int sdbg_writer_finish(sdbg_writer*, const uint8_t** doc_file, size_t* n, const sdbg_term_meta** terms, size_t* n_terms);
/* Synthetic corpus shard of SURVEY §8d: docs (doc0, doc0+n], terms [t0, t0+nt); fills norms (u8,
 * dl<=255) and stages everything into `seg`. Returns per-term docs_count and the shard's sum of dl. */
int sdbg_synth_corpus(sdbg_segment* seg, uint64_t doc0, uint32_t n_docs, uint32_t t0, uint32_t nt,
                      int threads, uint32_t* docs_count_out /* nt */, uint64_t* sum_dl_out);
/* Same with a probability floor: p_t = max(p_floor, min(0.5, 0.6 / (t + 1))) (flat tail; an index far larger than L2). */
int sdbg_synth_corpus_ex(sdbg_segment* seg, uint64_t doc0, uint32_t n_docs, uint32_t t0, uint32_t nt, int threads, double p_floor,
                         uint32_t* docs_count_out /* nt */, uint64_t* sum_dl_out);
/* Synthetic table column generated directly in HBM: kind as in SURVEY §8d (0 k,1 a,2 b,3 v,4 w,5+ raw),
 * or 6 = int32 n = h % 1000000 (hybrid INCLUDE column, stream 2). */
int sdbg_synth_column(sdbg_segment* seg, uint64_t field, uint64_t stream, int kind, uint64_t row0, uint64_t rows);
uint64_t sdbg_synth_hash(uint64_t stream, uint64_t index);
/* Host-only probe of the staging parser (no device): block table of a ".doc" stream, for tests. */
int sdbg_debug_stage_host(const uint8_t* doc_file, size_t n, const sdbg_term_meta* terms, size_t n_terms, int has_wand,
                          uint32_t cap, uint32_t* n_blocks, uint32_t* term_blk_begin, uint32_t* last_doc,
                          uint32_t* prev_last, uint32_t* packed, uint32_t* max_freq, uint32_t* max_norm,
                          uint64_t* arena_bytes);

/* ---- collectives over NVLink (NCCL, resolved at run time with dlopen: libsdbg.so does not link it) -------------
   One communicator per context, everything enqueued on the context's stream. A C++ host needs nothing but these
   calls: rank 0 creates the id, the host ships its 128 bytes to the other ranks by whatever channel it has
   (the reference's own RPC, MPI, a file), every rank calls sdbg_dist_init. */
#define SDBG_DIST_ID_BYTES 128
int sdbg_dist_unique_id(uint8_t* id128);
int sdbg_dist_init(sdbg_ctx*, const uint8_t* id128, int rank, int world);
int sdbg_dist_destroy(sdbg_ctx*);
int sdbg_dist_allreduce_i64(sdbg_ctx*, void* d_buf, size_t n);                          /* in place, SUM */
int sdbg_dist_allgather(sdbg_ctx*, const void* d_send, void* d_recv, size_t bytes_per_rank);
/* Dense GROUP BY partials (sdbg_filter_groupby_partial) of every rank -> global partials on every rank with ONE
   ncclAllReduce: counts, SUM(int) limbs and SUM(double) as 120-bit fixed point share one int64 buffer (exact and
   independent of the rank order). abs_bound >= |SUM(double column)| over all ranks, identical on every rank. */
int sdbg_dist_groupby_merge(sdbg_ctx*, void* d_i64, void* d_f64, uint64_t span, double abs_bound);
/* BM25 top-k over the segments of ALL ranks: local scan -> one all-gather of the k best keys per query -> local
   selection, back to back on the context's stream. hit.seg = rank, hit.doc = ordinal within the rank. out == NULL:
   nothing is copied and nothing waits (results stay in HBM). */
int sdbg_dist_bm25_topk_batch(sdbg_segment* const* segs, size_t n_segs, int kind, const sdbg_bm25_term* terms,
                              const uint32_t* term_off, size_t n_queries, float k1, float b, const sdbg_col_pred* filt,
                              uint32_t k, float threshold_in, sdbg_hit* out, uint32_t* n_out);

#ifdef __cplusplus
}
#endif
#endif /* SDBG_H_ */
