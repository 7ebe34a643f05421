/* oracle/oracle.h -- C API of the CPU oracle.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT. Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product (libsdbg.so)
 * never links or calls it.
 *
 * The oracle is a CPU restatement of SereneDB's query-time hot path:
 *   - posting block codec "1_5simd"   libs/iresearch/include/iresearch/formats/posting/format_block_128.hpp:57-640
 *   - postings writer + skip list     .../formats/posting/writer.hpp:305-331,443-488,617-641,699-779
 *                                     .../formats/posting/skip_list.hpp:93-118, skip_list.cpp:38-94
 *   - block-max (wand) data           .../formats/posting/wand_writer.hpp:69-75,142-175,196-206,302-327,366-381
 *   - BM25 statistics + arithmetic    .../search/bm25.cpp:90-107,279-310
 *   - top-k collector                 .../index/iterators.hpp:103-250, search/doc_collector.hpp:88-136
 *   - columnar filter / COUNT / SUM / GROUP BY  (DuckDB semantics; DEFINITION, see below)
 * Paths are relative to /root/reference.
 *
 * Parity status:
 *   - bit packing: pinned against the reference's own simdcomp sources compiled into
 *     oracle/_ref/libsimdcomp_ref.so (tests/test_oracle_codec.py).
 *   - BM25 arithmetic: pinned against the reference's sqllogic goldens 1.9693236,
 *     1.6739764, 0.8266785 (tests/golden/bm25_goldens.json).
 *   - collector: pinned against block_scoring_test.cpp:427-493 known answers.
 *   - filter/COUNT/SUM: pinned against search_table_scan_10k.test:34-110.
 *   - StreamVByte (un-vendored submodule) and GROUP BY SUM/AVG (DuckDB, un-vendored):
 *     "parity unpinned" -- restated from the public format / documented SQL semantics.
 */
#ifndef SDB_ORACLE_H_
#define SDB_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------- bit packing (third_party/simdcomp layout) ---------- */
void orc_pack128(const uint32_t* in128, uint32_t* out_words /*4*bits*/, uint32_t bits);
void orc_unpack128(const uint32_t* in_words, uint32_t* out128, uint32_t bits);
void orc_pack128_d1(uint32_t prev, const uint32_t* in128, uint32_t* out_words, uint32_t bits);
void orc_unpack128_d1(uint32_t prev, const uint32_t* in_words, uint32_t* out128, uint32_t bits);
/* Optional: route 128-value unpack through the reference's own simdcomp (oracle/_ref).
 * Returns 0 on success. Used to validate the restatement and for the "reference" CPU arm. */
int orc_use_simdcomp_ref(const char* so_path);

/* ---------- StreamVByte "1234" (public format; submodule absent => parity unpinned) ---------- */
size_t orc_svb_encode(const uint32_t* in, uint32_t len, uint8_t* out);
size_t orc_svb_decode(const uint8_t* in, uint32_t* out, uint32_t len);
size_t orc_svb_delta_encode(const uint32_t* in, uint32_t len, uint8_t* out, uint32_t prev);
size_t orc_svb_delta_decode(const uint8_t* in, uint32_t* out, uint32_t len, uint32_t prev);

/* ---------- 128-doc block codec (format_block_128.hpp) ---------- */
/* All return the number of bytes written / consumed (header byte included). */
size_t orc_encode_doc_block(const uint32_t* docs, uint32_t len, uint32_t prev, uint8_t* out);
size_t orc_decode_doc_block(const uint8_t* in, uint32_t len, uint32_t prev, uint32_t* docs_out);
size_t orc_encode_freq_block(const uint32_t* freqs, uint32_t len, uint8_t* out);
size_t orc_decode_freq_block(const uint8_t* in, uint32_t len, uint32_t* freqs_out);

/* ---------- BM25 (bm25.cpp) ---------- */
typedef struct { float idf, norm_const, norm_length; } orc_bm25_stats;
void orc_bm25_collect(uint64_t docs_with_field, uint64_t total_term_freq, uint64_t docs_with_term,
                      float k, float b, orc_bm25_stats* out);
/* c0 = boost*(k+1)*idf (bm25.cpp:224); r = c0 - c0*c1/(c1+freq), c1 = norm_const + norm_length*norm */
float orc_bm25_num(float k, float boost, float idf);
void orc_bm25_score(const uint32_t* freq, const uint32_t* norm, uint32_t n, float num,
                    float norm_const, float norm_length, float* out);

/* ---------- top-k collector (iterators.hpp:103-250) ---------- */
typedef struct { float score; uint32_t doc; uint32_t seg; } orc_hit;
/* Faithful NthPartitionScoreCollector: feeds (score,doc) in order, returns accepted count in
 * hits (capacity 2k), sorted by score desc (stable on ties is NOT guaranteed by the reference). */
uint64_t orc_collect_nth(const float* scores, const uint32_t* docs, uint64_t n, uint32_t k,
                         float threshold_in, orc_hit* hits_2k, uint32_t* accepted,
                         float* threshold_out);

/* ---------- segment: postings + norms + columns ---------- */
typedef struct orc_segment orc_segment;
typedef struct {
  uint32_t docs_count;   /* TermMetaImpl::docs_count */
  uint32_t freq;         /* total term frequency in the segment */
  uint64_t doc_start;    /* offset of the term's stream in .doc */
  uint64_t e_skip_start; /* == e_single_doc when docs_count == 1 (union in the reference) */
} orc_term_meta;

orc_segment* orc_segment_new(uint32_t docs_count /*N; doc ids are 1..N*/, int has_wand,
                             float wand_b /*BM25 b used by the block-max producer*/);
void orc_segment_free(orc_segment*);
/* norms[d-1] = field length of doc d (raw token count). Must be set before adding terms when
 * has_wand (the producer reads norms, wand_writer.hpp:289-293). byte width chosen like
 * norm_column_reader.hpp: smallest of 1/2/4 that holds max. */
void orc_segment_set_norms(orc_segment*, const uint32_t* norms);
/* Appends one term (docs ascending, 1-based) through the writer restatement. Returns term index. */
int64_t orc_segment_add_term(orc_segment*, const uint32_t* docs, const uint32_t* freqs, uint32_t n);
const uint8_t* orc_segment_doc_bytes(const orc_segment*, uint64_t* size);
uint32_t orc_segment_num_terms(const orc_segment*);
void orc_segment_term_meta(const orc_segment*, uint32_t term, orc_term_meta* out);
uint32_t orc_segment_docs(const orc_segment*);
uint64_t orc_segment_norm_sum(const orc_segment*);
const uint8_t* orc_segment_norm_bytes(const orc_segment*, uint32_t* byte_width);
/* Decode a whole posting list through the reader restatement. Returns docs_count. */
uint32_t orc_segment_decode_term(const orc_segment*, uint32_t term, uint32_t* docs, uint32_t* freqs);
/* Level-0 skip entries (one per full block that has a successor) + root block-max entry.
 * last_doc/doc_ptr/wand_freq/wand_norm arrays sized (docs_count-1)/128. Returns entry count. */
uint32_t orc_segment_skip_level0(const orc_segment*, uint32_t term, uint32_t* last_doc,
                                 uint64_t* doc_ptr, uint32_t* wand_freq, uint32_t* wand_norm,
                                 uint32_t* root_freq, uint32_t* root_norm, uint32_t* num_levels);
/* INCLUDE / table columns. type: 0=int64 1=float64 2=int32. validity may be NULL (all valid);
 * otherwise bit r of validity[r/64] set => row r is NOT NULL. Values are copied. */
/* DocumentMask of the segment (deleted doc ids): masked docs are not scored, collected or counted by the
 * BM25 calls (SegmentReaderImpl::mask, segment_reader_impl.cpp:318-326). n == 0 clears. */
int orc_segment_set_docs_mask(orc_segment*, const uint32_t* deleted_docs, size_t n);
int orc_segment_add_column(orc_segment*, uint64_t field, int type, const void* values,
                           const uint64_t* validity, uint64_t rows);

/* ---------- predicates ---------- */
enum { ORC_OP_LT = 0, ORC_OP_LE, ORC_OP_GT, ORC_OP_GE, ORC_OP_EQ, ORC_OP_NE, ORC_OP_BETWEEN,
       ORC_OP_IS_NULL, ORC_OP_IS_NOT_NULL };
typedef struct {
  uint64_t field;
  int32_t op;
  int32_t is_float; /* compare as double (lo_f/hi_f) or int64 (lo_i/hi_i) */
  int64_t lo_i, hi_i;
  double lo_f, hi_f;
} orc_pred;

/* ---------- BM25 top-k over segments (doc_collector.hpp:88-136 semantics, canonical ties) ---------- */
enum { ORC_QUERY_OR = 0, ORC_QUERY_AND = 1 };
typedef struct { float idf, norm_const, norm_length, boost; uint32_t term; } orc_bm25_term;
/* mode: 0 = exhaustive, dense accumulators; 1 = exhaustive, 4096-doc windows (CPU-baseline leg);
 *       2 = block-max pruned (single term: SingleWandIterator semantics; multi-term OR: window
 *           skipping by summed block-max), used for the "pruned == exhaustive" differential.
 * Results: hits sorted by (score desc, seg asc, doc asc); n_out = min(k, matches above threshold_in).
 * (k1, b) select the scoring form like BM25::PrepareScorer: k1 == 0 BM1 (scores 0), b == 0 BM15, else BM25;
 * block-max pruning (mode 2) only applies to the BM25 form.
 * Sum order for multi-term: ascending docs_count, ties by position in `terms` (conjunction.hpp:520-523). */
int orc_bm25_topk(orc_segment* const* segs, size_t n_segs, int kind, const orc_bm25_term* terms,
                  size_t n_terms, float k1, float b, const orc_pred* filt, uint32_t k, float threshold_in,
                  int mode, orc_hit* out, uint32_t* n_out, uint64_t* total_matches,
                  uint64_t* postings_scored);

/* Many queries on `threads` host cores (one query per thread at a time). out: n_queries*k hits. */
int orc_bm25_topk_batch(orc_segment* const* segs, size_t n_segs, int kind, const orc_bm25_term* terms,
                        const uint32_t* term_off, size_t n_queries, float k1, float b, const orc_pred* filt, uint32_t k,
                        float threshold_in, int mode, int threads, orc_hit* out, uint32_t* n_out,
                        uint64_t* total_matches, uint64_t* postings_scored);

/* ---------- columnar (full_scanner.cpp:81-147 + DuckDB aggregate; see header note) ---------- */
int orc_filter_bitmap(const orc_segment*, const orc_pred* preds, size_t n_preds, uint64_t* mask_out);
/* sum_field type decides which sum is filled. NULL sum inputs are skipped (SQL). threads>=1. */
int orc_filter_count_sum(orc_segment* const* segs, size_t n_segs, const orc_pred* preds,
                         size_t n_preds, uint64_t sum_field, int threads, uint64_t* count,
                         int64_t sum_i128[2], double* sum_f64);
typedef struct { int64_t key; uint64_t count; int64_t sum_i128[2]; double sum_f64; uint64_t cnt_f64; } orc_group_row;
/* rows come back sorted by key; *n_out groups; cap = capacity of out. */
int orc_filter_groupby(orc_segment* const* segs, size_t n_segs, const orc_pred* preds, size_t n_preds,
                       uint64_t key_field, uint64_t sum_int_field, uint64_t avg_f64_field,
                       int threads, orc_group_row* out, uint64_t cap, uint64_t* n_out);

/* ---------- deterministic synthetic inputs, seed 0x5EDB2026 ---------- */
/* 1: evaluate c1 = norm_const + norm_length * norm of the BM25 form as one fused multiply-add (what a clang -mfma
   build of bm25.cpp:105 does); 0 (default): the source order without contraction. Test-only switch. */
void orc_set_contract(int on);
/* TFIDF statistics (search/tfidf.cpp:149-150). The top-k entry points select TFIDF with k1 = -1 (b != 0: normalised). */
float orc_tfidf_idf(uint64_t docs_with_field, uint64_t docs_with_term);
uint32_t orc_count_max_levels(uint64_t skip_0, uint64_t skip_n, uint64_t count);
uint64_t orc_synth_hash(uint64_t stream, uint64_t index);
/* kind: 0 k=h%100000, 1 a=h%1e6, 2 b in [0,1), 3 v=(h%2001)-1000, 4 w in [0,1000), 5.. raw int64 */
void orc_synth_column(uint64_t stream, int kind, uint64_t row0, uint64_t rows, void* out);
/* doc lengths dl(d) = 16 + h%240 for docs doc0+1 .. doc0+n (global doc numbering). */
void orc_synth_doc_lengths(uint64_t doc0, uint32_t n, uint32_t* out);
/* posting list of synthetic term t restricted to global docs (doc0, doc0+n]; local ids 1..n.
 * Returns count; docs/freqs sized n. p_t = min(0.5, 0.6/(t+1)). */
uint32_t orc_synth_term(uint32_t t, uint64_t doc0, uint32_t n, const uint32_t* dl, uint32_t* docs,
                        uint32_t* freqs);

/* Whole synthetic shard built with `threads` workers (same bytes as term-by-term building). */
orc_segment* orc_synth_segment(uint64_t doc0, uint32_t n_docs, uint32_t t0, uint32_t nt, int threads,
                               uint32_t* docs_count_out, uint64_t* sum_dl_out);

#ifdef __cplusplus
}
#endif
#endif /* SDB_ORACLE_H_ */
