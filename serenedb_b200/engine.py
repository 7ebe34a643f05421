"""Host-side mirror of the reference's interfaces for the hot path, over the C ABI (include/sdbg.h).

Names follow the reference: `BM25` (irs/search/bm25.hpp:58), `ExecuteTopK`
(irs/search/doc_collector.hpp:88-136), `ScoreDoc` hits (irs/index/iterators.hpp:93-101), the
`iresearch_scan` column scan with pushed `TableFilterSet` predicates
(server/connector/duckdb_table_function.cpp:1178-1219). Everything that touches postings or rows runs
in libsdbg.so on the GPU; this module only marshals arguments.
"""
import ctypes as C

import numpy as np

from . import _native as N

FLT_MIN = float(np.finfo(np.float32).tiny)  # doc_collector.hpp:102 initial score_threshold
HIT_DTYPE = np.dtype([("score", "<f4"), ("doc", "<u4"), ("seg", "<u4")])
GROUP_DTYPE = np.dtype([("key", "<i8"), ("count", "<u8"), ("sum_lo", "<i8"), ("sum_hi", "<i8"),
                        ("sum_f64", "<f8"), ("cnt_f64", "<u8")])
TERM_META_DTYPE = np.dtype([("docs_count", "<u4"), ("freq", "<u4"), ("doc_start", "<u8"),
                            ("e_skip_start", "<u8")])
OPS = dict(LT=0, LE=1, GT=2, GE=3, EQ=4, NE=5, BETWEEN=6, IS_NULL=7, IS_NOT_NULL=8)
TYPES = {np.dtype("int64"): 0, np.dtype("float64"): 1, np.dtype("int32"): 2}
OR, AND = 0, 1
NO_FIELD = N.UINT64_MAX


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def pred(field, op, lo=0, hi=0):
    """One pushed column predicate (duckdb TableFilter). Float bounds select a double comparison on a double column;
    on an integer column the kernels compare integers, so the integer bounds are the tightest integers with the same
    truth set (v < 2.5 <=> v < 3, v >= 2.5 <=> v >= 3, v <= 2.5 <=> v <= 2, v > 2.5 <=> v > 2)."""
    import math
    p = N.ColPred()
    p.field = int(field)
    p.op = OPS[op] if isinstance(op, str) else int(op)
    is_float = isinstance(lo, float) or isinstance(hi, float)
    p.is_float = 1 if is_float else 0
    p.lo_f, p.hi_f = float(lo), float(hi)
    if not is_float:
        p.lo_i, p.hi_i = int(lo), int(hi)
        return p
    name = {v: k for k, v in OPS.items()}.get(p.op, "")
    big = (1 << 62)
    clamp = lambda x: int(max(-big, min(big, x)))
    lo_f, hi_f = float(lo), float(hi)
    if math.isnan(lo_f) or math.isinf(lo_f) or math.isnan(hi_f) or math.isinf(hi_f):
        p.lo_i, p.hi_i = clamp(-big if lo_f < 0 else big) if not math.isnan(lo_f) else 0, 0
        return p
    if name in ("LT", "GE"):
        p.lo_i = clamp(math.ceil(lo_f))
    elif name in ("LE", "GT"):
        p.lo_i = clamp(math.floor(lo_f))
    elif name == "BETWEEN":
        p.lo_i, p.hi_i = clamp(math.ceil(lo_f)), clamp(math.floor(hi_f))
    elif name in ("EQ", "NE"):
        # = / <> with a fractional constant on an integer column is a constant predicate the planner folds away
        # (DuckDB does); it is not representable here, so reject it instead of comparing with a rounded value
        if lo_f != math.floor(lo_f):
            raise ValueError("fold '= / <> fractional constant' on an integer column before pushing it down")
        p.lo_i = clamp(lo_f)
    return p


def _pred_array(preds):
    arr = (N.ColPred * max(len(preds), 1))()
    for i, p in enumerate(preds):
        arr[i] = p
    return arr


class Context:
    """One GPU + one stream (one per worker, like one DocIterator per worker in the reference)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        N.check(N.lib().sdbg_init(int(device), C.byref(self._h)))
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            N.lib().sdbg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        N.check(N.lib().sdbg_sync(self._h), self._h)

    def timer_start(self):
        N.check(N.lib().sdbg_timer_start(self._h), self._h)

    def timer_stop(self):
        ms = C.c_float()
        N.check(N.lib().sdbg_timer_stop(self._h, C.byref(ms)), self._h)
        return ms.value

    def flush_l2(self):
        N.check(N.lib().sdbg_flush_l2(self._h), self._h)

    def set_wand(self, enabled=True):
        """Block-max pruning on/off (WandContext of irs::ExecuteTopK). Off => exact total_matches."""
        N.check(N.lib().sdbg_set_wand(self._h, int(enabled)), self._h)

    def profile(self, on=True):
        N.check(N.lib().sdbg_profile_enable(self._h, 1 if on else 0), self._h)

    def scan_stats(self):
        """(2048-row blocks judged by the zonemap pass of the last GROUP BY scan, blocks proven dead = never read)."""
        a, b = C.c_uint64(), C.c_uint64()
        N.check(N.lib().sdbg_scan_stats(self._h, C.byref(a), C.byref(b)), self._h)
        return a.value, b.value

    # ---- collectives (NCCL behind the C ABI; the host only has to ship the 128-byte id to every rank) ----
    @staticmethod
    def dist_unique_id():
        buf = (C.c_uint8 * 128)()
        N.check(N.lib().sdbg_dist_unique_id(buf))
        return bytes(buf)

    def dist_init(self, id128, rank, world):
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(id128))
        N.check(N.lib().sdbg_dist_init(self._h, buf, int(rank), int(world)), self._h)

    def dist_allreduce_i64(self, d_ptr, n):
        N.check(N.lib().sdbg_dist_allreduce_i64(self._h, C.c_void_p(int(d_ptr)), int(n)), self._h)

    def dist_allgather(self, d_send, d_recv, bytes_per_rank):
        N.check(N.lib().sdbg_dist_allgather(self._h, C.c_void_p(int(d_send)), C.c_void_p(int(d_recv)), int(bytes_per_rank)), self._h)

    def dist_groupby_merge(self, d_i64, d_f64, span, abs_bound):
        """Dense GROUP BY partials of all ranks -> global partials on every rank: one ncclAllReduce on this context's stream."""
        N.check(N.lib().sdbg_dist_groupby_merge(self._h, C.c_void_p(int(d_i64)), C.c_void_p(int(d_f64)), int(span), float(abs_bound)), self._h)

    def profile_read(self, kernel):
        """kernel: 'groupby' | 'topk' | 'merge' | 'count_sum' -> (total_ms, launches) since profile(True)."""
        kid = dict(groupby=0, topk=1, merge=2, count_sum=3)[kernel]
        ms, n = C.c_double(), C.c_uint64()
        N.check(N.lib().sdbg_profile_read(self._h, kid, C.byref(ms), C.byref(n)), self._h)
        return ms.value, n.value

    @property
    def launches(self):
        return int(N.lib().sdbg_launch_count(self._h))


class PostingsWriter:
    """Host mirror of irs PostingsWriterImpl (formats/posting/writer.hpp): builds a ".doc" stream."""

    def __init__(self, segment_docs, norms=None, has_wand=True, wand_b=0.75):
        self._h = C.c_void_p()
        self._norms = None if norms is None else np.ascontiguousarray(norms, dtype=np.uint32)
        N.check(N.lib().sdbg_writer_create(int(segment_docs), 1 if has_wand else 0, float(wand_b),
                                           _ptr(self._norms), C.byref(self._h)))

    def add_term(self, docs, freqs):
        docs = np.ascontiguousarray(docs, dtype=np.uint32)
        freqs = np.ascontiguousarray(freqs, dtype=np.uint32)
        N.check(N.lib().sdbg_writer_add_term(self._h, _ptr(docs), _ptr(freqs), len(docs)))

    def finish(self):
        """-> (doc_bytes uint8[n], term metas structured array)."""
        p, n, t, nt = C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_size_t()
        N.check(N.lib().sdbg_writer_finish(self._h, C.byref(p), C.byref(n), C.byref(t), C.byref(nt)))
        doc = np.zeros(n.value, np.uint8)
        if n.value:
            C.memmove(doc.ctypes.data, p.value, n.value)
        metas = np.zeros(nt.value, TERM_META_DTYPE)
        if nt.value:
            C.memmove(metas.ctypes.data, t.value, nt.value * TERM_META_DTYPE.itemsize)
        return doc, metas

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            N.lib().sdbg_writer_destroy(self._h)
            self._h = C.c_void_p()


def stage_parse_host(doc_bytes, metas, has_wand=True):
    """Host-only probe of the staging parser: the block table the kernels would read."""
    doc_bytes = np.ascontiguousarray(doc_bytes, dtype=np.uint8)
    metas = np.ascontiguousarray(metas, dtype=TERM_META_DTYPE)
    cap = int(sum((int(m) + 127) // 128 for m in metas["docs_count"])) + 1
    nblk = C.c_uint32()
    tbb = np.zeros(len(metas) + 1, np.uint32)
    last = np.zeros(cap, np.uint32)
    prev = np.zeros(cap, np.uint32)
    packed = np.zeros(cap, np.uint32)
    mf = np.zeros(cap, np.uint32)
    mn = np.zeros(cap, np.uint32)
    arena = C.c_uint64()
    N.check(N.lib().sdbg_debug_stage_host(_ptr(doc_bytes), len(doc_bytes), _ptr(metas), len(metas),
                                          1 if has_wand else 0, cap, C.byref(nblk), _ptr(tbb), _ptr(last),
                                          _ptr(prev), _ptr(packed), _ptr(mf), _ptr(mn), C.byref(arena)))
    n = nblk.value
    return dict(term_blk_begin=tbb, last_doc=last[:n], prev_last=prev[:n], packed=packed[:n],
                max_freq=mf[:n], max_norm=mn[:n], arena_bytes=arena.value)


class Segment:
    """One index segment resident in HBM: postings, norms, table columns."""

    def __init__(self, ctx, n_docs):
        self.ctx = ctx
        self.n_docs = int(n_docs)
        self._h = C.c_void_p()
        N.check(N.lib().sdbg_segment_create(ctx._h, self.n_docs, C.byref(self._h)), ctx._h)
        self.term_docs = None  # docs_count per term (filled by staging)
        self._keep = []        # host buffers that must outlive async copies

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            N.lib().sdbg_segment_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- staging ----
    def stage_postings(self, doc_bytes, metas, has_wand=True, wand_b=0.75):
        doc_bytes = np.ascontiguousarray(doc_bytes, dtype=np.uint8)
        metas = np.ascontiguousarray(metas, dtype=TERM_META_DTYPE)
        N.check(N.lib().sdbg_stage_postings(self._h, _ptr(doc_bytes), len(doc_bytes), _ptr(metas), len(metas),
                                            1 if has_wand else 0), self.ctx._h)
        N.check(N.lib().sdbg_segment_set_wand_b(self._h, float(wand_b)), self.ctx._h)   # pruning only for scorers with this b
        self.term_docs = metas["docs_count"].astype(np.uint64)

    def stage_norms(self, norm_bytes, byte_width):
        norm_bytes = np.ascontiguousarray(norm_bytes, dtype=np.uint8)
        rg = N.NormRg(byte_width, self.n_docs, 0)
        N.check(N.lib().sdbg_stage_norms(self._h, _ptr(norm_bytes), len(norm_bytes), C.byref(rg), 1), self.ctx._h)

    def stage_column(self, field, values, validity=None):
        """values: numpy array (int64/float64/int32) or a (host_ptr, dtype, rows) triple of pinned memory."""
        if isinstance(values, tuple):
            ptr, dtype, rows = values
            t = TYPES[np.dtype(dtype)]
            vp = C.c_void_p(int(ptr))
        else:
            values = np.ascontiguousarray(values)
            t, rows, vp = TYPES[values.dtype], len(values), _ptr(values)
            self._keep.append(values)
        vv = None
        if validity is not None:
            validity = np.ascontiguousarray(validity, dtype=np.uint64)
            self._keep.append(validity)
            vv = _ptr(validity)
        N.check(N.lib().sdbg_stage_column(self._h, int(field), t, vp, vv, int(rows)), self.ctx._h)

    def stage_column_for(self, field, packed):
        """Stage an int64 column from its frame-of-reference bit-packed form (pack_for): only the packed bytes cross PCIe,
        the values are unpacked on the GPU."""
        headers, words, rows = packed
        self._keep.append(packed)
        N.check(N.lib().sdbg_stage_column_for(self._h, int(field), _ptr(headers), _ptr(words), len(words), int(rows)), self.ctx._h)

    def stage_docs_mask(self, deleted_docs):
        """DocumentMask of the segment: doc ids that queries must neither score nor count (None / empty clears)."""
        d = np.ascontiguousarray(deleted_docs if deleted_docs is not None else [], dtype=np.uint32)
        N.check(N.lib().sdbg_stage_docs_mask(self._h, _ptr(d) if len(d) else None, len(d)), self.ctx._h)

    def stage_column_device(self, field, device_ptr, dtype, rows):
        N.check(N.lib().sdbg_stage_column_device(self._h, int(field), TYPES[np.dtype(dtype)],
                                                 C.c_void_p(int(device_ptr)), int(rows)), self.ctx._h)

    def column_device_ptr(self, field):
        p, r = C.c_void_p(), C.c_uint64()
        N.check(N.lib().sdbg_column_device_ptr(self._h, int(field), C.byref(p), C.byref(r)), self.ctx._h)
        return p.value, r.value

    def gather(self, field, docs, dtype):
        """Column values of the given hit docs (late materialisation, HitBatcher::MaterializeColumn): (values, valid)."""
        docs = np.ascontiguousarray(docs, dtype=np.uint32)
        out = np.zeros(len(docs), dtype)
        valid = np.zeros(len(docs), np.uint8)
        N.check(N.lib().sdbg_gather_column(self._h, int(field), _ptr(docs) if len(docs) else None, len(docs),
                                           _ptr(out) if len(docs) else None, _ptr(valid) if len(docs) else None), self.ctx._h)
        return out, valid.astype(bool)

    def column_to_host(self, field, host_ptr, rows):
        N.check(N.lib().sdbg_column_to_host(self._h, int(field), C.c_void_p(int(host_ptr)), int(rows)), self.ctx._h)

    def synth_corpus(self, doc0, t0, nt, threads=8, p_floor=0.0):
        """SURVEY §8d corpus shard: returns (docs_count per term, sum of doc lengths). p_floor > 0 gives every term at
        least that inclusion probability (a flat tail: an index far larger than L2)."""
        dc = np.zeros(nt, np.uint32)
        sdl = C.c_uint64()
        N.check(N.lib().sdbg_synth_corpus_ex(self._h, int(doc0), self.n_docs, int(t0), int(nt), int(threads), float(p_floor),
                                             _ptr(dc), C.byref(sdl)), self.ctx._h)
        self.term_docs = dc.astype(np.uint64)
        return dc, sdl.value

    def synth_column(self, field, stream, kind, row0, rows):
        N.check(N.lib().sdbg_synth_column(self._h, int(field), int(stream), int(kind), int(row0), int(rows)),
                self.ctx._h)

    def posting_stats(self):
        a, b, c, d = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        N.check(N.lib().sdbg_segment_posting_stats(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return dict(payload_bytes=a.value, table_bytes=b.value, n_blocks=c.value, n_postings=d.value)

    def term_bytes(self, n_terms):
        out = np.zeros(int(n_terms), np.uint64)
        N.check(N.lib().sdbg_segment_term_bytes(self._h, _ptr(out), int(n_terms)), self.ctx._h)
        return out

    def column_minmax(self, field):
        mn, mx = C.c_int64(), C.c_int64()
        N.check(N.lib().sdbg_column_minmax_i64(self._h, int(field), C.byref(mn), C.byref(mx)), self.ctx._h)
        return mn.value, mx.value

    # ---- probes ----
    def decode_score_term(self, term, c0, norm_const, norm_length):
        n = int(self.term_docs[term])
        docs = np.zeros(max(n, 1), np.uint32)
        freqs = np.zeros(max(n, 1), np.uint32)
        scores = np.zeros(max(n, 1), np.float32)
        N.check(N.lib().sdbg_decode_score_term(self._h, int(term), float(c0), float(norm_const), float(norm_length),
                                               _ptr(docs), _ptr(freqs), _ptr(scores)), self.ctx._h)
        return docs[:n], freqs[:n], scores[:n]

    def filter_bitmap(self, preds, rows=None):
        rows = self.n_docs if rows is None else rows
        mask = np.zeros((rows + 63) // 64, np.uint64)
        N.check(N.lib().sdbg_filter_bitmap(self._h, _pred_array(preds), len(preds), _ptr(mask)), self.ctx._h)
        return mask


def _seg_array(segs):
    arr = (C.c_void_p * len(segs))()
    for i, s in enumerate(segs):
        arr[i] = s._h
    return arr


class BM25:
    """irs::BM25 (search/bm25.hpp:58): k, b and the statistics -> BM25Stats step (bm25.cpp:279-310)."""

    def __init__(self, k=1.2, b=0.75):
        self.k, self.b = float(k), float(b)

    def collect(self, docs_with_field, total_term_freq, docs_with_term, term=0, boost=1.0):
        t = N.BM25Term()
        N.check(N.lib().sdbg_bm25_collect(int(docs_with_field), int(total_term_freq), int(docs_with_term),
                                          self.k, self.b, C.byref(t)))
        t.term = int(term)
        t.boost = float(boost)
        return t

    def num(self, term):
        """c0 = boost*(k+1)*idf in fp32 (bm25.cpp:224)."""
        return np.float32(np.float32(np.float32(term.boost) * np.float32(self.k + np.float32(1))) * np.float32(term.idf))


class TFIDF:
    """irs::TFIDF (search/tfidf.cpp): sqrt(freq) * boost * idf, divided by sqrt(doc length) when `normalize`; idf from
    TFIDF::collect (:149-150). Runs through the same scan entry points: k = -1 is the ABI's reserved selector for it
    (sdbg_tfidf_topk_batch forwards exactly that), b != 0 means normalised. Always exhaustive."""

    def __init__(self, normalize=False):
        self.normalize = bool(normalize)
        self.k, self.b = -1.0, (1.0 if normalize else 0.0)

    def collect(self, docs_with_field, total_term_freq, docs_with_term, term=0, boost=1.0):
        t = N.BM25Term()
        N.check(N.lib().sdbg_tfidf_collect(int(docs_with_field), int(docs_with_term), C.byref(t)))
        t.term = int(term)
        t.boost = float(boost)
        return t

    def num(self, term):
        return np.float32(np.float32(term.boost) * np.float32(term.idf))


class IndexReader:
    """The segments of one snapshot on one GPU plus corpus-wide field statistics
    (FieldCollector / TermCollector sums over all segments, search/collectors.cpp:30-52)."""

    def __init__(self, segments, docs_with_field, total_term_freq, docs_with_term):
        self.segments = list(segments)
        self.docs_with_field = int(docs_with_field)
        self.total_term_freq = int(total_term_freq)
        self.docs_with_term = np.asarray(docs_with_term, dtype=np.uint64)  # per term id, global

    def stats(self, scorer, term, boost=1.0):
        return scorer.collect(self.docs_with_field, self.total_term_freq, int(self.docs_with_term[term]), term, boost)


def ExecuteTopKBatch(reader, queries, kind, scorer, k, filt=None, threshold=FLT_MIN):
    """Batch of ExecuteTopK calls (doc_collector.hpp:88-136). queries: list of term-id lists.
    Returns (hits [Q, k] structured, n_out [Q], total_matches [Q])."""
    nq = len(queries)
    flat = [reader.stats(scorer, t) for q in queries for t in q]
    terms = (N.BM25Term * max(len(flat), 1))()
    for i, t in enumerate(flat):
        terms[i] = t
    off = np.zeros(nq + 1, np.uint32)
    off[1:] = np.cumsum([len(q) for q in queries])
    hits = np.zeros((nq, k), HIT_DTYPE)
    n_out = np.zeros(nq, np.uint32)
    total = np.zeros(nq, np.uint64)
    ctx = reader.segments[0].ctx
    fp = C.byref(filt) if filt is not None else None
    N.check(N.lib().sdbg_bm25_topk_batch(_seg_array(reader.segments), len(reader.segments), int(kind), terms,
                                         _ptr(off), nq, scorer.k, scorer.b, fp, int(k), float(threshold), _ptr(hits),
                                         _ptr(n_out), _ptr(total)), ctx._h)
    return hits, n_out, total


FOR_BLOCK_DTYPE = np.dtype([("base", "<i8"), ("bits", "<u4"), ("off8", "<u4")])


def pack_for(values, out_words=None):
    """Host-side writer of the bit-packed column format (sdbg_pack_for): returns (headers, words, rows). `out_words` may be a
    preallocated (e.g. pinned) uint64 array."""
    values = np.ascontiguousarray(values, dtype=np.int64)
    rows = len(values)
    headers = np.zeros((rows + 2047) // 2048, FOR_BLOCK_DTYPE)
    n = C.c_uint64(0)
    words = out_words if out_words is not None else np.zeros(rows + 1, np.uint64)      # never larger than the raw column + slack
    rc = N.lib().sdbg_pack_for(_ptr(values), rows, _ptr(headers), _ptr(words), len(words), C.byref(n))
    N.check(rc)
    return headers, words[:n.value], rows


def StreamScoredDocs(reader, seg_idx, query, kind, scorer, filt=None, doc_min=1, doc_max=None):
    """The search scan's streaming mode (duckdb_search_full_scan.cpp:2370 RunStreamingScan over
    DocIterator::EmitScoredDocs): every match of `query` in docs [doc_min, doc_max) of segment `seg_idx` with its score,
    ascending by doc id. Returns (docs u32, scores f32)."""
    seg = reader.segments[seg_idx]
    terms = (N.BM25Term * len(query))(*[reader.stats(scorer, t) for t in query])
    fp = C.byref(filt) if filt is not None else None
    hi = int(doc_max) if doc_max is not None else 0xFFFFFFFF
    n = C.c_uint64(0)
    cap = 0
    docs = scores = None
    for _ in range(2):   # count-only call first, then one with exactly the room needed
        rc = N.lib().sdbg_bm25_scan(seg._h, int(kind), terms, len(query), scorer.k, scorer.b, fp, int(doc_min), hi,
                                    _ptr(docs) if docs is not None else None, _ptr(scores) if scores is not None else None,
                                    cap, C.byref(n))
        if rc == -6 and n.value > cap:
            cap = n.value
            docs, scores = np.zeros(cap, np.uint32), np.zeros(cap, np.float32)
            continue
        N.check(rc, seg.ctx._h)
        break
    if docs is None:
        return np.zeros(0, np.uint32), np.zeros(0, np.float32)
    return docs[:n.value], scores[:n.value]


def _flatten_queries(reader, queries, scorer):
    flat = [reader.stats(scorer, t) for q in queries for t in q]
    terms = (N.BM25Term * max(len(flat), 1))()
    for i, t in enumerate(flat):
        terms[i] = t
    off = np.zeros(len(queries) + 1, np.uint32)
    off[1:] = np.cumsum([len(q) for q in queries])
    return terms, off


class PreparedBatch:
    """Query descriptors marshalled once (terms + statistics), reusable across steps."""

    def __init__(self, reader, queries, kind, scorer, k, filt=None, threshold=FLT_MIN):
        self.reader, self.kind, self.scorer, self.k, self.filt, self.threshold = reader, int(kind), scorer, int(k), filt, float(threshold)
        self.nq = len(queries)
        self.terms, self.off = _flatten_queries(reader, queries, scorer)
        self.hits = np.zeros((self.nq, self.k), HIT_DTYPE)
        self.n_out = np.zeros(self.nq, np.uint32)
        self.total = np.zeros(self.nq, np.uint64)

    def run_host(self):
        """Full API call: host descriptors in, host hits out."""
        r = self.reader
        fp = C.byref(self.filt) if self.filt is not None else None
        N.check(N.lib().sdbg_bm25_topk_batch(_seg_array(r.segments), len(r.segments), self.kind, self.terms,
                                             _ptr(self.off), self.nq, self.scorer.k, self.scorer.b, fp, self.k, self.threshold,
                                             _ptr(self.hits), _ptr(self.n_out), _ptr(self.total)), r.segments[0].ctx._h)
        return self.hits, self.n_out, self.total

    def run_dist(self, to_host=True):
        """Distributed top-k (sdbg_dist_bm25_topk_batch): local scan, one all-gather, local selection -- all enqueued by
        the library on its stream. to_host=False leaves the merged keys in HBM and returns without waiting."""
        r = self.reader
        fp = C.byref(self.filt) if self.filt is not None else None
        if to_host:
            hits = np.zeros((self.nq, self.k), HIT_DTYPE)
            n_out = np.zeros(self.nq, np.uint32)
            hp, npp = _ptr(hits), _ptr(n_out)
        else:
            hits = n_out = None
            hp = npp = None
        N.check(N.lib().sdbg_dist_bm25_topk_batch(_seg_array(r.segments), len(r.segments), self.kind, self.terms, _ptr(self.off), self.nq,
                                                  self.scorer.k, self.scorer.b, fp, self.k, self.threshold, hp, npp), r.segments[0].ctx._h)
        return hits, n_out

    def run_device(self, rank, d_keys_ptr, d_totals_ptr=None):
        """Results stay in HBM as sortable keys (for the multi-GPU gather + merge)."""
        r = self.reader
        fp = C.byref(self.filt) if self.filt is not None else None
        N.check(N.lib().sdbg_bm25_topk_batch_device(_seg_array(r.segments), len(r.segments), self.kind, self.terms,
                                                    _ptr(self.off), self.nq, self.scorer.k, self.scorer.b, fp, self.k, self.threshold,
                                                    int(rank), C.c_void_p(int(d_keys_ptr)),
                                                    C.c_void_p(int(d_totals_ptr)) if d_totals_ptr else None),
                r.segments[0].ctx._h)


def merge_gathered(ctx, d_keys_all_ptr, n_ranks, nq, k, to_host=True):
    """Global top-k from the keys every rank contributed ([rank][query][k] u64 in HBM)."""
    if not to_host:
        N.check(N.lib().sdbg_topk_merge_gathered(ctx._h, C.c_void_p(int(d_keys_all_ptr)), int(n_ranks), int(nq), int(k),
                                                 None, None), ctx._h)
        return None, None
    hits = np.zeros((nq, k), HIT_DTYPE)
    n_out = np.zeros(nq, np.uint32)
    N.check(N.lib().sdbg_topk_merge_gathered(ctx._h, C.c_void_p(int(d_keys_all_ptr)), int(n_ranks), int(nq), int(k),
                                             _ptr(hits), _ptr(n_out)), ctx._h)
    return hits, n_out


def ExecuteTopK(reader, query_terms, kind, scorer, k, filt=None, threshold=FLT_MIN):
    """irs::ExecuteTopK for one query: hits sorted by (score desc, seg asc, doc asc), total matches."""
    hits, n_out, total = ExecuteTopKBatch(reader, [list(query_terms)], kind, scorer, k, filt, threshold)
    return hits[0, :n_out[0]].copy(), int(total[0])


class IResearchScan:
    """The ColScan / count shapes of the `iresearch_scan` table function
    (server/connector/duckdb_search_full_scan.hpp:56-77) with the aggregate above it pushed into the scan."""

    def __init__(self, segments):
        self.segments = list(segments)
        self.ctx = self.segments[0].ctx

    def count_sum(self, preds, sum_field=None):
        cnt = C.c_uint64()
        s128 = (C.c_int64 * 2)()
        sf = C.c_double()
        N.check(N.lib().sdbg_filter_count_sum(_seg_array(self.segments), len(self.segments), _pred_array(preds),
                                              len(preds), NO_FIELD if sum_field is None else int(sum_field),
                                              C.byref(cnt), s128, C.byref(sf)), self.ctx._h)
        si = (int(s128[1]) << 64) | (int(s128[0]) & 0xFFFFFFFFFFFFFFFF)
        return cnt.value, si, sf.value

    def prepare_count_sum(self, preds, sum_field=None):
        """The same call with its arguments marshalled once (a point query is ~10 us on the device side; building the
        ctypes arrays per call costs about as much). Returns a callable -> (count, sum_int, sum_f64)."""
        segs, n_segs, pa, n_preds = _seg_array(self.segments), len(self.segments), _pred_array(preds), len(preds)
        field = NO_FIELD if sum_field is None else int(sum_field)
        cnt, s128, sf = C.c_uint64(), (C.c_int64 * 2)(), C.c_double()
        pc, psf = C.byref(cnt), C.byref(sf)
        fn, h = N.lib().sdbg_filter_count_sum, self.ctx._h

        def run():
            rc = fn(segs, n_segs, pa, n_preds, field, pc, s128, psf)
            if rc:
                N.check(rc, h)
            return cnt.value, (int(s128[1]) << 64) | (int(s128[0]) & 0xFFFFFFFFFFFFFFFF), sf.value
        return run

    def groupby(self, preds, key_field, sum_int_field=None, avg_f64_field=None, cap=None, n_groups_hint=0):
        cap = int(cap if cap is not None else max(n_groups_hint, 1 << 20))
        out = np.zeros(cap, GROUP_DTYPE)
        n = C.c_uint64()
        N.check(N.lib().sdbg_filter_groupby(_seg_array(self.segments), len(self.segments), _pred_array(preds),
                                            len(preds), int(key_field), int(n_groups_hint),
                                            NO_FIELD if sum_int_field is None else int(sum_int_field),
                                            NO_FIELD if avg_f64_field is None else int(avg_f64_field),
                                            _ptr(out), cap, C.byref(n)), self.ctx._h)
        return out[:n.value]

    def groupby_partial(self, preds, key_field, key_min, key_span, sum_int_field, avg_f64_field, d_i64_ptr, d_f64_ptr):
        N.check(N.lib().sdbg_filter_groupby_partial(_seg_array(self.segments), len(self.segments), _pred_array(preds),
                                                    len(preds), int(key_field), int(key_min), int(key_span),
                                                    NO_FIELD if sum_int_field is None else int(sum_int_field),
                                                    NO_FIELD if avg_f64_field is None else int(avg_f64_field),
                                                    C.c_void_p(int(d_i64_ptr)), C.c_void_p(int(d_f64_ptr))), self.ctx._h)

    def groupby_finalize(self, key_min, key_span, d_i64_ptr, d_f64_ptr, cap, out=None):
        if out is None:
            out = np.empty(int(cap), GROUP_DTYPE)
        n = C.c_uint64()
        N.check(N.lib().sdbg_groupby_finalize(self.ctx._h, int(key_min), int(key_span), C.c_void_p(int(d_i64_ptr)),
                                              C.c_void_p(int(d_f64_ptr)), _ptr(out), int(cap), C.byref(n)), self.ctx._h)
        return out[:n.value]


def sum_i128(rows):
    """Python ints of the 128-bit SUM(int) column of a group result."""
    return [(int(h) << 64) | (int(l) & 0xFFFFFFFFFFFFFFFF) for l, h in zip(rows["sum_lo"], rows["sum_hi"])]
