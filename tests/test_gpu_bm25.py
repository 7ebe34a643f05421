"""GPU parity: posting decode + BM25 + top-k through the C ABI vs the CPU oracle (bit-exact doc ids,
freqs, fp32 scores and top-k order), reference goldens, edge cases, and size-independent properties."""
import json
import os

import numpy as np
import pytest

import orc
import serenedb_b200 as sdb
from gpu_util import assert_hits_equal, ctx, metas_of, oracle_terms, to_gpu

pytestmark = pytest.mark.gpu
G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bm25_goldens.json")))


@pytest.fixture(scope="module")
def corpus():
    n = 300_000
    terms = [0, 1, 2, 3, 9, 30, 60, 200, 255]
    oseg, dl, lists = orc.synth_segment(n, terms)
    nn = orc.synth_column(2, 1, 1, n).astype(np.int32)  # n:int32 = h % 1e6 over docs 1..n
    oseg.add_column(9, nn)
    gseg = to_gpu(oseg, columns={9: (nn, None)})
    reader = sdb.IndexReader([gseg], n, int(dl.sum()), [len(d) for d, _ in lists])
    return dict(oseg=oseg, gseg=gseg, dl=dl, lists=lists, n=n, reader=reader, nn=nn)


def test_decode_and_score_every_posting(corpus):
    """Row 1/2/6/7/8 of SURVEY §8a: decode docs+freqs, gather norms, score -- exhaustive, bit-exact."""
    scorer = sdb.BM25()
    for t, (docs, freqs) in enumerate(corpus["lists"]):
        st = corpus["reader"].stats(scorer, t)
        c0 = scorer.num(st)
        d, f, s = corpus["gseg"].decode_score_term(t, c0, st.norm_const, st.norm_length)
        assert np.array_equal(d, docs) and np.array_equal(f, freqs)
        ost = orc.BM25Stats(st.idf, st.norm_const, st.norm_length)
        exp = orc.bm25_score(freqs, corpus["dl"][docs - 1], ost)
        assert np.array_equal(s.view(np.uint32), exp.view(np.uint32))


def test_decode_all_encodings():
    """all-same, bitset, bit-packed, raw, (delta) streamvbyte tails, single-doc terms, no norms."""
    rng = np.random.default_rng(3)
    n = 70000
    oseg = orc.Segment(n, has_wand=False)
    lists = []
    shapes = [np.arange(1, n + 1, dtype=np.uint32),                       # gap 1 everywhere -> all_same_08
              np.arange(5, n + 1, 300, dtype=np.uint32),                  # all_same_16
              np.array([69999], np.uint32),                               # single doc (inline in term meta)
              np.array([7, 69000], np.uint32),                            # 2-doc tail
              np.sort(rng.choice(np.arange(1, 3000), 1500, replace=False)).astype(np.uint32),   # bitset blocks
              np.sort(rng.choice(np.arange(1, n + 1), 128, replace=False)).astype(np.uint32),   # exactly one block
              np.sort(rng.choice(np.arange(1, n + 1), 129, replace=False)).astype(np.uint32),   # block + 1-doc tail
              np.sort(rng.choice(np.arange(1, n + 1), 1000, replace=False)).astype(np.uint32),  # bitpack + svb tail
              np.sort(rng.choice(np.arange(1, n + 1), 40000, replace=False)).astype(np.uint32)]
    for docs in shapes:
        freqs = rng.integers(1, 5, size=len(docs)).astype(np.uint32)
        if len(docs) == 1000:
            freqs[-40:] = rng.integers(1, 70000, size=40)  # wide freq tail -> svb / raw
        if len(docs) == 129:
            freqs[:] = 3                                   # all_same freq block
        oseg.add_term(docs, freqs)
        lists.append((docs, freqs))
    g = to_gpu(oseg, has_wand=False)
    for t, (docs, freqs) in enumerate(lists):
        d, f, s = g.decode_score_term(t, 2.0, 0.3, 0.01)
        assert np.array_equal(d, docs), t
        assert np.array_equal(f, freqs), t
        exp = orc.bm25_score(freqs, None, orc.BM25Stats(0, 0.3, 0.01))  # no norm column => norm = 1
        num = np.float32(2.0)
        # recompute with explicit c0 through the oracle helper
        out = np.zeros(len(freqs), np.float32)
        orc.lib().orc_bm25_score(orc.ptr(freqs), None, len(freqs), num, np.float32(0.3), np.float32(0.01), orc.ptr(out))
        assert np.array_equal(s.view(np.uint32), out.view(np.uint32)), t


QUERIES = [("OR", [3], 10), ("OR", [0, 4], 100), ("OR", [2, 5, 6], 1000), ("OR", [8], 1000), ("OR", [0, 1], 1000),
           ("AND", [0, 1, 2], 50), ("AND", [0, 1, 2, 3, 4], 1000), ("AND", [0, 8], 1000), ("OR", [0, 1, 2, 3, 4, 5, 6, 7], 500),
           ("OR", [7], 5000), ("OR", [4], 1)]


@pytest.mark.parametrize("kind,tis,k", QUERIES)
def test_topk_matches_oracle(corpus, kind, tis, k):
    scorer = sdb.BM25()
    hits, total = sdb.ExecuteTopK(corpus["reader"], tis, sdb.AND if kind == "AND" else sdb.OR, scorer, k)
    oh, ototal, _ = orc.bm25_topk([corpus["oseg"]], kind, oracle_terms(corpus["reader"], scorer, tis), k, mode=1)
    assert_hits_equal(hits, oh)
    assert total == ototal


def test_topk_batch_equals_single_queries(corpus):
    scorer = sdb.BM25()
    rng = np.random.default_rng(1)
    queries = [list(rng.choice(9, size=2, replace=False)) for _ in range(64)]
    hits, n_out, total = sdb.ExecuteTopKBatch(corpus["reader"], queries, sdb.OR, scorer, 100)
    for qi in (0, 7, 31, 63):
        oh, ototal, _ = orc.bm25_topk([corpus["oseg"]], "OR", oracle_terms(corpus["reader"], scorer, queries[qi]), 100, mode=1)
        assert_hits_equal(hits[qi, :n_out[qi]], oh)
        assert total[qi] == ototal


def test_hybrid_filter(corpus):
    """5-term AND + int range filter (config 4 shape): TableFilterDocIterator semantics."""
    scorer = sdb.BM25()
    tis = [0, 1, 2, 3, 4]
    hits, total = sdb.ExecuteTopK(corpus["reader"], tis, sdb.AND, scorer, 1000, filt=sdb.pred(9, "BETWEEN", 250000, 749999))
    oh, ototal, _ = orc.bm25_topk([corpus["oseg"]], "AND", oracle_terms(corpus["reader"], scorer, tis), 1000,
                                  filt=orc.make_pred(9, "BETWEEN", 250000, 749999), mode=1)
    assert_hits_equal(hits, oh)
    assert total == ototal and total > 0
    assert np.all((corpus["nn"][hits["doc"] - 1] >= 250000) & (corpus["nn"][hits["doc"] - 1] <= 749999))


def test_threshold_seed_and_small_results(corpus):
    scorer = sdb.BM25()
    full, _ = sdb.ExecuteTopK(corpus["reader"], [6], sdb.OR, scorer, 50)
    thr = float(full["score"][9])
    hits, _ = sdb.ExecuteTopK(corpus["reader"], [6], sdb.OR, scorer, 50, threshold=thr)
    assert np.all(hits["score"] > np.float32(thr))          # strict '>' like iterators.hpp:206-208
    assert np.array_equal(hits["doc"], full["doc"][full["score"] > np.float32(thr)])
    # k larger than the number of matches
    n_match = len(corpus["lists"][8][0])
    hits, total = sdb.ExecuteTopK(corpus["reader"], [8], sdb.OR, scorer, 8192)
    assert len(hits) == min(8192, n_match) and total == n_match
    assert np.all(np.diff(hits["score"]) <= 0)


def test_multisegment_global_stats(corpus):
    """Two segments by doc range with corpus-wide statistics == one segment (collectors.cpp:36-52)."""
    n, half = corpus["n"], corpus["n"] // 2
    terms = [0, 1, 2, 3, 9, 30, 60, 200, 255]
    oA, dlA, lA = orc.synth_segment(half, terms, doc0=0)
    oB, dlB, lB = orc.synth_segment(n - half, terms, doc0=half)
    gA, gB = to_gpu(oA), to_gpu(oB)
    dwt = [len(a[0]) + len(b[0]) for a, b in zip(lA, lB)]
    reader2 = sdb.IndexReader([gA, gB], n, int(dlA.sum() + dlB.sum()), dwt)
    scorer = sdb.BM25()
    for kind, tis, k in [(sdb.OR, [0, 4], 100), (sdb.AND, [0, 1, 2], 200), (sdb.OR, [7], 300)]:
        h1, t1 = sdb.ExecuteTopK(corpus["reader"], tis, kind, scorer, k)
        h2, t2 = sdb.ExecuteTopK(reader2, tis, kind, scorer, k)
        glob = np.where(h2["seg"] == 0, h2["doc"], h2["doc"] + half)
        assert np.array_equal(glob, h1["doc"]) and np.array_equal(h2["score"], h1["score"]) and t1 == t2


def test_reference_goldens_on_gpu():
    """ranking.test 1.9693236 and the WAND + table-filter top-2 answers, through the GPU path."""
    g = G["wand_table_filter"]
    rows = g["rows"]
    toks = [r["body"].split() for r in rows]
    n = len(rows)
    oseg = orc.Segment(n, has_wand=True)
    oseg.set_norms([len(t) for t in toks])
    vocab = {}
    for w in ("term", "fill", "rare"):
        vocab[w] = oseg.add_term([i + 1 for i, t in enumerate(toks) if w in t], [t.count(w) for t in toks if w in t])
    nn = np.array([r["n"] for r in rows], np.int32)
    gseg = to_gpu(oseg, columns={7: (nn, None)})
    reader = sdb.IndexReader([gseg], n, sum(len(t) for t in toks), [oseg.term_meta(t).docs_count for t in range(3)])
    scorer = sdb.BM25(1.2, 0.75)
    for case in ("case1", "case2"):
        hits, _ = sdb.ExecuteTopK(reader, [vocab[w] for w in g[case]["terms"]], sdb.OR, scorer, g["k"], filt=sdb.pred(7, "GT", 20))
        assert [rows[d - 1]["n"] for d in hits["doc"]] == g[case]["expected_n"]
    # ranking.test: 4 docs, no norm column, 'alien' in one doc
    r = G["ranking_alien"]
    oseg = orc.Segment(4, has_wand=False)
    oseg.add_term([3], [1])
    gseg = to_gpu(oseg, has_wand=False)
    reader = sdb.IndexReader([gseg], r["docs_with_field"], r["total_term_freq"], [1])
    hits, total = sdb.ExecuteTopK(reader, [0], sdb.OR, scorer, 10)
    assert total == 1 and hits["doc"][0] == 3
    assert np.format_float_positional(hits["score"][0], unique=True, trim="-") == r["expected"]


def test_sequential_order_goldens_on_gpu():
    """bm25_test.cpp:163,214 (rank order of a ByRange disjunction under BM25 with norms) through the GPU path,
    at every pruning level, bit-exact against the oracle that test_oracle_goldens pins to the same golden."""
    g = G["sequential_order"]
    n = len(g["docs"])
    dl = np.array([len(d["field"]) for d in g["docs"]], np.uint32)
    oseg = orc.Segment(n, has_wand=True)
    oseg.set_norms(dl)
    for t in range(10):
        docs = [i + 1 for i, d in enumerate(g["docs"]) if str(t) in d["field"]]
        oseg.add_term(np.array(docs, np.uint32), np.array([g["docs"][i - 1]["field"].count(str(t)) for i in docs], np.uint32))
    gseg = to_gpu(oseg)
    reader = sdb.IndexReader([gseg], n, int(dl.sum()), [oseg.term_meta(t).docs_count for t in range(10)])
    scorer = sdb.BM25(g["k"], g["b"])
    try:
        for level in (0, 1, 2):
            ctx().set_wand(level)
            for c in g["cases"]:
                tis = [int(t) for t in c["terms"]]
                hits, total = sdb.ExecuteTopK(reader, tis, sdb.OR, scorer, 8)
                assert [g["docs"][d - 1]["seq"] for d in hits["doc"]] == c["expected_seq_order"], (level, c["range"])
                oh, _, _ = orc.bm25_topk([oseg], "OR", oracle_terms(reader, scorer, tis), 8, k1=g["k"], b=g["b"], mode=1)
                assert_hits_equal(hits, oh)
    finally:
        ctx().set_wand(0)


def test_sequential_order_goldens_without_norm_column_on_gpu():
    """bm25_test.cpp:506,908,958,1006,1054 (Bm25TestCase.test_query: the field has no Norm feature, every doc scores with
    norm = 1, bm25.cpp:353-360) through the GPU path at every pruning level, bit-exact against the pinned oracle."""
    g, t = G["sequential_order"], G["sequential_order_no_norms"]
    n = len(g["docs"])
    oseg = orc.Segment(n, has_wand=False)
    for ti in range(10):
        docs = [i + 1 for i, d in enumerate(g["docs"]) if str(ti) in d["field"]]
        oseg.add_term(np.array(docs, np.uint32), np.array([g["docs"][i - 1]["field"].count(str(ti)) for i in docs], np.uint32))
    gseg = to_gpu(oseg, has_wand=False)
    reader = sdb.IndexReader([gseg], n, 52, [oseg.term_meta(ti).docs_count for ti in range(10)])
    scorer = sdb.BM25(t["k"], t["b"])
    try:
        for level in (0, 2):
            ctx().set_wand(level)
            for c in t["cases"]:
                tis = [int(x) for x in c["terms"]]
                hits, total = sdb.ExecuteTopK(reader, tis, sdb.OR, scorer, 8)
                assert [g["docs"][d - 1]["seq"] for d in hits["doc"]] == c["expected_seq_order"], (level, c["range"])
                oh, _, _ = orc.bm25_topk([oseg], "OR", oracle_terms(reader, scorer, tis), 8, k1=t["k"], b=t["b"], mode=1)
                assert_hits_equal(hits, oh)
    finally:
        ctx().set_wand(0)


def test_tfidf_sequential_order_goldens_on_gpu():
    """tfidf_test.cpp:531,934,984,1032,1080 (rank orders under TFIDF without norms) through the GPU path, bit-exact against the
    oracle that test_oracle_goldens pins to the same goldens."""
    g, t = G["sequential_order"], G["tfidf_sequential_order"]
    n = len(g["docs"])
    dl = np.array([len(d["field"]) for d in g["docs"]], np.uint32)
    oseg = orc.Segment(n, has_wand=True)
    oseg.set_norms(dl)
    for ti in range(10):
        docs = [i + 1 for i, d in enumerate(g["docs"]) if str(ti) in d["field"]]
        oseg.add_term(np.array(docs, np.uint32), np.array([g["docs"][i - 1]["field"].count(str(ti)) for i in docs], np.uint32))
    gseg = to_gpu(oseg)
    reader = sdb.IndexReader([gseg], n, int(dl.sum()), [oseg.term_meta(ti).docs_count for ti in range(10)])
    scorer = sdb.TFIDF(normalize=t["normalize"])
    for c in t["cases"]:
        tis = [int(x) for x in c["terms"]]
        hits, total = sdb.ExecuteTopK(reader, tis, sdb.OR, scorer, 8)
        assert [g["docs"][d - 1]["seq"] for d in hits["doc"]] == c["expected_seq_order"], c["range"]
        assert total == len(c["expected_seq_order"])
        terms = []
        for ti in tis:
            x = orc.BM25Term()
            x.idf = orc.tfidf_idf(n, int(reader.docs_with_term[ti]))
            x.norm_const, x.norm_length, x.boost, x.term = 0.0, 0.0, 1.0, ti
            terms.append(x)
        oh, _, _ = orc.bm25_topk([oseg], "OR", terms, 8, k1=-1.0, b=0.0, mode=1)
        assert_hits_equal(hits, oh)


def test_collector_worst_case_increasing_scores():
    """Scores increasing with doc id defeat every threshold: the candidate buffer must keep compacting."""
    n = 200_000
    oseg = orc.Segment(n, has_wand=True)
    dl = np.arange(n, 0, -1, dtype=np.uint32) // 4 + 1   # shorter docs later => higher scores later
    oseg.set_norms(dl)
    docs = np.arange(1, n + 1, dtype=np.uint32)
    oseg.add_term(docs, np.ones(n, np.uint32))
    gseg = to_gpu(oseg)
    reader = sdb.IndexReader([gseg], n, int(dl.sum()), [n])
    scorer = sdb.BM25()
    hits, total = sdb.ExecuteTopK(reader, [0], sdb.OR, scorer, 1000)
    oh, ototal, _ = orc.bm25_topk([oseg], "OR", oracle_terms(reader, scorer, [0]), 1000, mode=1)
    assert_hits_equal(hits, oh)
    assert total == ototal == n


def test_property_checks_at_scale():
    """Size-independent properties on a 4M-doc shard built by the product's own corpus builder:
    sortedness, idempotence, every hit really matches, AND subset of OR."""
    n = 4_000_000
    g = sdb.Segment(ctx(), n)
    dc, sum_dl = g.synth_corpus(0, 0, 16, threads=8)
    reader = sdb.IndexReader([g], n, sum_dl, dc)
    scorer = sdb.BM25()
    h_or, t_or = sdb.ExecuteTopK(reader, [3, 12], sdb.OR, scorer, 1000)
    h_or2, _ = sdb.ExecuteTopK(reader, [3, 12], sdb.OR, scorer, 1000)
    assert np.array_equal(h_or, h_or2)
    assert len(h_or) == 1000 and np.all(np.diff(h_or["score"]) <= 0)
    ties = np.diff(h_or["score"]) == 0
    assert np.all(np.diff(h_or["doc"].astype(np.int64))[ties] > 0)   # canonical tie order: doc asc
    h_and, t_and = sdb.ExecuteTopK(reader, [3, 12], sdb.AND, scorer, 1000)
    assert t_and <= min(dc[3], dc[12]) and t_or == dc[3] + dc[12] - t_and   # inclusion-exclusion
    d3, _, s3 = g.decode_score_term(3, scorer.num(reader.stats(scorer, 3)), reader.stats(scorer, 3).norm_const, reader.stats(scorer, 3).norm_length)
    d12, _, s12 = g.decode_score_term(12, scorer.num(reader.stats(scorer, 12)), reader.stats(scorer, 12).norm_const, reader.stats(scorer, 12).norm_length)
    assert np.all(np.isin(h_and["doc"], d3)) and np.all(np.isin(h_and["doc"], d12))
    both = dict(zip(d3.tolist(), s3.tolist()))
    exp = np.array([np.float32(both[d]) for d in h_and["doc"].tolist()], np.float32)
    s12m = dict(zip(d12.tolist(), s12.tolist()))
    exp2 = np.array([np.float32(s12m[d]) for d in h_and["doc"].tolist()], np.float32)
    lo_first = dc[12] <= dc[3]   # sum order: ascending docs_count
    tot = (exp2 + exp) if lo_first else (exp + exp2)
    assert np.array_equal(h_and["score"], tot)


@pytest.mark.parametrize("k1,b", [(1.2, 0.0), (0.0, 0.75), (2.0, 1.0)])
@pytest.mark.parametrize("kind,tis,k", [("OR", [3], 100), ("OR", [0, 4], 1000), ("OR", [2, 5, 6], 300), ("AND", [0, 1, 2], 50)])
def test_scorer_forms(corpus, k1, b, kind, tis, k):
    """BM25::PrepareScorer's three forms (bm25.cpp:312-365): b == 0 -> Bm15 (no norms, its own operation
    order), k == 0 -> Bm1 (all scores 0: no hit beats the FLT_MIN seed, matches still counted), b == 1 -> the
    BM25 form (BM11). Bit-exact against the oracle with block-max pruning left on (it must disable itself for
    the non-BM25 forms: the staged pairs are BM25's)."""
    scorer = sdb.BM25(k1, b)
    # pruning stays switched on only where it has to disable itself; for (2.0, 1.0) the corpus' block-max pairs
    # (written for k = 1.2, b = 0.75) would not be valid bounds -- the reference only uses wand data of an
    # index-time scorer that equals the query's (Scorer::equals), which is the caller's side of has_wand
    ctx().set_wand(1 if (b == 0.0 or k1 == 0.0) else 0)
    try:
        hits, total = sdb.ExecuteTopK(corpus["reader"], tis, sdb.AND if kind == "AND" else sdb.OR, scorer, k)
    finally:
        ctx().set_wand(0)
    oh, ototal, _ = orc.bm25_topk([corpus["oseg"]], kind, oracle_terms(corpus["reader"], scorer, tis), k, k1=k1, b=b, mode=1)
    assert_hits_equal(hits, oh)
    if k1 == 0.0:
        assert len(hits) == 0
    assert total == ototal


def test_docs_mask(corpus):
    """DocumentMask: deleted docs are neither scored nor counted (SegmentReaderImpl::mask). OR, AND, hybrid
    filter and pruning on, bit-exact against the oracle; clearing the mask restores the full result."""
    n = corpus["n"]
    rng = np.random.default_rng(41)
    deleted = np.unique(rng.integers(1, n + 1, size=n // 4)).astype(np.uint32)
    scorer = sdb.BM25()
    corpus["oseg"].set_docs_mask(deleted)
    corpus["gseg"].stage_docs_mask(deleted)
    try:
        for wand in (0, 1):
            ctx().set_wand(wand)
            for kind, tis, k in (("OR", [0, 4], 1000), ("OR", [3], 100), ("AND", [0, 1, 2], 50), ("OR", [2, 5, 6], 300)):
                hits, total = sdb.ExecuteTopK(corpus["reader"], tis, sdb.AND if kind == "AND" else sdb.OR, scorer, k)
                oh, ototal, _ = orc.bm25_topk([corpus["oseg"]], kind, oracle_terms(corpus["reader"], scorer, tis), k, mode=1)
                assert_hits_equal(hits, oh)
                assert not np.isin(hits["doc"], deleted).any()
                if wand == 0 or len(tis) > 1:
                    assert total == ototal
        ctx().set_wand(0)
        filt = sdb.pred(9, "BETWEEN", 250000, 749999)
        hits, total = sdb.ExecuteTopK(corpus["reader"], [0, 1], sdb.OR, scorer, 500, filt=filt)
        oh, ototal, _ = orc.bm25_topk([corpus["oseg"]], "OR", oracle_terms(corpus["reader"], scorer, [0, 1]), 500,
                                      filt=orc.make_pred(9, "BETWEEN", 250000, 749999), mode=1)
        assert_hits_equal(hits, oh)
        assert total == ototal
    finally:
        ctx().set_wand(0)
        corpus["oseg"].set_docs_mask([])
        corpus["gseg"].stage_docs_mask(None)
    hits, total = sdb.ExecuteTopK(corpus["reader"], [3], sdb.OR, scorer, 10)
    assert total == len(corpus["lists"][3][0])


@pytest.mark.parametrize("normalize", [False, True])
@pytest.mark.parametrize("kind,tis,k", [("OR", [3, 5], 100), ("OR", [0], 50), ("AND", [0, 1, 4], 100), ("OR", [2, 6, 7, 8], 200)])
def test_tfidf_scorer(corpus, normalize, kind, tis, k):
    """irs::TFIDF (search/tfidf.cpp:59-80, 149-150) on the same scan: sqrt(freq) * idf [/ sqrt(doc length)], bit-exact
    against the oracle's restatement (pinned to the reference's rank-order goldens by test_tfidf_sequential_order_goldens)."""
    scorer = sdb.TFIDF(normalize=normalize)
    hits, total = sdb.ExecuteTopK(corpus["reader"], tis, sdb.AND if kind == "AND" else sdb.OR, scorer, k)
    terms = []
    for t in tis:
        x = orc.BM25Term()
        x.idf = orc.tfidf_idf(corpus["reader"].docs_with_field, int(corpus["reader"].docs_with_term[t]))
        x.norm_const, x.norm_length, x.boost, x.term = 0.0, 0.0, 1.0, t
        terms.append(x)
    oh, ototal, _ = orc.bm25_topk([corpus["oseg"]], kind, terms, k, k1=-1.0, b=1.0 if normalize else 0.0, mode=1)
    assert_hits_equal(hits, oh)
    assert total == ototal
    # spot value: the top hit's score recomputed in float32 with numpy (sqrt, *, / are all correctly rounded)
    d0 = int(oh["doc"][0])
    s = np.float32(0)
    for t, x in sorted(zip(tis, terms), key=lambda p: len(corpus["lists"][p[0]][0])):
        docs, freqs = corpus["lists"][t]
        i = np.searchsorted(docs, d0)
        if i < len(docs) and docs[i] == d0:
            v = np.float32(np.sqrt(np.float32(freqs[i]))) * np.float32(x.idf)
            if normalize:
                v = np.float32(v / np.float32(np.sqrt(np.float32(corpus["dl"][d0 - 1]))))
            s = np.float32(s + v)
    assert s == oh["score"][0]


STREAM_QUERIES = [("OR", [3]), ("OR", [0, 4]), ("OR", [2, 5, 6]), ("OR", [0, 1, 2, 3]), ("AND", [0, 1, 2]), ("AND", [0, 8]),
                  ("AND", [0, 1, 2, 3, 4]), ("OR", [8])]


@pytest.mark.parametrize("kind,tis", STREAM_QUERIES)
def test_stream_scored_docs(corpus, kind, tis):
    """Streaming mode (RunStreamingScan / EmitScoredDocs): every match with its score, ascending by doc, equals the
    oracle's exhaustive result re-sorted by doc -- bit-exact; doc windows partition the stream."""
    scorer = sdb.BM25()
    knd = sdb.AND if kind == "AND" else sdb.OR
    docs, scores = sdb.StreamScoredDocs(corpus["reader"], 0, tis, knd, scorer)
    oh, ototal, _ = orc.bm25_topk([corpus["oseg"]], kind, oracle_terms(corpus["reader"], scorer, tis), corpus["n"], mode=0)
    order = np.argsort(oh["doc"], kind="stable")
    assert len(docs) == ototal == len(oh)
    assert np.array_equal(docs, oh["doc"][order])
    assert np.array_equal(scores.view(np.uint32), oh["score"][order].view(np.uint32))
    assert np.all(np.diff(docs.astype(np.int64)) > 0)
    # chunked like the scan's EmitChunk windows: [1, a) + [a, b) + [b, end) == the whole stream
    a, b = 77_777, 200_001
    parts = [sdb.StreamScoredDocs(corpus["reader"], 0, tis, knd, scorer, doc_min=lo, doc_max=hi) for lo, hi in ((1, a), (a, b), (b, None))]
    assert np.array_equal(np.concatenate([p[0] for p in parts]), docs)
    assert np.array_equal(np.concatenate([p[1] for p in parts]).view(np.uint32), scores.view(np.uint32))


def test_stream_scored_docs_filter_and_mask(corpus):
    scorer = sdb.BM25()
    n = corpus["n"]
    deleted = np.unique(np.random.default_rng(5).integers(1, n + 1, size=n // 5)).astype(np.uint32)
    corpus["oseg"].set_docs_mask(deleted)
    corpus["gseg"].stage_docs_mask(deleted)
    try:
        for kind, tis in (("OR", [0, 1]), ("AND", [0, 1, 2])):
            docs, scores = sdb.StreamScoredDocs(corpus["reader"], 0, tis, sdb.AND if kind == "AND" else sdb.OR, scorer,
                                                filt=sdb.pred(9, "BETWEEN", 250000, 749999))
            oh, ototal, _ = orc.bm25_topk([corpus["oseg"]], kind, oracle_terms(corpus["reader"], scorer, tis), n,
                                          filt=orc.make_pred(9, "BETWEEN", 250000, 749999), mode=0)
            order = np.argsort(oh["doc"], kind="stable")
            assert len(docs) == ototal and ototal > 0
            assert np.array_equal(docs, oh["doc"][order])
            assert np.array_equal(scores.view(np.uint32), oh["score"][order].view(np.uint32))
            assert not np.isin(docs, deleted).any()
    finally:
        corpus["oseg"].set_docs_mask([])
        corpus["gseg"].stage_docs_mask(None)
    # empty window and unsupported shapes
    d, s = sdb.StreamScoredDocs(corpus["reader"], 0, [3], sdb.OR, scorer, doc_min=50, doc_max=50)
    assert len(d) == 0 and len(s) == 0
    with pytest.raises(Exception, match="EUNSUPPORTED"):
        sdb.StreamScoredDocs(corpus["reader"], 0, [0, 1, 2, 3, 4], sdb.OR, scorer)
