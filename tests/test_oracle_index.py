"""Writer/reader restatement: .doc layout, skip list, block-max validity, pruned == exhaustive."""
import numpy as np
import pytest

import orc


def _q(seg, t, n_docs, total_tf, k=1.2, b=0.75, boost=1.0):
    m = seg.term_meta(t)
    st = orc.bm25_stats(n_docs, total_tf, m.docs_count, k, b)
    q = orc.BM25Term()
    q.idf, q.norm_const, q.norm_length, q.boost, q.term = st.idf, st.norm_const, st.norm_length, boost, t
    return q


@pytest.fixture(scope="module")
def corpus():
    n = 200_000
    terms = [0, 1, 3, 9, 30, 60, 200, 255]
    seg, dl, lists = orc.synth_segment(n, terms)
    return dict(seg=seg, dl=dl, lists=lists, n=n, terms=terms, total_tf=int(dl.sum()))


def test_roundtrip_all_lengths():
    """docs_count = 1 (inline), <128 (wand first), ==128 (wand after), >128 (skip list), multiples of 128."""
    n = 5000
    rng = np.random.default_rng(7)
    seg = orc.Segment(n, has_wand=True)
    dl = rng.integers(1, 300, size=n).astype(np.uint32)
    seg.set_norms(dl)
    cases = [1, 2, 5, 127, 128, 129, 255, 256, 257, 1000, 4096, 4097, 5000]
    lists = []
    for c in cases:
        docs = np.sort(rng.choice(np.arange(1, n + 1), size=c, replace=False)).astype(np.uint32)
        freqs = np.minimum(rng.geometric(0.5, size=c), dl[docs - 1]).astype(np.uint32)
        t = seg.add_term(docs, freqs)
        lists.append((t, docs, freqs))
    for t, docs, freqs in lists:
        m = seg.term_meta(t)
        assert m.docs_count == len(docs) and m.freq == int(freqs.sum())
        d, f = seg.decode_term(t)
        assert np.array_equal(d, docs) and np.array_equal(f, freqs)
        sk = seg.skip_level0(t)
        nblk = (len(docs) - 1) // 128
        if len(docs) > 128:
            assert len(sk["last_doc"]) == nblk            # no terminal entry (writer.hpp:763)
            assert np.array_equal(sk["last_doc"], docs[127::128][:nblk])
            assert sk["num_levels"] == (1 if nblk < 32 else 2)
            assert m.e_skip_start > 0
        else:
            assert len(sk["last_doc"]) == 0
        if len(docs) == 1:
            assert m.e_skip_start == docs[0] - 1          # e_single_doc
    # single-doc terms put nothing in .doc
    assert seg.term_meta(0).doc_start == 0 and seg.term_meta(1).doc_start == 0


def test_block_max_is_attained_upper_bound(corpus):
    """Each level-0 wand pair is a real (freq,norm) of its block and bounds every score in it."""
    seg, dl, n, total_tf = corpus["seg"], corpus["dl"], corpus["n"], corpus["total_tf"]
    for ti, (docs, freqs) in enumerate(corpus["lists"]):
        if len(docs) <= 128:
            continue
        q = _q(seg, ti, n, total_tf)
        st = orc.BM25Stats(q.idf, q.norm_const, q.norm_length)
        sk = seg.skip_level0(ti)
        scores = orc.bm25_score(freqs, dl[docs - 1], st)
        root = orc.bm25_score([sk["root"][0]], [sk["root"][1]], st)[0]
        assert root == scores.max()
        for j in range(len(sk["last_doc"])):
            blk = slice(128 * j, 128 * j + 128)
            pair = (sk["wand_freq"][j], sk["wand_norm"][j])
            assert pair in set(zip(freqs[blk], dl[docs[blk] - 1]))
            bound = orc.bm25_score([pair[0]], [pair[1]], st)[0]
            assert bound == scores[blk].max()


def _expected_topk(corpus, kind, tis, k, filt_mask=None):
    seg, dl, n, total_tf = corpus["seg"], corpus["dl"], corpus["n"], corpus["total_tf"]
    acc = np.zeros(n + 1, np.float32)
    cnt = np.zeros(n + 1, np.int32)
    order = sorted(range(len(tis)), key=lambda i: (len(corpus["lists"][tis[i]][0]), i))
    for i in order:
        docs, freqs = corpus["lists"][tis[i]]
        q = _q(seg, tis[i], n, total_tf)
        st = orc.BM25Stats(q.idf, q.norm_const, q.norm_length)
        s = orc.bm25_score(freqs, dl[docs - 1], st)
        acc[docs] = acc[docs] + s
        cnt[docs] += 1
    need = len(tis) if kind == "AND" else 1
    m = cnt >= need
    m[0] = False
    if filt_mask is not None:
        m[1:] &= filt_mask
    d = np.nonzero(m)[0]
    o = np.lexsort((d, -acc[d].astype(np.float64)))[:k]
    return d[o], acc[d][o], int(m.sum())


@pytest.mark.parametrize("kind,tis,k", [("OR", [3], 10), ("OR", [0, 4], 100), ("OR", [2, 5, 6], 1000),
                                        ("AND", [0, 1, 2], 50), ("AND", [0, 1, 2, 3, 4], 1000),
                                        ("OR", [7], 5000)])
def test_topk_modes_agree_with_numpy(corpus, kind, tis, k):
    seg, n, total_tf = corpus["seg"], corpus["n"], corpus["total_tf"]
    terms = [_q(seg, t, n, total_tf) for t in tis]
    ed, es, etotal = _expected_topk(corpus, kind, tis, k)
    for mode in (0, 1, 2):
        hits, total, scored = orc.bm25_topk([seg], kind, terms, k, mode=mode)
        assert np.array_equal(hits["doc"], ed), mode
        assert np.array_equal(hits["score"], es), mode
        if mode < 2:
            assert total == etotal
        else:
            assert total <= etotal  # with WAND TotalMatches is a lower bound (wand_scoring_test.cpp:382-384)


def test_pruning_skips_blocks(corpus):
    seg, n, total_tf = corpus["seg"], corpus["n"], corpus["total_tf"]
    terms = [_q(seg, 0, n, total_tf)]
    h0, _, s0 = orc.bm25_topk([seg], "OR", terms, 10, mode=1)
    h2, _, s2 = orc.bm25_topk([seg], "OR", terms, 10, mode=2)
    assert np.array_equal(h0, h2)
    assert s2 < s0 // 2  # single-term WAND with k=10 must skip most blocks


def test_hybrid_filter_and_multisegment(corpus):
    n = corpus["n"]
    # two segments by doc range, global statistics (collectors.cpp:36-52): results equal one big segment
    half = n // 2
    segA, dlA, _ = orc.synth_segment(half, corpus["terms"], doc0=0)
    segB, dlB, _ = orc.synth_segment(n - half, corpus["terms"], doc0=half)
    nn = np.array([orc.lib().orc_synth_hash(2, d) % 1000000 for d in range(1, n + 1)], np.int32)
    corpus["seg"].add_column(9, nn)
    segA.add_column(9, nn[:half])
    segB.add_column(9, nn[half:])
    filt = orc.make_pred(9, "BETWEEN", 250000, 749999)
    tis = [0, 1, 2]
    total_tf = corpus["total_tf"]

    def q(seg_list, t):
        dwt = sum(s.term_meta(t).docs_count for s in seg_list)
        st = orc.bm25_stats(n, total_tf, dwt)
        x = orc.BM25Term()
        x.idf, x.norm_const, x.norm_length, x.boost, x.term = st.idf, st.norm_const, st.norm_length, 1.0, t
        return x

    one = [q([corpus["seg"]], t) for t in tis]
    two = [q([segA, segB], t) for t in tis]
    for a, b in zip(one, two):
        assert a.idf == b.idf
    mask = (nn >= 250000) & (nn <= 749999)
    ed, es, etotal = _expected_topk(corpus, "AND", tis, 100, filt_mask=mask)
    h1, t1, _ = orc.bm25_topk([corpus["seg"]], "AND", one, 100, filt=filt, mode=1)
    assert np.array_equal(h1["doc"], ed) and np.array_equal(h1["score"], es) and t1 == etotal
    h2, t2, _ = orc.bm25_topk([segA, segB], "AND", two, 100, filt=filt, mode=0)
    glob = np.where(h2["seg"] == 0, h2["doc"], h2["doc"] + half)
    assert np.array_equal(glob, ed) and np.array_equal(h2["score"], es) and t2 == etotal


def _f32(x):
    return np.float32(x)


@pytest.mark.parametrize("kind,tis", [("OR", [3]), ("OR", [0, 4]), ("AND", [0, 1, 2])])
def test_bm15_and_bm1_forms(corpus, kind, tis):
    """b == 0 selects Bm15 (bm25.cpp:70-87: c0 - c0 / (1 + freq / c1), c1 = k, no norms) and k == 0 Bm1
    (:112-126: every score 0 without a filter boost, so nothing beats the FLT_MIN seed). Numpy float32
    arithmetic is IEEE single without contraction, i.e. the reference's operation order."""
    seg, n, total_tf = corpus["seg"], corpus["n"], corpus["total_tf"]
    k1 = 1.2
    terms = [_q(seg, t, n, total_tf, k=k1, b=0.0) for t in tis]
    acc = np.zeros(n + 1, np.float32)
    cnt = np.zeros(n + 1, np.int32)
    order = sorted(range(len(tis)), key=lambda i: (len(corpus["lists"][tis[i]][0]), i))
    for i in order:
        docs, freqs = corpus["lists"][tis[i]]
        q = terms[i]
        assert q.norm_length == 0.0 and q.norm_const == _f32(k1)
        c0 = _f32(_f32(_f32(q.boost) * _f32(_f32(k1) + _f32(1))) * _f32(q.idf))
        s = c0 - c0 / (_f32(1) + freqs.astype(np.float32) / _f32(q.norm_const))
        assert s.dtype == np.float32
        acc[docs] = acc[docs] + s
        cnt[docs] += 1
    m = cnt >= (len(tis) if kind == "AND" else 1)
    d = np.nonzero(m)[0]
    o = np.lexsort((d, -acc[d].astype(np.float64)))[:500]
    for mode in (0, 1, 2):        # mode 2 must fall back to an exhaustive scan: the block-max pairs are BM25's
        hits, total, _ = orc.bm25_topk([seg], kind, terms, 500, k1=k1, b=0.0, mode=mode)
        assert np.array_equal(hits["doc"], d[o]) and np.array_equal(hits["score"], acc[d][o]), mode
        assert total == int(m.sum())
    # BM1
    terms0 = [_q(seg, t, n, total_tf, k=0.0, b=0.75) for t in tis]
    hits, total, _ = orc.bm25_topk([seg], kind, terms0, 500, k1=0.0, b=0.75, mode=0)
    assert len(hits) == 0 and total == int(m.sum())


def test_docs_mask(corpus):
    """Deleted docs (DocumentMask) are invisible to scoring, the collector and the match count
    (MaskDocIterator wraps the query iterator, segment_reader_impl.cpp:95-157,318-326)."""
    seg, n, total_tf = corpus["seg"], corpus["n"], corpus["total_tf"]
    rng = np.random.default_rng(31)
    deleted = np.unique(rng.integers(1, n + 1, size=n // 3)).astype(np.uint32)
    keep = np.ones(n, bool)
    keep[deleted - 1] = False
    seg.set_docs_mask(deleted)
    try:
        for kind, tis, k in (("OR", [0, 4], 100), ("AND", [0, 1, 2], 50), ("OR", [7], 5000)):
            terms = [_q(seg, t, n, total_tf) for t in tis]
            ed, es, etotal = _expected_topk(corpus, kind, tis, k, filt_mask=keep)
            for mode in (0, 1, 2):
                hits, total, _ = orc.bm25_topk([seg], kind, terms, k, mode=mode)
                assert np.array_equal(hits["doc"], ed) and np.array_equal(hits["score"], es), (kind, mode)
                assert total == etotal if mode < 2 else total <= etotal
    finally:
        seg.set_docs_mask([])
    hits, total, _ = orc.bm25_topk([seg], "OR", [_q(seg, 0, n, total_tf)], 10)
    assert total == len(corpus["lists"][0][0])        # cleared
