"""Block-max pruning on the GPU: the differential property of the reference's wand tests
(tests/libs/iresearch/search/wand_scoring_test.cpp:361-393, CompareWandVsNonWand): pruned top-k ==
exhaustive top-k (same docs, same fp32 scores); TotalMatches is only a lower bound with WAND (:382-384)."""
import numpy as np
import pytest

import orc
import serenedb_b200 as sdb
from gpu_util import assert_hits_equal, ctx, oracle_terms, to_gpu

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def corpus():
    n = 2_000_000
    g = sdb.Segment(ctx(), n)
    dc, sum_dl = g.synth_corpus(0, 0, 64, threads=8)
    oseg, odc, osdl = orc.synth_segment_mt(n, 0, 64, threads=8)
    assert np.array_equal(dc, odc) and sum_dl == osdl
    return dict(g=g, oseg=oseg, reader=sdb.IndexReader([g], n, sum_dl, dc), dc=dc)


@pytest.fixture(autouse=True, params=[1, 2])
def wand_on(request):
    ctx().set_wand(request.param)
    yield request.param
    ctx().set_wand(0)


QUERIES = [("OR", [0], 10), ("OR", [5], 1000), ("OR", [0, 40], 100), ("OR", [0, 40], 1000), ("OR", [1, 12], 1000),
           ("OR", [3, 5], 1000), ("OR", [2, 30, 60], 500), ("OR", [0, 1, 2, 3, 4], 1000), ("OR", [63], 1000), ("OR", [20, 21], 1)]


@pytest.mark.parametrize("kind,tis,k", QUERIES)
def test_wand_equals_exhaustive(corpus, kind, tis, k):
    scorer = sdb.BM25()
    hits, total = sdb.ExecuteTopK(corpus["reader"], tis, sdb.OR, scorer, k)
    oh, ototal, _ = orc.bm25_topk([corpus["oseg"]], kind, oracle_terms(corpus["reader"], scorer, tis), k, mode=1)
    assert_hits_equal(hits, oh)
    assert total <= ototal


def test_pruning_actually_skips(corpus):
    """Single dense term, k = 10, and a dense+sparse disjunction: most blocks cannot beat the threshold,
    so far fewer docs are even looked at (visible as a much smaller TotalMatches)."""
    scorer = sdb.BM25()
    # one chain per query (the batch shape): thresholds are published when a chain's buffer fills
    q1, q2 = [[0]] * 4, [[0, 40]] * 4
    ctx().set_wand(0)
    _, _, exact1 = sdb.ExecuteTopKBatch(corpus["reader"], q1 * 256, sdb.OR, scorer, 10)
    _, _, exact2 = sdb.ExecuteTopKBatch(corpus["reader"], q2 * 256, sdb.OR, scorer, 100)
    ctx().set_wand(1)
    _, _, l1 = sdb.ExecuteTopKBatch(corpus["reader"], q1 * 256, sdb.OR, scorer, 10)
    ctx().set_wand(2)
    _, _, l2 = sdb.ExecuteTopKBatch(corpus["reader"], q2 * 256, sdb.OR, scorer, 100)
    assert exact1[0] == corpus["dc"][0], (exact1[:4], corpus["dc"][0])
    assert l1[0] < exact1[0] // 2, (l1[:4], exact1[:4])     # single term: block skip in the planner
    assert l2[0] < exact2[0] * 3 // 4, (l2[:4], exact2[:4])  # dense + sparse: exact-partial-score skip


def test_wand_with_threshold_seed_filter_and_batch(corpus):
    scorer = sdb.BM25()
    n = corpus["reader"].docs_with_field
    nn = orc.synth_column(2, 1, 1, n).astype(np.int32)
    corpus["oseg"].add_column(9, nn)
    corpus["g"].stage_column(9, nn)
    filt_g, filt_o = sdb.pred(9, "BETWEEN", 250000, 749999), orc.make_pred(9, "BETWEEN", 250000, 749999)
    hits, total = sdb.ExecuteTopK(corpus["reader"], [1, 30], sdb.OR, scorer, 200, filt=filt_g)
    oh, ototal, _ = orc.bm25_topk([corpus["oseg"]], "OR", oracle_terms(corpus["reader"], scorer, [1, 30]), 200, filt=filt_o, mode=1)
    assert_hits_equal(hits, oh)
    assert total <= ototal
    rng = np.random.default_rng(4)
    queries = [[int(a), int(b)] for a, b in (rng.choice(64, size=2, replace=False) for _ in range(96))]
    bh, bn, bt = sdb.ExecuteTopKBatch(corpus["reader"], queries, sdb.OR, scorer, 100)
    for qi in range(0, 96, 7):
        oh, _, _ = orc.bm25_topk([corpus["oseg"]], "OR", oracle_terms(corpus["reader"], scorer, queries[qi]), 100, mode=1)
        assert_hits_equal(bh[qi, :bn[qi]], oh)
