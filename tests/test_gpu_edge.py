"""Edge cases and error behaviour of the C ABI on the GPU: empty / single-doc / full lists, 16-term
queries, k = 1, nullable filter columns, int32 predicates, and the error codes a caller sees."""
import numpy as np
import pytest

import orc
import serenedb_b200 as sdb
from serenedb_b200 import _native
from gpu_util import assert_hits_equal, ctx, oracle_terms, to_gpu

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def small():
    n = 20_000
    rng = np.random.default_rng(17)
    oseg = orc.Segment(n, has_wand=True)
    dl = rng.integers(1, 70000, size=n).astype(np.uint32)     # wide norms -> 4-byte width... capped below
    dl = np.minimum(dl, 60000).astype(np.uint32)
    oseg.set_norms(dl)
    lists = []
    sizes = [0, 1, n, 7, 128, 129, 3000] + [200 + 37 * i for i in range(12)]
    for c in sizes:
        docs = np.sort(rng.choice(np.arange(1, n + 1), size=c, replace=False)).astype(np.uint32) if c else np.zeros(0, np.uint32)
        freqs = np.minimum(rng.geometric(0.3, size=c), dl[docs - 1] if c else 1).astype(np.uint32) if c else np.zeros(0, np.uint32)
        oseg.add_term(docs, freqs)
        lists.append((docs, freqs))
    vals = rng.integers(-50, 50, size=n).astype(np.int32)
    valid = rng.integers(0, 2**63, size=(n + 63) // 64, dtype=np.int64).astype(np.uint64)
    oseg.add_column(5, vals, valid)
    g = to_gpu(oseg, columns={5: (vals, valid)})
    reader = sdb.IndexReader([g], n, int(dl.sum()), [len(d) for d, _ in lists])
    return dict(oseg=oseg, g=g, reader=reader, lists=lists, n=n, vals=vals, valid=valid)


@pytest.mark.parametrize("kind,tis,k", [("OR", [0], 10), ("OR", [1], 10), ("OR", [2], 1), ("OR", [0, 1], 5), ("AND", [0, 2], 5),
                                        ("AND", [1, 2], 5), ("OR", [1, 3, 4], 100), ("OR", list(range(3, 19)), 1000),
                                        ("AND", list(range(2, 18)), 1000), ("OR", [2, 6], 8192), ("AND", [2, 6, 5], 1)])
def test_odd_lists_and_wide_queries(small, kind, tis, k):
    scorer = sdb.BM25(0.9, 0.4)
    hits, total = sdb.ExecuteTopK(small["reader"], tis, sdb.AND if kind == "AND" else sdb.OR, scorer, k)
    oh, ototal, _ = orc.bm25_topk([small["oseg"]], kind, oracle_terms(small["reader"], scorer, tis), k, k1=0.9, mode=1)
    assert_hits_equal(hits, oh)
    assert total == ototal


def test_nullable_int32_filter(small):
    scorer = sdb.BM25()
    for op, lo, hi in [("GE", 0, 0), ("BETWEEN", -10, 10), ("IS_NULL", 0, 0), ("IS_NOT_NULL", 0, 0), ("NE", 3, 0)]:
        hits, total = sdb.ExecuteTopK(small["reader"], [2, 6], sdb.OR, scorer, 300, filt=sdb.pred(5, op, lo, hi))
        oh, ototal, _ = orc.bm25_topk([small["oseg"]], "OR", oracle_terms(small["reader"], scorer, [2, 6]), 300,
                                      filt=orc.make_pred(5, op, lo, hi), mode=1)
        assert_hits_equal(hits, oh)
        assert total == ototal


def test_error_codes(small):
    scorer = sdb.BM25()
    with pytest.raises(_native.SdbgError, match="EINVAL"):
        sdb.ExecuteTopK(small["reader"], [999], sdb.OR, scorer, 10) if False else sdb.ExecuteTopKBatch(
            sdb.IndexReader([small["g"]], small["n"], 1, np.ones(2000)), [[1500]], sdb.OR, scorer, 10)
    with pytest.raises(_native.SdbgError, match="EUNSUPPORTED"):
        sdb.ExecuteTopK(small["reader"], list(range(3, 19)) + [2], sdb.OR, scorer, 10)      # 17 terms
    with pytest.raises(_native.SdbgError, match="EUNSUPPORTED"):
        sdb.ExecuteTopK(small["reader"], [2], sdb.OR, scorer, 10000)                         # k > 8192
    with pytest.raises(_native.SdbgError, match="ENOTFOUND"):
        sdb.ExecuteTopK(small["reader"], [2], sdb.OR, scorer, 10, filt=sdb.pred(77, "GT", 0))
    scan = sdb.IResearchScan([small["g"]])
    with pytest.raises(_native.SdbgError, match="ENOTFOUND"):
        scan.count_sum([sdb.pred(123, "GT", 0)])
    with pytest.raises(_native.SdbgError, match="ECAPACITY"):
        g2 = sdb.Segment(ctx(), 1000)
        g2.stage_column(1, np.arange(1000, dtype=np.int64))
        sdb.IResearchScan([g2]).groupby([], 1, cap=10)
    with pytest.raises(_native.SdbgError, match="EFORMAT"):
        bad = small["oseg"].doc_bytes().copy()
        bad[:64] = 0xFF
        from gpu_util import metas_of
        sdb.Segment(ctx(), small["n"]).stage_postings(bad, metas_of(small["oseg"]))


def test_count_sum_without_predicates_and_tiny_tables():
    for rows in (1, 2, 3):
        g = sdb.Segment(ctx(), rows)
        x = np.arange(1, rows + 1, dtype=np.int64) * 7
        y = np.arange(rows, dtype=np.float64) + 0.5
        g.stage_column(1, x)
        g.stage_column(2, y)
        scan = sdb.IResearchScan([g])
        assert scan.count_sum([], 1)[:2] == (rows, int(x.sum()))
        assert scan.count_sum([], 2)[2] == pytest.approx(float(y.sum()))
        got = scan.groupby([], 1, sum_int_field=1, avg_f64_field=2)
        assert got["key"].tolist() == x.tolist() and got["count"].tolist() == [1] * rows
