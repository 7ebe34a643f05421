"""The C++ adapters (serenedb_b200/host) driven like the reference's callers, checked against the oracle."""
import json
import subprocess

import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu


def test_adapters_match_oracle():
    from serenedb_b200 import build as b
    exe = b.build_adapters()
    n = 200_000
    res = subprocess.run([exe, str(n)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    topk, stream, agg, mt = [json.loads(l) for l in res.stdout.strip().splitlines()]
    # ---- GpuTopKIterator::Collect vs oracle (2-term OR + INCLUDE-column range filter, k = 100) ----
    oseg, dc, sum_dl = orc.synth_segment_mt(n, 0, 8, threads=4)
    nn = orc.synth_column(2, 1, 1, n).astype(np.int32)
    oseg.add_column(9, nn)
    terms = []
    for t in (2, 5):
        st = orc.bm25_stats(n, sum_dl, int(dc[t]))
        x = orc.BM25Term()
        x.idf, x.norm_const, x.norm_length, x.boost, x.term = st.idf, st.norm_const, st.norm_length, 1.0, t
        terms.append(x)
    oh, ototal, _ = orc.bm25_topk([oseg], "OR", terms, 100, filt=orc.make_pred(9, "BETWEEN", 250000, 749999), mode=1)
    assert [d for d, _ in topk["topk"]] == oh["doc"].tolist()
    assert np.array_equal(np.array([s for _, s in topk["topk"]], np.float32), oh["score"])
    assert topk["total"] == ototal
    assert np.float32(topk["threshold"]) == oh["score"][-1]      # threshold raised to the k-th score
    # the reference's accessors: irs::get<ScoreThresholdAttr> / irs::get<CostAttr> through AttributeProvider::GetMutable
    assert np.float32(topk["attr_threshold"]) == oh["score"][-1] and topk["attr_cost"] == ototal
    # FillBlock over docs [1, 4097): one bit and one score per hit in the window, value() = first hit beyond it
    inwin = oh[oh["doc"] < 4097]
    later = oh["doc"][oh["doc"] >= 4097]
    assert topk["fill_bits"] == len(inwin)
    assert topk["fill_sum"] == pytest.approx(float(inwin["score"].astype(np.float64).sum()), rel=1e-6)
    assert topk["fill_next"] == (int(later.min()) if len(later) else 0xFFFFFFFF)
    # ---- streaming mode (k = 0): EmitScoredDocs windows drain every match in doc order ----
    allh, alltotal, _ = orc.bm25_topk([oseg], "OR", terms, n, filt=orc.make_pred(9, "BETWEEN", 250000, 749999), mode=0)
    assert stream["stream_n"] == alltotal == len(allh) == stream["stream_count"]
    assert stream["stream_ordered"] == 1
    assert stream["stream_doc_sum"] == int(allh["doc"].astype(np.uint64).sum())
    by_doc = allh[np.argsort(allh["doc"], kind="stable")]
    assert stream["stream_score_sum"] == pytest.approx(float(by_doc["score"].astype(np.float64).sum()), rel=1e-10)
    assert stream["stream_chunks"] == len(np.unique((by_doc["doc"] - 1) // 2048))
    # ---- GpuAggScan chunks vs oracle GROUP BY ----
    cols = {10: (10, 0), 11: (11, 1), 12: (12, 2), 13: (13, 3), 14: (14, 4)}
    for f, (stream, kind) in cols.items():
        oseg.add_column(f, orc.synth_column(stream, kind, 0, n))
    exp = orc.filter_groupby([oseg], [orc.make_pred(11, "LT", 500000), orc.make_pred(12, "GE", 0.25, is_float=True)], 10, 13, 14, cap=100001)
    assert agg["groups"] == len(exp) and agg["rows"] == int(exp["count"].sum())
    assert agg["chunks"] == (len(exp) + 2047) // 2048           # <= STANDARD_VECTOR_SIZE rows per call
    assert agg["sum_v"] == int(exp["sum_lo"].astype(object).sum())
    assert agg["avg_sum"] == pytest.approx(float((exp["sum_f64"] / exp["cnt_f64"]).sum()), rel=1e-9)
    # ---- the same mode under DuckDB's threading contract: four workers share one global state and claim chunks atomically ----
    assert mt["mt_ok"] == 1 and mt["mt_groups"] == len(exp) == mt["mt_emitted"] and mt["mt_rows"] == int(exp["count"].sum())
    assert mt["mt_chunks"] == (len(exp) + 2047) // 2048 and mt["mt_sum_v"] == int(exp["sum_lo"].astype(object).sum())
