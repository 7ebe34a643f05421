"""GROUP BY / SUM / AVG / COUNT pinned to the answers the reference's sqllogic tests hold (tests/golden/groupby_goldens.json
cites each file:line): the oracle's aggregate -- the definition the GPU path is compared with everywhere else -- has
to reproduce them. String group keys are mapped to integers; the aggregate semantics (COUNT(*), exact SUM(int), AVG =
SUM(double) / COUNT) are what is being pinned."""
import json
import os

import numpy as np

import orc

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "groupby_goldens.json")))


def run(keys, ints=None, floats=None, preds=()):
    n = len(keys)
    seg = orc.Segment(n, has_wand=False)
    seg.add_column(1, np.asarray(keys, np.int64))
    seg.add_column(2, np.asarray(ints if ints is not None else [0] * n, np.int64))
    seg.add_column(3, np.asarray(floats if floats is not None else [0.0] * n, np.float64))
    return orc.filter_groupby([seg], list(preds), 1, 2, 3, cap=64)


def test_sum_by_region():
    g = G["sales_sum_by_region"]
    out = run(g["rows"]["key"], ints=g["rows"]["amount"])
    got = {g["key_names"][int(k)]: int(s) for k, s in zip(out["key"], out["sum_lo"])}
    assert got == g["expect_sum"] and sum(got.values()) == g["expect_total_sum"]
    assert np.all(out["sum_hi"] == 0)


def test_count_by_city():
    g = G["addresses_count_by_city"]
    out = run(g["rows"]["key"])
    assert {g["key_names"][int(k)]: int(c) for k, c in zip(out["key"], out["count"])} == g["expect_count"]


def test_avg_by_city_street():
    g = G["addresses_avg_income_by_city_street"]
    out = run(g["rows"]["key"], floats=g["rows"]["income"])
    got = {g["key_names"][int(k)]: float(s / c) for k, s, c in zip(out["key"], out["sum_f64"], out["cnt_f64"])}
    assert got == {k: float(v) for k, v in g["expect_avg"].items()}


def test_facet_counts():
    g = G["products_facets"]
    for col, names, exp in (("category", "category_names", "expect_category"), ("brand", "brand_names", "expect_brand"), ("band", "band_names", "expect_band")):
        out = run(g["rows"][col])
        assert {g[names][int(k)]: int(c) for k, c in zip(out["key"], out["count"])} == g[exp]
