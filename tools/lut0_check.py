"""Diagnostic for the SDBG_STREAM_LUT=0 total_matches anomaly: the configs[2] batch with pruning off, once with
the score table and once without; prints the queries whose totals or hits differ. Run under compute-sanitizer to look
for shared-memory hazards:  compute-sanitizer --tool racecheck python tools/lut0_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import serenedb_b200 as sdb
import bench

ctx = sdb.Context(0)
n = int(os.environ.get("PROBE_DOCS", 10_000_000))
nq = int(os.environ.get("PROBE_QUERIES", 4096))
g = sdb.Segment(ctx, n)
dc, sum_dl = g.synth_corpus(0, 0, 256, threads=32)
reader = sdb.IndexReader([g], n, sum_dl, dc)
queries = bench.make_queries(nq)
batch = sdb.PreparedBatch(reader, queries, sdb.OR, sdb.BM25(), 1000)
ctx.set_wand(0)
res = {}
for lut in os.environ.get("PROBE_LUTS", "1,0,0").split(","):
    os.environ["SDBG_STREAM_LUT"] = lut
    h, nout, tot = batch.run_host()
    exp = np.array([int(dc[a]) + int(dc[b]) for a, b in queries])    # |A| + |B| >= |A u B|
    print("lut", lut, "sum totals", int(tot.sum()), flush=True)
    if lut in res:
        print("   repeat identical:", bool(np.array_equal(tot, res[lut][2])))
    res.setdefault(lut, (h.copy(), nout.copy(), tot.copy()))
if "1" in res and "0" in res:
    a, b = res["1"], res["0"]
    bad = np.nonzero(a[2] != b[2])[0]
    print("totals differ for", len(bad), "queries; hits identical:", bool(np.array_equal(a[0]["doc"], b[0]["doc"]) and np.array_equal(a[0]["score"], b[0]["score"])))
    for i in bad[:12]:
        print("  q", int(i), queries[i], "docs", int(dc[queries[i][0]]), int(dc[queries[i][1]]), "lut1", int(a[2][i]), "lut0", int(b[2][i]), "diff", int(a[2][i]) - int(b[2][i]))
