"""Diagnostic: what does block-max pruning buy, per level? (TotalMatches with pruning = docs actually
looked at.) Also asserts that every level returns bit-identical hits to level 0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import serenedb_b200 as sdb
import bench

ctx = sdb.Context(0)
n = int(os.environ.get("PROBE_DOCS", 10_000_000))
g = sdb.Segment(ctx, n)
dc, sum_dl = g.synth_corpus(0, 0, 256, threads=32)
reader = sdb.IndexReader([g], n, sum_dl, dc)
scorer = sdb.BM25()


def run(batch, reps=2):
    batch.run_host()
    ctx.sync(); ctx.timer_start()
    for _ in range(reps):
        r = batch.run_host()
    return r, ctx.timer_stop() / reps


for q in ([0, 120], [0, 59], [5, 59], [2, 200], [0, 1], [30, 200], [0, 5, 59], [3, 40, 41, 200], [0]):
    out, ref = [], None
    for wand in (0, 1, 2):
        ctx.set_wand(wand)
        batch = sdb.PreparedBatch(reader, [q] * 64, sdb.OR, scorer, 1000)
        (h, nout, tot), ms = run(batch)
        if ref is None:
            ref = (h[0, :nout[0]].copy(), int(nout[0]))
        else:
            assert int(nout[0]) == ref[1] and np.array_equal(h[0, :nout[0]]["doc"], ref[0]["doc"]) and \
                np.array_equal(h[0, :nout[0]]["score"], ref[0]["score"]), ("pruning changed the result", q, wand)
        out.append((int(tot[0]), round(ms, 2)))
    print(q, [int(dc[t]) for t in q], "off(total,ms)", out[0], "L1", out[1], "L2", out[2], flush=True)

# the bench's own query mix (configs[2]): 4096 two-term disjunctions, top-1000
queries = bench.make_queries(int(os.environ.get("PROBE_QUERIES", 4096)))
postings = sum(int(dc[t]) for q in queries for t in q)
ref = None
for wand in (0, 1, 2):
    ctx.set_wand(wand)
    batch = sdb.PreparedBatch(reader, queries, sdb.OR, scorer, 1000)
    (h, nout, tot), ms = run(batch, 1)
    if ref is None:
        ref = (h.copy(), nout.copy())
    else:
        assert np.array_equal(nout, ref[1])
        for i in range(len(queries)):
            assert np.array_equal(h[i, :nout[i]]["doc"], ref[0][i, :nout[i]]["doc"]) and \
                np.array_equal(h[i, :nout[i]]["score"], ref[0][i, :nout[i]]["score"]), ("pruning changed the result", queries[i], wand)
    print("mix level", wand, "ms", round(ms, 2), "G postings/s (nominal)", round(postings / ms / 1e6, 2),
          "looked-at fraction", round(float(tot.sum()) / float(ref_tot) if wand else 1.0, 3) if wand else 1.0, flush=True)
    if wand == 0:
        ref_tot = tot.sum()
