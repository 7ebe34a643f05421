"""Debug helper: a few queries on a small corpus through the stream kernel with / without the score table; prints
total_matches against the legacy kernel. Run under compute-sanitizer when hunting races."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import serenedb_b200 as sdb

ctx = sdb.Context(0)
ctx.set_wand(0)
n = int(os.environ.get("PROBE_DOCS", 2_000_000))
g = sdb.Segment(ctx, n)
dc, sum_dl = g.synth_corpus(0, 0, 96, threads=16)
reader = sdb.IndexReader([g], n, sum_dl, dc)
scorer = sdb.BM25()
queries = [[81, 1], [0, 1], [5, 59], [1, 36], [0], [0, 1, 2], [3, 40, 70, 90]]
ref = None
for env in ({"SDBG_STREAM": "0"}, {"SDBG_STREAM": "1"}, {"SDBG_STREAM": "1", "SDBG_STREAM_LUT": "0"}):
    for k_ in ("SDBG_STREAM", "SDBG_STREAM_LUT"):
        os.environ.pop(k_, None)
    os.environ.update(env)
    for rep in range(int(os.environ.get("PROBE_REPS", 2))):
        batch = sdb.PreparedBatch(reader, queries, sdb.OR, scorer, 100)
        h, nout, tot = batch.run_host()
        if ref is None:
            ref = (h.copy(), nout.copy(), tot.copy())
        ok = all(np.array_equal(h[i][: nout[i]]["doc"], ref[0][i][: ref[1][i]]["doc"]) and np.array_equal(h[i][: nout[i]]["score"], ref[0][i][: ref[1][i]]["score"]) for i in range(len(queries)))
        print(env, "hits ok", ok, "totals", tot.tolist(), "ref", ref[2].tolist(), flush=True)
