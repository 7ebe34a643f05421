"""Diagnostic: the 4096-query configs[2] batch through the legacy window kernel and the warp-autonomous stream
kernel (SDBG_STREAM), hits compared bit for bit, device time for each."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import serenedb_b200 as sdb
import bench

ctx = sdb.Context(0)
n = int(os.environ.get("PROBE_DOCS", 10_000_000))
nq = int(os.environ.get("PROBE_QUERIES", 4096))
g = sdb.Segment(ctx, n)
dc, sum_dl = g.synth_corpus(0, 0, 256, threads=32)
reader = sdb.IndexReader([g], n, sum_dl, dc)
scorer = sdb.BM25()
queries = bench.make_queries(nq)
postings = sum(int(dc[t]) for q in queries for t in q)
batch = sdb.PreparedBatch(reader, queries, sdb.OR, scorer, 1000)
d_keys = torch.empty(len(queries) * 1000, dtype=torch.int64, device="cuda:0")
ref = None
configs = [({"SDBG_STREAM": "0"}, 0), ({"SDBG_STREAM": "1"}, 0), ({"SDBG_STREAM": "1", "SDBG_STREAM_OCC": "2"}, 0), ({"SDBG_STREAM": "1", "SDBG_STREAM_LUT": "0"}, 0),
           ({"SDBG_STREAM": "0"}, 2), ({"SDBG_STREAM": "1"}, 2)]
if os.environ.get("PROBE_ONLY_STREAM"):
    configs = [({"SDBG_STREAM": "1", "SDBG_STREAM_OCC": os.environ.get("PROBE_OCC", "3")}, int(os.environ.get("PROBE_WAND", "0")))]
for env, wand in configs:
    for k_ in ("SDBG_STREAM", "SDBG_STREAM_LUT", "SDBG_STREAM_OCC"):
        os.environ.pop(k_, None)
    os.environ.update(env)
    ctx.set_wand(wand)
    h, nout, tot = batch.run_host()
    if ref is None:
        ref = (h.copy(), nout.copy(), tot.copy())
    elif os.environ.get("PROBE_ONLY_STREAM"):
        pass
    else:
        ok = np.array_equal(nout, ref[1]) and np.array_equal(h["doc"], ref[0]["doc"]) and np.array_equal(h["score"], ref[0]["score"])
        print("   hits identical to legacy/wand0:", ok, " totals equal:", bool(np.array_equal(tot, ref[2])), flush=True)
        if not np.array_equal(tot, ref[2]):
            bad = np.nonzero(tot != ref[2])[0]
            print("   totals differ for", len(bad), "queries:", [(int(i), queries[i], int(tot[i]), int(ref[2][i])) for i in bad[:8]])
        if not ok:
            bad = [i for i in range(len(queries)) if not (np.array_equal(h[i]["doc"], ref[0][i]["doc"]) and np.array_equal(h[i]["score"], ref[0][i]["score"]))]
            print("   first bad queries:", bad[:10], [queries[i] for i in bad[:10]])
    batch.run_device(0, d_keys.data_ptr())
    ctx.flush_l2(); ctx.sync(); ctx.timer_start()
    for _ in range(3):
        batch.run_device(0, d_keys.data_ptr())
    ms = ctx.timer_stop() / 3
    print(env, "wand", wand, "ms", round(ms, 3), "G postings/s", round(postings / ms / 1e6, 1), "seen %.0f%%" % (100.0 * float(tot.sum()) / float(ref[2].sum())), flush=True)
if os.environ.get("PROBE_ONLY_STREAM"):
    sys.exit(0)
# configs[3]: 5-term AND + range filter
g.synth_column(9, 2, 6, 1, n)
filt = sdb.pred(9, "BETWEEN", 250000, 749999)
b4 = sdb.PreparedBatch(reader, [[0, 1, 2, 3, 4]] * 64, sdb.AND, scorer, 1000, filt=filt)
p4 = sum(int(dc[t]) for t in range(5))
r4 = None
for env in ({"SDBG_STREAM": "0"}, {"SDBG_STREAM": "1"}):
    os.environ.update(env)
    ctx.set_wand(0)
    h4, n4, t4 = b4.run_host()
    if r4 is None:
        r4 = (h4.copy(), n4.copy(), t4.copy())
    else:
        print("   AND hits identical:", bool(np.array_equal(h4["doc"], r4[0]["doc"]) and np.array_equal(h4["score"], r4[0]["score"])), "totals", int(t4[0]), int(r4[2][0]))
    ctx.flush_l2(); ctx.sync(); ctx.timer_start()
    h4, n4, t4 = b4.run_host()
    ms4 = ctx.timer_stop()
    print(env, "configs[3] 64x 5-term AND + filter: ms/query", round(ms4 / 64, 4), "G postings/s", round(64 * p4 / ms4 / 1e6, 1), flush=True)
