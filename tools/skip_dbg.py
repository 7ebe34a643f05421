import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import serenedb_b200 as sdb
ctx = sdb.Context(0)
n = 2_000_000
g = sdb.Segment(ctx, n)
dc, sum_dl = g.synth_corpus(0, 0, 96, threads=16)
reader = sdb.IndexReader([g], n, sum_dl, dc)
scorer = sdb.BM25()
qs = [[0, 1, 2]]
ctx.set_wand(0)
h0, n0, t0 = sdb.ExecuteTopKBatch(reader, qs, sdb.OR, scorer, 100)
for dbg in (0, 16, 32, 64, 128, 16 + 128, 32 + 64, 16 + 32 + 64, 240):
    os.environ["SDBG_STREAM_DBG"] = str(dbg)
    ctx.set_wand(1)
    bad = 0
    for rep in range(5):
        h, nn, tot = sdb.ExecuteTopKBatch(reader, qs, sdb.OR, scorer, 100)
        ok = np.array_equal(h["doc"], h0["doc"]) and np.array_equal(h["score"], h0["score"])
        bad += not ok
        if not ok and rep == 0:
            a = set(h[0]["doc"].tolist()); b = set(h0[0]["doc"].tolist())
            missing = sorted(b - a)[:5]; extra = sorted(a - b)[:5]
            ms = {int(d): float(s) for d, s in zip(h0[0]["doc"], h0[0]["score"])}
            es = {int(d): float(s) for d, s in zip(h[0]["doc"], h[0]["score"])}
            print("   missing", [(d, ms[d]) for d in missing], "extra", [(d, es[d]) for d in extra], "kth ref", float(h0[0][-1]["score"]))
            diff = [(int(d), float(s1), float(s2)) for d, s1, d2, s2 in zip(h[0]["doc"], h[0]["score"], h0[0]["doc"], h0[0]["score"]) if d == d2 and s1 != s2][:5]
            print("   score diffs", diff)
    print("dbg", dbg, "bad runs", bad, "of 5, seen", int(tot[0]), "of", int(t0[0]), flush=True)
