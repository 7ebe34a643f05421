"""Diagnostic: how much does block-max pruning skip? (TotalMatches with WAND = docs actually looked at.)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import serenedb_b200 as sdb

ctx = sdb.Context(0)
n = 10_000_000
g = sdb.Segment(ctx, n)
dc, sum_dl = g.synth_corpus(0, 0, 256, threads=32)
reader = sdb.IndexReader([g], n, sum_dl, dc)
scorer = sdb.BM25()
for q in ([0, 120], [0, 59], [5, 59], [2, 200], [0, 1], [30, 200], [0]):
    out = []
    for wand in (0, 1, 2):
        ctx.set_wand(wand)
        batch = sdb.PreparedBatch(reader, [q] * 64, sdb.OR, scorer, 1000)
        batch.run_host()
        ctx.sync(); t = time.perf_counter()
        h, nout, tot = batch.run_host()
        ctx.sync(); dt = time.perf_counter() - t
        out.append((int(tot[0]), round(dt * 1e3, 2), float(h[0, nout[0] - 1]["score"])))
    print(q, [int(dc[t]) for t in q], "off(total,ms,kth)", out[0], "L1", out[1], "L2", out[2], flush=True)
