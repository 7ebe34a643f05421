"""Batch time of the configs[2] workload for {pruning off, default} x {score table on, off}; hits compared with the first run."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import serenedb_b200 as sdb
import bench

ctx = sdb.Context(0)
n, nq = 10_000_000, 4096
g = sdb.Segment(ctx, n)
dc, sum_dl = g.synth_corpus(0, 0, 256, threads=32)
reader = sdb.IndexReader([g], n, sum_dl, dc)
queries = bench.make_queries(nq)
postings = sum(int(dc[t]) for q in queries for t in q)
batch = sdb.PreparedBatch(reader, queries, sdb.OR, sdb.BM25(), 1000)
d_keys = torch.empty(nq * 1000, dtype=torch.int64, device="cuda:0")
ref = None
for rep in range(2):
    for wand in (0, 2):
        for lut in ("1", "0"):
            os.environ["SDBG_STREAM_LUT"] = lut
            ctx.set_wand(wand)
            h, nout, tot = batch.run_host()
            if ref is None:
                ref = (h.copy(), nout.copy())
            same = np.array_equal(h["doc"], ref[0]["doc"]) and np.array_equal(h["score"], ref[0]["score"]) and np.array_equal(nout, ref[1])
            batch.run_device(0, d_keys.data_ptr())
            ts = []
            for _ in range(5):
                ctx.flush_l2(); ctx.sync(); ctx.timer_start()
                batch.run_device(0, d_keys.data_ptr())
                ts.append(ctx.timer_stop())
            ms = float(np.median(ts))
            print("wand", wand, "lut", lut, "ms %.3f" % ms, "G postings/s %.1f" % (postings / ms / 1e6), "hits identical", same, flush=True)
