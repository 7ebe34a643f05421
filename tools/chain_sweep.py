"""Batch time of the configs[2] workload (shipped pruning) over the work-split and candidate-buffer switches."""
import os, sys, itertools
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
ctx.set_wand(2)
ref = None
configs = [{}] + [{"SDBG_TOPK_CHAIN_DIV": str(d), "SDBG_TOPK_CHAIN_MIN": str(m)} for d, m in itertools.product((2, 4, 8, 16), (32768, 65536, 131072))] + \
          [{"SDBG_TOPK_CAP": "4096"}, {"SDBG_STREAM_LEAD_CHAINS": "8"}, {"SDBG_STREAM_LEAD_CHAINS": "32"}, {"SDBG_STREAM_LEAD": "0"}, {}]
for env in configs:
    for k_ in ("SDBG_TOPK_CHAIN_DIV", "SDBG_TOPK_CHAIN_MIN", "SDBG_TOPK_CAP", "SDBG_STREAM_LEAD_CHAINS", "SDBG_STREAM_LEAD"):
        os.environ.pop(k_, None)
    os.environ.update(env)
    h, nout, tot = batch.run_host()
    if ref is None:
        ref = (h.copy(), nout.copy())
    same = np.array_equal(h["doc"], ref[0]["doc"]) and np.array_equal(h["score"], ref[0]["score"])
    batch.run_device(0, d_keys.data_ptr())
    ts = []
    for _ in range(5):
        ctx.flush_l2(); ctx.sync(); ctx.timer_start()
        batch.run_device(0, d_keys.data_ptr())
        ts.append(ctx.timer_stop())
    ms = float(np.median(ts))
    print(env, "ms %.3f" % ms, "G postings/s %.1f" % (postings / ms / 1e6), "hits identical", same, flush=True)
