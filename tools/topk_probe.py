"""Diagnostic: BM25 batch time against the chain-split granularity (SDBG_TOPK_CHAIN_DIV: target postings per
chain = batch postings / (SMs * DIV)) and the window budget. Corpus built once; env re-read on every call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import serenedb_b200 as sdb
import bench

ctx = sdb.Context(0)
n = int(os.environ.get("PROBE_DOCS", 10_000_000))
g = sdb.Segment(ctx, n)
dc, sum_dl = g.synth_corpus(0, 0, 256, threads=32)
reader = sdb.IndexReader([g], n, sum_dl, dc)
scorer = sdb.BM25()
queries = bench.make_queries(4096)
postings = sum(int(dc[t]) for q in queries for t in q)
batch = sdb.PreparedBatch(reader, queries, sdb.OR, scorer, 1000)
d_keys = torch.empty(len(queries) * 1000, dtype=torch.int64, device="cuda:0")
ref = None
for env in ({"SDBG_TOPK_CHAIN_DIV": "4"}, {"SDBG_TOPK_CHAIN_DIV": "1"}, {"SDBG_TOPK_CHAIN_DIV": "2"}, {"SDBG_TOPK_CHAIN_DIV": "3"},
            {"SDBG_TOPK_CHAIN_DIV": "6"}, {"SDBG_TOPK_CHAIN_DIV": "16"}, {"SDBG_TOPK_CHAIN_DIV": "2", "SDBG_TOPK_BUDGET": "16"}):
    for k_ in ("SDBG_TOPK_CHAIN_DIV", "SDBG_TOPK_BUDGET", "SDBG_TOPK_CAP"):
        os.environ.pop(k_, None)
    os.environ.update(env)
    h, nout, tot = batch.run_host()
    if ref is None:
        ref = (h.copy(), nout.copy())
    else:
        assert np.array_equal(nout, ref[1]) and np.array_equal(h["doc"], ref[0]["doc"]) and np.array_equal(h["score"], ref[0]["score"]), env
    batch.run_device(0, d_keys.data_ptr())
    ctx.flush_l2(); ctx.sync(); ctx.timer_start()
    for _ in range(3):
        batch.run_device(0, d_keys.data_ptr())
    ms = ctx.timer_stop() / 3
    print(env, "ms", round(ms, 3), "G postings/s", round(postings / ms / 1e6, 1), flush=True)
