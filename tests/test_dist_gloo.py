"""world_size-2 gloo tests (CPU) of the N>1 host logic: shard -> partial -> one collective -> merge
gives exactly the unsharded answer (oracle partials stand in for the GPU kernels here; the GPU
versions of the same flows run in bench.py --gpus N)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

K, A, B, V, W = 10, 11, 12, 13, 14
ROWS = 200_000
SPAN = 100000
N_DOCS = 120_000
TERMS = [0, 3, 9, 30]
TOPK = 50


def _table(row0, rows):
    import orc
    seg = orc.Segment(rows, has_wand=False)
    for f, (stream, kind) in {K: (10, 0), A: (11, 1), B: (12, 2), V: (13, 3), W: (14, 4)}.items():
        seg.add_column(f, orc.synth_column(stream, kind, row0, rows))
    return seg


def _preds():
    import orc
    return [orc.make_pred(A, "LT", 500000), orc.make_pred(B, "GE", 0.25, is_float=True)]


def _partials(groups):
    """Oracle group rows -> the flat buffers sdbg_filter_groupby_partial produces (wide limbs)."""
    i64 = np.zeros(4 * SPAN, np.int64)
    f64 = np.zeros(SPAN, np.float64)
    for g in groups:
        k = int(g["key"])
        total = (int(g["sum_hi"]) << 64) + (int(g["sum_lo"]) & 0xFFFFFFFFFFFFFFFF)   # signed 128-bit value
        i64[k] = int(g["count"])
        i64[SPAN + k] = total & 0xFFFFFFFF              # low limb: sum of v & 0xFFFFFFFF (non-negative)
        i64[2 * SPAN + k] = (total - (total & 0xFFFFFFFF)) >> 32
        i64[3 * SPAN + k] = int(g["cnt_f64"])
        f64[k] = float(g["sum_f64"])
    return i64, f64


def _wide_table(shard):
    """Sparse int64 keys shared between shards, values that need the full 128-bit SUM, doubles of mixed sign."""
    import orc
    rng = np.random.default_rng(100 + shard)
    rows = 30_000 + 17 * shard
    pool = np.random.default_rng(5).integers(-2**62, 2**62, size=400)          # same key pool on every shard
    key = pool[rng.integers(0, len(pool), size=rows)].astype(np.int64)
    v = rng.integers(-2**62, 2**62, size=rows).astype(np.int64)
    a = rng.integers(0, 100, size=rows).astype(np.int64)
    w = rng.standard_normal(rows)
    seg = orc.Segment(rows, has_wand=False)
    for f, vals in {1: key, 2: v, 3: a, 4: w}.items():
        seg.add_column(f, vals)
    return seg, rows


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import orc
    from serenedb_b200 import dist as sd
    res = {}
    # ---- GROUP BY: shard rows, oracle partial per rank, one SUM all-reduce per dtype ----
    lo, hi = sd.shard_rows(ROWS, rank, world)
    part = orc.filter_groupby([_table(lo, hi - lo)], _preds(), K, V, W, cap=SPAN + 1)
    i64, f64 = _partials(part)
    ti, tf = torch.from_numpy(i64), torch.from_numpy(f64)
    sd.merge_groupby_partials(dist, ti, tf)
    res["gb_i64"], res["gb_f64"] = ti.numpy().copy(), tf.numpy().copy()
    # ---- BM25: shard docs, global statistics, per-rank top-k keys, one all-gather, local select ----
    per = N_DOCS // world
    seg, dl, lists = orc.synth_segment(per, TERMS, doc0=rank * per)
    dwt = torch.tensor([len(d) for d, _ in lists], dtype=torch.int64)
    ttf = torch.tensor([int(dl.sum())], dtype=torch.int64)
    dwf = torch.tensor([per], dtype=torch.int64)
    sd.global_term_stats(dist, dwt, ttf, dwf)
    res["stats"] = (dwt.numpy().copy(), int(ttf.item()), int(dwf.item()))
    queries = [[0, 2], [1, 3], [2]]
    keys = np.zeros(len(queries) * TOPK, np.int64)
    for qi, qt in enumerate(queries):
        terms = []
        for t in qt:
            st = orc.bm25_stats(int(dwf.item()), int(ttf.item()), int(dwt[t]))
            x = orc.BM25Term()
            x.idf, x.norm_const, x.norm_length, x.boost, x.term = st.idf, st.norm_const, st.norm_length, 1.0, t
            terms.append(x)
        hits, _, _ = orc.bm25_topk([seg], "OR", terms, TOPK, mode=1)
        for i, h in enumerate(hits):
            keys[qi * TOPK + i] = np.uint64(sd.rebase_key(sd.make_key(h["score"], h["doc"]), rank)).view(np.int64)
    gathered = sd.gather_topk_keys(dist, torch.from_numpy(keys))
    res["topk"] = sd.select_topk_host(gathered.numpy(), world, len(queries), TOPK)
    # ---- hash-table GROUP BY (wide keys): group rows per shard, one object all-gather, key-wise merge ----
    wseg, _ = _wide_table(rank)
    wrows = orc.filter_groupby([wseg], [orc.make_pred(3, "LT", 70)], 1, 2, 4, cap=5000)
    res["wide"] = sd.merge_group_rows(dist, wrows)
    if rank == 0:
        q.put(res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.fixture(scope="module")
def two_rank_result():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_groupby_sharded_equals_unsharded(two_rank_result):
    import orc
    full = orc.filter_groupby([_table(0, ROWS)], _preds(), K, V, W, cap=SPAN + 1)
    i64, f64 = two_rank_result["gb_i64"], two_rank_result["gb_f64"]
    live = np.nonzero(i64[:SPAN])[0]
    assert np.array_equal(live, full["key"])
    assert np.array_equal(i64[live], full["count"].astype(np.int64))
    tot = [(int(i64[2 * SPAN + k]) << 32) + int(i64[SPAN + k]) for k in live]
    exp = [(int(h) << 64) + (int(l) & 0xFFFFFFFFFFFFFFFF) for l, h in zip(full["sum_lo"], full["sum_hi"])]
    assert tot == exp                                   # SUM(int) exact through the limb all-reduce
    assert np.array_equal(i64[3 * SPAN + live], full["cnt_f64"].astype(np.int64))
    assert np.allclose(f64[live] / i64[3 * SPAN + live], full["sum_f64"] / full["cnt_f64"], rtol=1e-5)


def test_topk_sharded_equals_unsharded(two_rank_result):
    import orc
    from serenedb_b200 import dist as sd
    seg, dl, lists = orc.synth_segment(N_DOCS, TERMS, doc0=0)
    dwt, ttf, dwf = two_rank_result["stats"]
    assert list(dwt) == [len(d) for d, _ in lists] and ttf == int(dl.sum()) and dwf == N_DOCS
    per = N_DOCS // 2
    for qi, qt in enumerate([[0, 2], [1, 3], [2]]):
        terms = []
        for t in qt:
            st = orc.bm25_stats(N_DOCS, ttf, int(dwt[t]))
            x = orc.BM25Term()
            x.idf, x.norm_const, x.norm_length, x.boost, x.term = st.idf, st.norm_const, st.norm_length, 1.0, t
            terms.append(x)
        hits, _, _ = orc.bm25_topk([seg], "OR", terms, TOPK, mode=1)
        got = [sd.split_key(k) for k in two_rank_result["topk"][qi]]
        assert len(got) == len(hits)
        for (score, rank, ordinal), h in zip(got, hits):
            assert np.float32(score) == h["score"]
            assert rank * per + ordinal == h["doc"]     # shards are doc ranges: global doc = rank*per + local


def test_hash_groupby_rows_merge(two_rank_result):
    """Sparse-key GROUP BY: per-shard group rows merged across ranks == one GROUP BY over both shards."""
    import orc
    segs = [_wide_table(r)[0] for r in range(2)]
    full = orc.filter_groupby(segs, [orc.make_pred(3, "LT", 70)], 1, 2, 4, cap=5000)
    got = two_rank_result["wide"]
    for f in ("key", "count", "sum_lo", "sum_hi", "cnt_f64"):
        assert np.array_equal(got[f], full[f]), f
    assert np.allclose(got["sum_f64"], full["sum_f64"], rtol=1e-9, atol=1e-9)
