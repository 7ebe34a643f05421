"""Multi-GPU plumbing: one process per GPU, shards = index segments / row ranges, and exactly one
collective step per query batch. The reference has no distributed layer; its nearest
analogues are the per-segment statistics merge in PreparePhase and the atomic cross-thread threshold
(server/connector/duckdb_search_full_scan.cpp:1359-1383, 1886-1920).

Works with any torch.distributed backend: NCCL over NVLink on the GPU box, gloo on CPU in the tests.
The tensors handed in are the flat buffers the C ABI fills (sdbg_filter_groupby_partial,
sdbg_bm25_topk_batch_device), so the collective touches no intermediate copies.
"""
import numpy as np

RANK_SLOT_BITS = 28  # each rank owns 2^28 ordinals in the merged key space (sdbg_bm25_topk_batch_device)


def global_term_stats(dist, docs_with_term, total_term_freq, docs_with_field):
    """Corpus-wide BM25 statistics = sums over shards (collectors.cpp:36-52). Depends only on the
    snapshot, so it runs once at index-build time, not per query. Tensors are reduced in place."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(docs_with_term)
        dist.all_reduce(total_term_freq)
        dist.all_reduce(docs_with_field)
    return docs_with_term, total_term_freq, docs_with_field


def merge_groupby_partials(dist, part_i64, part_f64):
    """SUM all-reduce of dense partial aggregates. part_i64 = [count | sum_lo | sum_hi | cnt_f64]
    (4*span int64). torch.distributed fallback of sdbg_dist_groupby_merge (the C path, which also normalises the
    SUM(int) limbs to lo < 2^32 before the all-reduce so that no number of ranks can wrap them). Here the limbs travel
    as the kernels left them: exact while world_size * rows_per_gpu < 2^31; part_f64 = float64 sums (order of addition differs from one GPU: AVG
    agrees to ~1e-15 relative, far inside the 1e-5 bar)."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(part_i64)
        dist.all_reduce(part_f64)
    return part_i64, part_f64


def combine_group_rows(rows):
    """Merge finalized group rows (engine.GROUP_DTYPE) that may repeat a key: COUNT / cnt add, SUM(double) adds,
    SUM(int) adds as a 128-bit two's-complement value carried in (sum_lo unsigned, sum_hi signed).
    Returns rows sorted by key, one per distinct key."""
    rows = np.asarray(rows)
    if len(rows) == 0:
        return rows.copy()
    r = rows[np.argsort(rows["key"], kind="stable")]
    keys, start = np.unique(r["key"], return_index=True)
    out = np.zeros(len(keys), rows.dtype)
    out["key"] = keys
    out["count"] = np.add.reduceat(r["count"], start)
    out["cnt_f64"] = np.add.reduceat(r["cnt_f64"], start)
    out["sum_f64"] = np.add.reduceat(r["sum_f64"], start)
    lo = r["sum_lo"].view(np.uint64)
    m32 = np.uint64(0xFFFFFFFF)
    lo_l = np.add.reduceat(lo & m32, start)                 # 32-bit halves: exact in uint64 below 2^32 rows per key
    lo_h = np.add.reduceat(lo >> np.uint64(32), start)
    lo_h = lo_h + (lo_l >> np.uint64(32))
    carry = lo_h >> np.uint64(32)
    out["sum_lo"] = (((lo_h & m32) << np.uint64(32)) | (lo_l & m32)).view(np.int64)
    out["sum_hi"] = np.add.reduceat(r["sum_hi"], start) + carry.astype(np.int64)
    return out


def merge_group_rows(dist, rows):
    """Cross-rank merge for the hash-table GROUP BY (key ranges too wide for a dense partial, so there is
    nothing flat to all-reduce): every rank contributes the group rows of its shard, one all-gather of the
    (small, variable-length) results, and the same key-wise merge on every rank."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        gathered = [None] * dist.get_world_size()
        dist.all_gather_object(gathered, np.asarray(rows))
        rows = np.concatenate(gathered)
    return combine_group_rows(rows)


def gather_topk_keys(dist, keys, out=None):
    """All-gather of every rank's k best sortable keys per query ([Q*k] int64 each) -> [world, Q*k]."""
    import torch
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    if out is None:
        out = torch.empty(world * keys.numel(), dtype=keys.dtype, device=keys.device)
    if world == 1:
        out.copy_(keys)
    else:
        dist.all_gather_into_tensor(out, keys)
    return out


def shard_rows(total_rows, rank, world):
    """Contiguous row-range shard of a table (row-group-unit claiming, duckdb_search_full_scan.cpp:1293-1327)."""
    per = (total_rows + world - 1) // world
    lo = min(total_rows, rank * per)
    return lo, min(total_rows, lo + per)


def rebase_key(key, rank):
    """Host mirror of shift_keys_kernel: move a key's ordinal into the rank's slot."""
    key = int(key) & 0xFFFFFFFFFFFFFFFF
    if key == 0:
        return 0
    ordinal = (~key) & 0xFFFFFFFF
    ordinal += rank << RANK_SLOT_BITS
    return (key & 0xFFFFFFFF00000000) | ((~ordinal) & 0xFFFFFFFF)


def make_key(score, ordinal):
    bits = int(np.float32(score).view(np.uint32))
    return (bits << 32) | ((~int(ordinal)) & 0xFFFFFFFF)


def split_key(key):
    """-> (score fp32, rank, ordinal within rank)."""
    key = int(key) & 0xFFFFFFFFFFFFFFFF
    score = np.uint32(key >> 32).view(np.float32)
    ordinal = (~key) & 0xFFFFFFFF
    return float(score), ordinal >> RANK_SLOT_BITS, ordinal & ((1 << RANK_SLOT_BITS) - 1)


def select_topk_host(keys_all, world, nq, k):
    """CPU reference of sdbg_topk_merge_gathered: per query, the k largest non-zero keys over ranks."""
    a = np.asarray(keys_all, dtype=np.int64).view(np.uint64).reshape(world, nq, -1)
    out = []
    for q in range(nq):
        v = a[:, q, :].reshape(-1)
        v = v[v != 0]
        v = np.sort(v)[::-1][:k]
        out.append(v)
    return out
