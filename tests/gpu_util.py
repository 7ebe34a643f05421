"""Helpers shared by the GPU parity tests: build the same segment in the oracle and in HBM."""
import numpy as np

import orc
import serenedb_b200 as sdb
from serenedb_b200.engine import TERM_META_DTYPE

_ctx = None


def ctx():
    global _ctx
    if _ctx is None:
        _ctx = sdb.Context(0)
        _ctx.set_wand(False)   # exact total_matches (ExecuteTopKWithCount semantics); tests/test_gpu_wand.py turns pruning on
    return _ctx


def metas_of(oseg):
    return np.array([(m.docs_count, m.freq, m.doc_start, m.e_skip_start) for m in oseg.term_metas()],
                    dtype=TERM_META_DTYPE)


def to_gpu(oseg, has_wand=True, columns=None):
    """Stage an oracle segment's .doc bytes, norms and columns into a GPU segment."""
    g = sdb.Segment(ctx(), oseg.n_docs)
    if oseg.num_terms():
        g.stage_postings(oseg.doc_bytes(), metas_of(oseg), has_wand=has_wand)
    if oseg.has_norms:
        nb, w = oseg.norm_bytes()
        g.stage_norms(nb, w)
    for field, (vals, validity) in (columns or {}).items():
        g.stage_column(field, vals, validity)
    return g


def oracle_terms(reader, scorer, term_ids):
    """orc.BM25Term list with the same statistics the GPU path derives."""
    out = []
    for t in term_ids:
        s = reader.stats(scorer, t)
        q = orc.BM25Term()
        q.idf, q.norm_const, q.norm_length, q.boost, q.term = s.idf, s.norm_const, s.norm_length, s.boost, t
        out.append(q)
    return out


def assert_hits_equal(gpu_hits, orc_hits):
    assert len(gpu_hits) == len(orc_hits), (len(gpu_hits), len(orc_hits))
    assert np.array_equal(gpu_hits["doc"], orc_hits["doc"])
    assert np.array_equal(gpu_hits["seg"], orc_hits["seg"])
    # bit-exact fp32 scores (same op order, no FMA contraction on either side)
    assert np.array_equal(gpu_hits["score"].view(np.uint32), orc_hits["score"].view(np.uint32))
