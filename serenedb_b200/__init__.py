"""serenedb_b200 -- B200-native (sm_100a) implementation of SereneDB's query-time hot path:
IResearch BM25 posting scan + top-k, and the `iresearch_scan` columnar filter -> aggregate.

The package is a thin host layer over libsdbg.so (hand-written CUDA behind the C ABI in
include/sdbg.h). Importing it loads the native library; if the library cannot be built or loaded the
import raises -- there is no Python or CPU fallback for any operation.
"""
from . import _native
from .engine import (AND, OR, BM25, TFIDF, FLT_MIN, Context, ExecuteTopK, ExecuteTopKBatch, IndexReader,
                     IResearchScan, PostingsWriter, PreparedBatch, Segment, merge_gathered, pred,
                     stage_parse_host, sum_i128, StreamScoredDocs, pack_for)

_native.lib()  # fail loudly at import time when the CUDA extension is missing

__all__ = ["AND", "OR", "BM25", "TFIDF", "FLT_MIN", "Context", "ExecuteTopK", "ExecuteTopKBatch", "IndexReader",
           "IResearchScan", "PostingsWriter", "PreparedBatch", "Segment", "merge_gathered", "pred",
           "stage_parse_host", "sum_i128", "StreamScoredDocs", "pack_for"]
