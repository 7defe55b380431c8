"""nextsearch-api_b200 — B200-native BM25 scoring + top-k for NextSearch's /api/search path.

Import through the repo-root shim ``import nsb200`` (the directory name has a hyphen), which
registers this package as ``nextsearch_api_b200``.
"""
from . import _lib
from .corpus import CorpusSpec, SPEC_10K, SPEC_1M, QUERY_SEED, build_index, make_queries, seg_name, write_manifest, write_segment
from .engine import (HIT_DTYPE, QTERM_DTYPE, Batch, BatchResult, DeviceIndex, Engine, Exchange, clamp_k, merge_device,
                     query_terms)

__all__ = [
    "_lib", "CorpusSpec", "SPEC_10K", "SPEC_1M", "QUERY_SEED", "build_index", "make_queries", "seg_name",
    "write_manifest", "write_segment", "HIT_DTYPE", "QTERM_DTYPE", "Batch", "BatchResult", "DeviceIndex", "Engine", "Exchange",
    "clamp_k", "merge_device", "query_terms",
]
