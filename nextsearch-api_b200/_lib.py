"""ctypes binding of include/nextsearch_b200.h (libnsb200.so).

The shared library is built in-tree by ``__graft_entry__.build()`` /
``make -C nextsearch-api_b200/csrc``.  There is no Python or CPU implementation of the
search path behind this module: if the library is missing, importing fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NSB200_LIB") or os.path.join(_HERE, "libnsb200.so")  # NSB200_LIB: an experimental build

NS_OK = 0
NS_MAX_K = 100
NS_MAX_TERMS = 256
NS_SEG_DROP_RAW = 1
NS_IPC_HANDLE_BYTES = 64
STATUS_NAMES = {0: "NS_OK", 1: "NS_ERR_INVALID", 2: "NS_ERR_CUDA", 3: "NS_ERR_IO", 4: "NS_ERR_FORMAT",
                5: "NS_ERR_NOMEM", 6: "NS_ERR_STATE"}


class NsError(RuntimeError):
    def __init__(self, status: int, text: str):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {text}")
        self.status = status


class QTerm(C.Structure):
    _fields_ = [("seg", C.c_uint32), ("row", C.c_uint32), ("idf", C.c_float), ("weight", C.c_float)]


class Hit(C.Structure):
    _fields_ = [("score", C.c_float), ("seg", C.c_uint32), ("doc", C.c_uint32)]


class CorpusSpecC(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("vocab", C.c_uint32), ("zipf_s", C.c_double), ("zipf_q", C.c_double),
                ("len_lo", C.c_uint32), ("len_hi", C.c_uint32)]


# every symbol include/nextsearch_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_u8p = C.POINTER(C.c_uint8)
_strs = C.POINTER(C.c_char_p)
SYMBOLS = {
    "ns_last_error": (C.c_char_p, []),
    "ns_device_count": (C.c_int, []),
    "ns_index_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "ns_index_destroy": (None, [_P]),
    "ns_index_add_segment": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_float, _P, C.c_uint32, _P, _P, _P, C.c_uint64]),
    "ns_index_add_segment_ex": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_float, _P, C.c_uint32, _P, _P, _P, _P, C.c_uint64,
                                          C.c_uint32]),
    "ns_upload_begin": (C.c_int, [_P, C.c_uint64, C.POINTER(_P)]),
    "ns_upload_buffer": (_P, [_P]),
    "ns_upload_push": (C.c_int, [_P, C.c_uint64, C.c_uint64]),
    "ns_upload_finish": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_float, _P, C.c_uint32, _P, _P, _P, C.c_uint32]),
    "ns_upload_abort": (None, [_P]),
    "ns_index_commit": (C.c_int, [_P]),
    "ns_index_abort": (C.c_int, [_P]),
    "ns_index_num_segments": (C.c_int, [_P]),
    "ns_index_device_bytes": (C.c_uint64, [_P]),
    "ns_search_batch": (C.c_int, [_P, C.c_uint32, C.c_int, _P, _P, _P, _P, _P]),
    "ns_batch_prepare": (C.c_int, [_P, C.c_uint32, C.c_int, _P, _P, C.POINTER(_P)]),
    "ns_batch_launch": (C.c_int, [_P, _P]),
    "ns_batch_sync": (C.c_int, [_P]),
    "ns_batch_fetch": (C.c_int, [_P, _P, _P, _P]),
    "ns_batch_destroy": (None, [_P]),
    "ns_batch_device_results": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    "ns_batch_posting_count": (C.c_uint64, [_P]),
    "ns_batch_num_launches": (C.c_uint32, [_P]),
    "ns_batch_last_kernel_ms": (C.c_float, [_P, C.c_int]),
    "ns_batch_set_splits": (C.c_int, [_P, C.c_uint32]),
    "ns_batch_stream": (_P, [_P]),
    "ns_batch_upload_bytes": (C.c_uint64, [_P]),
    "ns_batch_result_bytes": (C.c_uint64, [_P]),
    "ns_exchange_create": (C.c_int, [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(_P)]),
    "ns_exchange_destroy": (None, [_P]),
    "ns_exchange_ipc_handle": (C.c_int, [_P, _P]),
    "ns_exchange_attach_ipc": (C.c_int, [_P, C.c_uint32, _P]),
    "ns_exchange_attach_local": (C.c_int, [_P, _P]),
    "ns_batch_launch_exchange": (C.c_int, [_P, _P, C.c_uint64, _P]),
    "ns_exchange_merge": (C.c_int, [_P, C.c_uint64, C.c_uint32, C.c_int, C.c_int, _P]),
    "ns_exchange_result_device": (C.c_int, [_P, C.c_uint64, C.POINTER(_P)]),
    "ns_exchange_fetch": (C.c_int, [_P, C.c_uint64, C.c_uint32, C.c_int, _P, _P, _P]),
    "ns_merge_device": (C.c_int, [C.c_int, C.c_uint32, C.c_int, C.c_uint32, _P, _P, _P, _P, _P, _P, _P]),
    "ns_batch_result_blob": (C.c_int, [_P, C.POINTER(_P), _u64p, _u64p, _u64p]),
    "ns_merge_blobs_device": (C.c_int, [C.c_int, C.c_uint32, C.c_int, C.c_uint32, _P, C.c_uint64, C.c_uint64, C.c_uint64,
                                        _P, _P, _P, _P]),
    "ns_semantic_upload": (C.c_int, [C.c_int, C.c_uint32, C.c_uint32, _P, C.POINTER(_P)]),
    "ns_semantic_destroy": (None, [_P]),
    "ns_semantic_scan": (C.c_int, [_P, C.c_uint32, _P, C.c_float, C.c_uint32, _P, _P, _P]),
    "ns_selftest_fastdiv": (C.c_int, [C.c_int, C.c_uint64, C.c_uint64, _u64p]),
    "ns_debug_violations": (C.c_int, [C.c_int, _u64p, C.c_int]),
    "ns_debug_selftest": (C.c_int, [C.c_int]),
    "ns_engine_create": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(_P)]),
    "ns_engine_create_multi": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(_P)]),
    "ns_engine_num_devices": (C.c_int, [_P]),
    "ns_engine_device_index": (_P, [_P, C.c_int]),
    "ns_engine_reload_stats": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), _u64p, _u64p]),
    "ns_engine_search_terms_batch": (C.c_int, [_P, C.c_uint32, _P, _strs, _P, C.c_int, _P, _P, _P, _P]),
    "ns_engine_expand": (C.c_int, [_P, C.c_char_p, C.c_char_p, C.c_size_t, _P, C.c_int, C.POINTER(C.c_int)]),
    "ns_engine_search_one": (C.c_int, [_P, C.c_char_p, C.c_int, _P, _u32p, _u64p, _u8p]),
    "ns_engine_coalescer_start": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_int]),
    "ns_engine_coalescer_stop": (C.c_int, [_P]),
    "ns_engine_coalescer_stats": (C.c_int, [_P, _u64p, _u64p, _u64p]),
    "ns_engine_load_test": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_char_p, C.c_size_t, C.c_int,
                                      C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "ns_engine_destroy": (None, [_P]),
    "ns_engine_set_shard": (C.c_int, [_P, C.c_int, C.c_int]),
    "ns_engine_reload": (C.c_int, [_P]),
    "ns_engine_num_segments": (C.c_int, [_P]),
    "ns_engine_segment_name": (C.c_int, [_P, C.c_int, C.c_char_p, C.c_size_t]),
    "ns_engine_segment_stats": (C.c_int, [_P, C.c_int, _u32p, C.POINTER(C.c_float), _u32p, _u64p]),
    "ns_engine_term_stats": (C.c_int, [_P, C.c_int, C.c_char_p, _u32p, _u32p]),
    "ns_engine_search_json": (C.c_int, [_P, C.c_char_p, C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "ns_engine_search_batch": (C.c_int, [_P, C.c_uint32, _strs, C.c_int, _P, _P, _P, _P]),
    "ns_engine_resolve_batch_packed": (C.c_int, [_P, C.c_uint32, C.c_char_p, C.c_size_t, _P, _P, C.c_uint64, _u64p, _P]),
    "ns_engine_search_batch_packed": (C.c_int, [_P, C.c_uint32, C.c_char_p, C.c_size_t, C.c_int, _P, _P, _P, _P]),
    "ns_engine_resolve_batch": (C.c_int, [_P, C.c_uint32, _strs, _P, _P, C.c_uint64, _u64p, _P]),
    "ns_engine_prepare_batch_packed": (C.c_int, [_P, C.c_uint32, C.c_char_p, C.c_size_t, C.c_int, C.POINTER(_P), _P]),
    "ns_engine_index": (_P, [_P]),
    "ns_engine_cord_uid": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_char_p, C.c_size_t]),
    "ns_text_query_terms": (C.c_int, [C.c_char_p, C.c_char_p, C.c_size_t]),
    "ns_corpus_write_segment": (C.c_int, [C.POINTER(CorpusSpecC), C.c_uint64, C.c_uint32, C.c_char_p, C.c_int,
                                          C.c_char_p, C.c_int]),
    "ns_corpus_write_manifest": (C.c_int, [C.c_char_p, C.c_uint32, _strs]),
    "ns_corpus_make_queries": (C.c_int, [C.POINTER(CorpusSpecC), C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
                                         C.c_uint32, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
}

_lib = None


def load() -> C.CDLL:
    """Load libnsb200.so and type every exported function.  Raises if the build is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no fallback implementation)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != NS_OK:
        raise NsError(status, (load().ns_last_error() or b"").decode("utf-8", "replace"))
