import os
import shutil
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import nsb200  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "ref: needs the compiled reference (oracle/_ref/ref_engine)")


def pytest_collection_modifyitems(config, items):
    skip_ref = pytest.mark.skip(reason="oracle/_ref/ref_engine not built (needs /root/reference)")
    for item in items:
        if "ref" in item.keywords and not orc.have_ref():
            item.add_marker(skip_ref)


@pytest.fixture(scope="session")
def workdir():
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    d = tempfile.mkdtemp(prefix="nsb200_test_", dir=base)
    yield d
    shutil.rmtree(d, ignore_errors=True)


class IndexCase:
    def __init__(self, path, spec, ndocs, nseg):
        self.path, self.spec, self.ndocs, self.nseg = path, spec, ndocs, nseg
        self._oracle = None

    @property
    def oracle(self):
        if self._oracle is None:
            self._oracle = orc.OracleIndex(self.path)
        return self._oracle


def make_case(workdir, name, spec, ndocs, nseg):
    path = os.path.join(workdir, name)
    if not os.path.isdir(path):
        nsb200.build_index(path, spec, ndocs, nseg)
    return IndexCase(path, spec, ndocs, nseg)


@pytest.fixture(scope="session")
def small_case(workdir):
    """2 ragged segments x 1500 docs, tiny vocabulary: many multi-term overlaps and ties."""
    return make_case(workdir, "small", nsb200.CorpusSpec(vocab=3000), 3000, 2)


@pytest.fixture(scope="session")
def config1_case(workdir):
    """BASELINE configs[0]: single segment, 10k docs, V=50k."""
    return make_case(workdir, "config1", nsb200.SPEC_10K, 10_000, 1)


EDGE_QUERIES = [
    "t3 t3",                 # duplicates are scored twice (src/api_engine.cpp:391-397)
    "t3 t5 t3 t5 t3",
    "the of and",            # all stopwords -> no "found"
    "",                      # empty
    "a b c",                 # all too short
    "zzzz",                  # unknown term -> found 0
    "zzzz t2",
    "T7, the t9!",           # case folding, punctuation, stopword
    "t1\tt2\nt4",
    "t1-t2_t3",
    "café t2",          # bytes >= 0x80 separate tokens
    "t1 t2 t3 t4 t5 t6 t7 t8 t9 t10 t11 t12",
]


def assert_same_as_oracle(res, oracle_index, queries, k):
    """Bit-exact comparison of a BatchResult with the oracle (scores as u32 bit patterns)."""
    _, s, g, d, nh, fo, hf = oracle_index.search_many(queries, k, nthreads=4)
    assert np.array_equal(res.nhits, nh)
    assert np.array_equal(res.found, fo)
    assert np.array_equal(res.has_found, hf)
    for q in range(len(queries)):
        n = int(nh[q])
        assert np.array_equal(res.hits["score"][q, :n].view(np.uint32), s[q, :n].view(np.uint32)), queries[q]
        assert np.array_equal(res.hits["seg"][q, :n], g[q, :n]), queries[q]
        assert np.array_equal(res.hits["doc"][q, :n], d[q, :n]), queries[q]
