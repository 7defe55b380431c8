"""The memcheck substitute: the GPU workloads of tools/debug_checks_probe.py against the DEBUG library
(libnsb200_dbg.so, -DNSB_DEBUG_CHECKS), whose score kernel checks every index it derives from a posting, a tile table
or a descriptor before using it.  compute-sanitizer is closed on the pool this was developed on ("find a bad access
with bounds checks and asserts of your own"); this is that."""
import ctypes as C
import json
import os
import subprocess
import sys

import pytest

import nsb200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DBG = os.path.join(ROOT, "nextsearch-api_b200", "libnsb200_dbg.so")


def test_product_library_has_no_debug_counters():
    """The product build compiles none of the checks; asking it for the counters is an error, not a zero."""
    lib = nsb200._lib.load()
    counts = (C.c_uint64 * 8)()
    assert lib.ns_debug_violations(0, counts, 8) == 6      # NS_ERR_STATE
    assert lib.ns_debug_selftest(0) == 6


@pytest.mark.gpu
def test_debug_library_counts_no_index_violations():
    if not os.path.exists(DBG):
        pytest.skip("debug library not built (make -C nextsearch-api_b200/csrc debug)")
    env = dict(os.environ, NSB200_LIB=DBG)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "debug_checks_probe.py")], env=env, cwd=ROOT,
                       capture_output=True, text=True, timeout=240)
    assert r.stdout.strip(), r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["debug_build"] and line["lib"] == "libnsb200_dbg.so", line
    assert line["parity_ok"], line["parity"]
    assert line["counters_live"], line
    assert sum(line["violations"].values()) == 0, line["violations"]
    assert r.returncode == 0
