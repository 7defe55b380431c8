"""bench.py pieces that do not need a GPU: the workload description, the committed ncu traffic figure the
roofline object quotes, and the measured-peak lookup."""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_workload_config_names_the_baseline_configs():
    bench = importlib.import_module("bench")
    one = bench.workload_config(1, 1)
    many = bench.workload_config(8, 8)
    assert "configs[1]" in one["workload"] and one["segments"] == 1 and one["batch_queries"] == 4096 and one["k"] == 10
    assert "configs[2]" in many["workload"] and many["parallelism"] == "segments%8"
    assert "model" not in one and "l2_policy" in one


def test_roofline_traffic_comes_from_the_committed_ncu_capture():
    bench = importlib.import_module("bench")
    assert os.path.exists(bench.ncu_summary_path()), "bench.py quotes a profile that is not committed"
    t = bench.ncu_traffic()
    # DRAM read+write of one score-kernel launch: well below the 6.11 GB of algorithmic bytes (L2 sharing)
    assert t is not None and 1e8 < t < 6.11e9


def test_measured_peak_is_the_driver_file_or_the_stated_fallback():
    bench = importlib.import_module("bench")
    peak, src = bench.measured_peak()
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        assert peak == float(json.load(open(p))["hbm_gbs"]) and src.startswith("measured")
    else:
        assert peak == 6650.0 and src.startswith("fallback")
