"""Engine::reload() on damaged or unusual index directories (src/api_engine.cpp:50-90, src/api_segment.cpp:14-136):
the product's loader and the oracle's accept and refuse exactly what the reference does.  EXPECT was recorded from
the reference itself (oracle/_ref/ref_engine); the `ref` test re-runs it where the binary exists."""
import json
import os
import subprocess

import pytest

import fmt
import nsb200
from oracle import oracle as orc


def _base(idx):
    seg = os.path.join(idx, "segments", "seg_000001")
    fmt.write_segment(seg, fmt.handmade_docs())
    fmt.write_manifest(idx, ["seg_000001"])
    return seg


def _rm(seg, name):
    os.remove(os.path.join(seg, name))


def _make(case, idx):
    if case == "zero_doc_segment":
        fmt.write_segment(os.path.join(idx, "segments", "seg_000001"), [])
        fmt.write_manifest(idx, ["seg_000001"])
        return
    if case == "empty_index_dir":
        os.makedirs(idx)
        return
    seg = _base(idx)
    if case.startswith("missing_"):
        _rm(seg, case[len("missing_"):] + ".bin")
    elif case == "segment_listed_twice":
        fmt.write_manifest(idx, ["seg_000001", "seg_000001"])
    elif case == "manifest_names_a_missing_segment":
        fmt.write_manifest(idx, ["seg_000001", "seg_000009"])
    elif case == "empty_manifest":
        fmt.write_manifest(idx, [])
    elif case == "no_manifest":
        os.remove(os.path.join(idx, "manifest.bin"))
    elif case == "dir_scan_skips_other_dirs":
        os.makedirs(os.path.join(idx, "segments", "notes"))
        os.remove(os.path.join(idx, "manifest.bin"))
    elif case == "dir_scan_skips_a_seg_file":
        open(os.path.join(idx, "segments", "seg_readme"), "w").write("x")
        os.remove(os.path.join(idx, "manifest.bin"))
    elif case != "ok":
        raise AssertionError(case)


# case -> (reload succeeds, segments, found for "alpha")
EXPECT = {
    "ok": (True, 1, 6),
    "missing_docs": (False, 0, None),
    "missing_stats": (False, 0, None),
    "missing_lexicon_b010": (False, 0, None),          # every lexicon barrel must open (api_segment.cpp:84-86)
    "missing_inverted_b063": (False, 0, None),         # every inverted barrel must open (:75-79)
    "missing_inverted_b000": (False, 0, None),         # has_barrels() false -> legacy loader -> no lexicon.bin
    "missing_barrels": (False, 0, None),
    "segment_listed_twice": (True, 2, 12),             # loaded twice, scored twice
    "manifest_names_a_missing_segment": (False, 0, None),
    "empty_manifest": (True, 1, 6),                    # falls back to the directory scan (api_engine.cpp:58-70)
    "no_manifest": (True, 1, 6),
    "zero_doc_segment": (True, 1, 0),
    "empty_index_dir": (False, 0, None),
    "dir_scan_skips_other_dirs": (True, 1, 6),
    "dir_scan_skips_a_seg_file": (True, 1, 6),
}


@pytest.mark.parametrize("case", sorted(EXPECT))
def test_reload_accepts_and_refuses_like_the_reference(workdir, case):
    ok, nseg, found = EXPECT[case]
    idx = os.path.join(workdir, "reload_case_" + case)
    _make(case, idx)
    e = nsb200.Engine(idx, device=None)
    assert e.reload() is ok, (case, e.last_error)
    if ok:
        assert e.num_segments == nseg
    else:
        assert e.last_error                              # the refusal says why
    e.close()
    if ok:
        oi = orc.OracleIndex(idx)
        assert oi.num_segments == nseg
        r = oi.search("alpha", 10)
        assert r["found"] == found and r["segments"] == nseg
    else:
        with pytest.raises(RuntimeError):
            orc.OracleIndex(idx)


@pytest.mark.ref
@pytest.mark.parametrize("case", sorted(EXPECT))
def test_expectations_are_the_live_references(workdir, case):
    ok, nseg, found = EXPECT[case]
    idx = os.path.join(workdir, "reload_ref_" + case)
    _make(case, idx)
    qf = os.path.join(workdir, "reload_ref_q.txt")
    open(qf, "w").write("alpha\n")
    out = os.path.join(workdir, f"reload_ref_{case}.jsonl")
    r = subprocess.run([orc.REF_ENGINE, "search", idx, qf, "10", out], capture_output=True, text=True)
    assert (r.returncode == 0) is ok, (case, r.stderr[-200:])
    if ok:
        row = json.loads(open(out).readline())
        assert row["segments"] == nseg and row.get("found") == found


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["segment_listed_twice", "zero_doc_segment"])
def test_unusual_but_loadable_indexes_on_the_gpu(workdir, case):
    """A segment listed twice is uploaded and scored twice (two global segment ids, found doubles); a segment
    without documents loads and answers nothing."""
    from conftest import assert_same_as_oracle

    idx = os.path.join(workdir, "reload_gpu_" + case)
    _make(case, idx)
    eng = nsb200.Engine(idx, device=0)
    assert eng.reload(), eng.last_error
    oi = orc.OracleIndex(idx)
    for k in (10, 3):
        assert_same_as_oracle(eng.search_batch(fmt.HANDMADE_QUERIES, k), oi, fmt.HANDMADE_QUERIES, k)
    eng.close()
