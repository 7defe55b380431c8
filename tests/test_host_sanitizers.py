"""The product's own HOST sources under AddressSanitizer + UBSan and under ThreadSanitizer (tests/tools/host_stress.cpp).
No GPU involved: a host-only engine loads segments and resolves queries; it never scores.

  * fuzzload (ASan + UBSan + LeakSanitizer): hundreds of indexes with ONE damaged file each (truncated, bytes flipped, a
    count field blown up, emptied).  Every reload ends in NS_OK or in an error code with a text — never in a crash, an
    out-of-bounds read, a huge allocation or an exception across the C ABI.  (The reference reads such files unchecked.)
  * stress (TSan): a reloader flips the index between two different corpora while worker threads resolve batches and
    read names / stats / uids: no data race, and every resolved batch is the answer of ONE corpus as a whole — the
    generation snapshot that makes `reload()` under load safe (the reference holds one mutex around everything)."""
import json
import os
import shutil
import subprocess

import pytest

import nsb200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "nextsearch-api_b200", "csrc")
DEVICE_OBJ = os.path.join(CSRC, "build", "device_api.o")
HOST_SRC = [os.path.join(CSRC, "host", f) for f in ("common.cpp", "segment_io.cpp", "corpus.cpp", "engine.cpp")]
CUDART = "/usr/local/cuda/lib64"

pytestmark = pytest.mark.skipif(
    shutil.which("g++") is None or not os.path.exists(DEVICE_OBJ) or not os.path.exists(os.path.join(CUDART, "libcudart_static.a")),
    reason="needs g++, the built device object (make -C nextsearch-api_b200/csrc) and the static CUDA runtime")


def _build(workdir, name, flags):
    out = os.path.join(workdir, name)
    cmd = ["g++", "-std=c++17", "-O1", "-g1", "-fno-omit-frame-pointer", "-ffp-contract=off", "-pthread", *flags, "-o", out,
           os.path.join(ROOT, "tests", "tools", "host_stress.cpp"), *HOST_SRC, DEVICE_OBJ, "-L" + CUDART, "-lcudart_static", "-ldl",
           "-lrt", "-lpthread"]
    return out, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)


@pytest.fixture(scope="module")
def drivers(workdir):
    """Both sanitizer builds, compiled side by side."""
    asan, pa = _build(workdir, "host_asan", ["-fsanitize=address,undefined"])
    tsan, pt = _build(workdir, "host_tsan", ["-fsanitize=thread"])
    for p in (pa, pt):
        log, _ = p.communicate(timeout=900)
        assert p.returncode == 0, log[-3000:]
    return {"asan": asan, "tsan": tsan}


@pytest.fixture(scope="module")
def corpora(workdir):
    a, b = os.path.join(workdir, "san_idx_a"), os.path.join(workdir, "san_idx_b")
    nsb200.build_index(a, nsb200.CorpusSpec(vocab=500), 600, 2)
    nsb200.build_index(b, nsb200.CorpusSpec(vocab=700, seed=5), 800, 3)
    # result decoration and expansion data, so that the fuzz damages their parsers' inputs as well
    import random
    rng = random.Random(4)
    with open(os.path.join(a, "metadata.csv"), "w", newline="") as f:
        f.write("cord_uid,sha,source_x,title,doi,publish_time,authors,journal,url\n")
        for d in range(0, 600, 3):
            f.write(f'uid{d},s{d},PMC,"Title, {d} ""quoted""",10.1/{d},2020-0{1 + d % 9}-11,"Last{d}, First; Other, A",J,https://x.example/{d}; https://y.example/{d}\n')
    with open(os.path.join(a, "embeddings.vec"), "w", newline="") as f:
        f.write("200 16\n")
        for r in range(1, 201):
            f.write(f"t{r} " + " ".join(f"{rng.gauss(0, 1):.4f}" for _ in range(16)) + "\n")
    return a, b


def _last_json(text):
    return json.loads([ln for ln in text.strip().splitlines() if ln.startswith("{")][-1])


@pytest.mark.parametrize("seed", [1, 2])
def test_damaged_segment_files_never_crash_the_loader(drivers, corpora, workdir, seed):
    scratch = os.path.join(workdir, f"san_scratch_{seed}")
    os.makedirs(scratch, exist_ok=True)
    r = subprocess.run([drivers["asan"], "fuzzload", corpora[0], scratch, "250", str(seed)], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1"))
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    assert "ERROR: AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr and "LeakSanitizer" not in r.stderr, r.stderr[-3000:]
    line = _last_json(r.stdout)
    assert line["loaded"] + line["refused"] == 250 and line["refused"] > 100, line     # most damage is noticed and refused


def test_reload_under_load_is_race_free_and_never_mixes_generations(drivers, corpora, workdir):
    link = os.path.join(workdir, "san_link")
    r = subprocess.run([drivers["tsan"], "stress", corpora[0], corpora[1], link, "6", "4"], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, TSAN_OPTIONS="halt_on_error=0:report_signal_unsafe=0"))
    assert "WARNING: ThreadSanitizer" not in r.stderr, r.stderr[-4000:]
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    line = _last_json(r.stdout)
    assert line["mixed"] == 0 and line["errors"] == 0 and line["reloads"] > 10 and line["batches"] > 50, line
