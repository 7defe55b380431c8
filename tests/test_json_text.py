"""The product's JSON text emitter (csrc/host/json_text.hpp) against nlohmann::json::dump(), the library the
reference serialises its answer with (an un-vendored dependency of the reference; the image carries 3.11.3
inside cudnn_frontend).  Compiles tests/tools/json_text_check.cpp and runs it on a few million f32 scores
widened to double plus strings with every escape class and valid / invalid UTF-8."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JSON_INC = "/opt/prime-rl/.venv/lib/python3.12/site-packages/include/cudnn_frontend/thirdparty"


@pytest.mark.skipif(not os.path.exists(os.path.join(JSON_INC, "nlohmann", "json.hpp")), reason="nlohmann/json.hpp not in this image")
def test_number_and_string_text_equals_nlohmann_dump(workdir):
    exe = os.path.join(workdir, "json_text_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", JSON_INC, "-I", os.path.join(ROOT, "nextsearch-api_b200", "csrc", "host"),
                    os.path.join(ROOT, "tests", "tools", "json_text_check.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe, "3000000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "bad=0" in r.stdout
