"""CPU test of the skewed Myers score pass (genomeworks_b200/csrc/myers_skew.cuh).

The kernel's lane-level functions live in a host+device header; tests/cpp/myers_skew_model.cpp compiles them with g++, runs them
lane by lane in the kernel's step order and record layout, and compares score_at() with the oracle's get_myers_score()
(reference formulation: cudaaligner/src/myers_gpu.cu:243-255, 629-846) for every cell of every band pass of a set of pairs:
similar and unrelated sequences, query longer / shorter than the target, clamped bands (asymmetric), the unbanded case and the
narrowest supported band. The C4 shape (10 000 x 10 000, bands 501 and 1001) runs with --big."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def model_binary():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], stdout=subprocess.DEVNULL)
    exe = os.path.join(ROOT, "build", "myers_skew_model")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "cpp", "myers_skew_model.cpp"), "-o", exe,
                           "-L", os.path.join(ROOT, "oracle"), "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    return exe


def test_skewed_pass_equals_reference_band_cell_by_cell(model_binary):
    out = subprocess.run([model_binary], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:]
    assert "PASS" in out.stdout.splitlines()[-1]


def test_skewed_pass_c4_shape(model_binary):
    out = subprocess.run([model_binary, "--big"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:]
    lines = [l for l in out.stdout.splitlines() if "q=10000 t=10000" in l]
    assert len(lines) >= 2 and all(l.startswith("ok") for l in lines), lines
