"""CPU checks of the host side around the POA path: the cudapoa CLI's option handling (reference: application_parameters.cpp)
and the window/FASTA parsers of utils.hpp. No compute happens here; on a box without a GPU the tool must fail loudly."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "build", "cudapoa")
UTEST = os.path.join(ROOT, "build", "test_utils_cpu")


@pytest.fixture(scope="module", autouse=True)
def built():
    import __graft_entry__ as g
    g.build()
    assert os.path.exists(CLI) and os.path.exists(UTEST)


def run(args):
    return subprocess.run([CLI] + args, capture_output=True, text=True, timeout=120)


def test_utils_parsers(tmp_path):
    r = subprocess.run([UTEST, str(tmp_path)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stderr


def test_cli_help_and_option_errors(tmp_path):
    r = run(["-h"])
    assert r.returncode == 0 and "Usage: cudapoa" in r.stderr and "--band-mode" in r.stderr
    wf = tmp_path / "w.txt"
    wf.write_text("2\nACGT\nACGA\n")
    assert run(["-i", str(wf), "-b", "7"]).returncode != 0
    assert "band-mode must be" in run(["-i", str(wf), "-b", "7"]).stderr
    assert "max-groups cannot be 0" in run(["-i", str(wf), "-M", "0"]).stderr
    assert "gap score must be non-positive" in run(["-i", str(wf), "-g", "3"]).stderr
    assert "pred-distance must be" in run(["-i", str(wf), "-D", "0"]).stderr
    assert "gpu-mem-alloc" in run(["-i", str(wf), "-R", "1.5"]).stderr
    assert "Invalid input file" in run(["-i", str(tmp_path / "missing.txt")]).stderr
    # two non-FASTA inputs are rejected with the usage text
    r = run(["-i", str(wf), "-i", str(wf)])
    assert r.returncode == 1 and "Invalid input." in r.stderr


def test_cli_without_gpu_fails_loudly(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    wf = tmp_path / "w.txt"
    wf.write_text("2\nACGT\nACGA\n")
    r = run(["-i", str(wf)])
    assert r.returncode != 0 and r.stdout == ""
    assert "CUDA device" in r.stderr
