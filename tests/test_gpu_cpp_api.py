"""Runs the C++ API test program (tests/cpp/test_cpp_api.cpp) that uses the reference-named classes on the GPU."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_api_program():
    exe = os.path.join(ROOT, "build", "test_cpp_api")
    if not os.path.exists(exe):
        cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include",
               os.path.join(ROOT, "tests", "cpp", "test_cpp_api.cpp"), "-o", exe, "-L", os.path.join(ROOT, "genomeworks_b200"), "-lgwb200",
               "-L", "/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + os.path.join(ROOT, "genomeworks_b200"), "-Wl,-rpath,/usr/local/cuda/lib64"]
        subprocess.check_call(cmd)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert "CPP_API_OK" in out.stdout, out.stdout + out.stderr
