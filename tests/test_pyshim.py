"""The reference's Cython shim (pygenomeworks/genomeworks/*.pyx, unmodified, cythonized from /root/reference) builds against
this repo's include/ + libgwb200.so (CPU test) and passes the cases of the reference's own binding tests on a B200 (GPU test,
own process). Build recipe: oracle/build_pyshim.py; INTEGRATION.md section 4."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import build_pyshim  # noqa: E402


def _ensure_lib():
    import __graft_entry__ as g
    g.build()


@pytest.mark.skipif(not build_pyshim.have_reference(), reason="reference sources absent (GPU box uses the prebuilt shim)")
def test_reference_cython_shim_builds_on_this_engine():
    _ensure_lib()
    out = build_pyshim.build(force=True)
    assert build_pyshim.built()
    # imports without a device, and is bound to this engine's library (not to libcudapoa / libcudaaligner / libgwbase)
    code = ("import sys; sys.path.insert(0, %r); import genomeworks.cuda.cuda, genomeworks.cudapoa.cudapoa as p, "
            "genomeworks.cudaaligner.cudaaligner as a; print(p.CudaPoaBatch.__name__, a.CudaAlignerBatch.__name__)" % out)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and "CudaPoaBatch CudaAlignerBatch" in r.stdout, r.stderr[-2000:]
    for m in build_pyshim.MODULES[1:]:
        ldd = subprocess.run(["ldd", os.path.join(out, "genomeworks", m + build_pyshim.suffix())], capture_output=True, text=True).stdout
        assert "libgwb200.so" in ldd and "libcudapoa" not in ldd and "libcudaaligner" not in ldd and "libgwbase" not in ldd


@pytest.mark.gpu
def test_reference_binding_test_cases_through_the_shim():
    if not build_pyshim.built():
        if not build_pyshim.have_reference():
            pytest.skip("shim not prebuilt and reference sources absent")
        _ensure_lib()
        build_pyshim.build()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "pyshim_cases.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "PYSHIM_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
