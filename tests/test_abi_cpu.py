"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/gwb200.h declares, the
host logic that needs no GPU (BatchConfig derivation, error decoding, synthetic generators) matches the oracle / reference
semantics, and every engine entry fails loudly without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from genomeworks_b200 import _lib
    L = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "gwb200.h")).read()
    names = sorted(set(re.findall(r"\b(gwb200_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 45
    missing = [n for n in names if not hasattr(L, n)]
    assert missing == []
    assert b"sm_100a" in L.gwb200_version()


@pytest.mark.parametrize("args", [
    (1024, 100, 256, 0, 2.0, 3.0, 0), (1024, 16, 256, 1, 2.0, 3.0, 0), (10240, 32, 256, 2, 2.0, 3.0, 0), (10240, 32, 256, 2, 6.0, 3.0, 0),
    (32768, 64, 256, 2, 2.0, 3.0, 0), (1024, 2, 128, 3, 2.0, 3.0, 0), (500, 7, 200, 4, 2.5, 2.0, 77),
])
def test_batch_config_matches_reference_derivation(args):
    from genomeworks_b200 import cudapoa
    cfg = cudapoa.make_config(args[0], args[1], args[2], args[3], adaptive_storage_factor=args[4], graph_length_factor=args[5], max_pred_dist=args[6])
    got = [cfg.max_sequence_size, cfg.max_consensus_size, cfg.max_nodes_per_graph, cfg.matrix_sequence_dimension, cfg.alignment_band_width,
           cfg.max_sequences_per_poa, cfg.band_mode, cfg.max_banded_pred_distance]
    assert got == list(ol.batch_config(*args))


def test_batch_config_known_values():
    from genomeworks_b200 import cudapoa
    c2 = cudapoa.make_config(1024, 16, 256, "static_band")  # C2
    assert (c2.max_nodes_per_graph, c2.matrix_sequence_dimension, c2.max_consensus_size) == (3072, 264, 2048)
    c3 = cudapoa.make_config(10240, 32, 256, "adaptive_band")  # C3
    assert (c3.max_nodes_per_graph, c3.matrix_sequence_dimension, c3.max_consensus_size) == (30720, 528, 20480)
    c3b = cudapoa.make_config(10240, 32, 256, "adaptive_band", adaptive_storage_factor=6.0)
    assert c3b.matrix_sequence_dimension == 1584


def test_config_errors_are_invalid_argument():
    from genomeworks_b200 import cudapoa
    with pytest.raises(ValueError):
        cudapoa.make_config(-1, 10)
    with pytest.raises(ValueError):
        cudapoa.make_config_explicit(1024, 100, 3072, 256, 10, 264, "static_band")  # max_consensus < max_sequence
    with pytest.raises(ValueError):
        cudapoa.make_config_explicit(100, 200, 300, 256, 10, 264, "static_band")  # band wider than max sequence


def test_decode_error_and_status_names():
    from genomeworks_b200 import _lib, cudapoa
    L = _lib.lib()
    m = C.create_string_buffer(256)
    h = C.create_string_buffer(256)
    for st in range(1, 13):
        assert L.gwb200_poa_decode_error(C.c_int32(st), m, C.c_int32(256), h, C.c_int32(256)) == 0
        assert len(m.value) > 5
    assert L.gwb200_poa_decode_error(C.c_int32(99), m, C.c_int32(256), h, C.c_int32(256)) == _lib.E_RUNTIME
    assert cudapoa.status_to_str(6) == "exceeded_adaptive_banded_matrix_size"


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from genomeworks_b200 import cudaaligner, cudapoa
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cudapoa.CudaPoaBatch(10, 1024, 1 << 30, config=cudapoa.make_config(1024, 10))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cudaaligner.FixedBandAligner(64)


def test_synthetic_generators_are_deterministic_and_shaped():
    from genomeworks_b200 import synth
    a = synth.poa_windows(3, 200, 5, 8, 4, 4, seed0=1000)
    b = synth.poa_windows(3, 200, 5, 8, 4, 4, seed0=1000)
    assert (a[1] == b[1]).all() and (a[2] == b[2]).all()
    w = synth.split_windows(*a)
    assert len(w) == 3 and all(len(x) == 5 for x in w)
    for win in w:
        assert len(win[0]) == 200 and set(win[0]) <= set("ACGT")  # read 0 is the backbone
        assert all(abs(len(r) - 200) <= 4 for r in win)
    ql, qd, tl, td = synth.aligner_pairs(4, 3000, seed=1)
    assert (ql == 3000).all() and (tl <= 3000).all() and (tl > 2800).all()


def test_cpp_api_headers_compile_and_link():
    """The reference-named C++ API (include/claraparabricks/genomeworks/...) compiles against the C ABI."""
    out = os.path.join(ROOT, "build", "test_cpp_api_cpu")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O0", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include",
           os.path.join(ROOT, "tests", "cpp", "test_cpp_api.cpp"), "-o", out, "-L", os.path.join(ROOT, "genomeworks_b200"), "-lgwb200",
           "-L", "/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + os.path.join(ROOT, "genomeworks_b200"), "-Wl,-rpath,/usr/local/cuda/lib64"]
    subprocess.check_call(cmd)
    assert os.path.exists(out)


@pytest.mark.parametrize("src", ["tests/cpp/test_cpp_api.cpp", "tests/cpp/test_utils_cpu.cpp", "tools/cudapoa_cli.cpp"])
def test_cpp_headers_are_cxx14_clean(src):
    """The reference builds its callers with -std=c++14 (cmake/CXX.cmake): the drop-in headers must compile there, warning-free."""
    cmd = ["g++", "-std=c++14", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include",
           os.path.join(ROOT, src)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_c_header_is_plain_c():
    """include/gwb200.h is the FFI boundary: it must compile as C (no C++ or torch types in the signatures)."""
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", "gwb200.h")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_pygenomeworks_import_alias():
    """`import genomeworks.cudapoa` style imports of pygenomeworks callers resolve to this engine through genomeworks_b200/compat."""
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r);"
            "from genomeworks.cudapoa import CudaPoaBatch, status_to_str; from genomeworks.cudaaligner import CudaAlignerBatch;"
            "import genomeworks.cuda as cuda; assert cuda.CudaStream and cuda.CudaRuntimeError and cuda.cuda_get_mem_info;"
            "assert cuda.cuda_get_device_count() >= 0; import genomeworks_b200.cudapoa as m; assert CudaPoaBatch is m.CudaPoaBatch;"
            "print('ok')") % (ROOT, os.path.join(ROOT, "genomeworks_b200", "compat"))
    r = subprocess.run([os.sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stderr
