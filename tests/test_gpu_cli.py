"""GPU test of the cudapoa command line tool (reference: cudapoa/src/main.cpp): the first windows of the reference's sample file,
8 reads each (BASELINE config C1 shape), through the tool; its consensus lines must be the oracle's."""
import gzip
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "build", "cudapoa")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def sample_windows(n_windows, n_reads):
    lines = gzip.open(os.path.join(GOLDEN, "sample-windows.txt.gz"), "rt").read().split("\n")
    windows, i = [], 0
    while i < len(lines) and len(windows) < n_windows:
        if not lines[i].strip():
            break
        n = int(lines[i])
        windows.append(lines[i + 1:i + 1 + n][:n_reads])
        i += 1 + n
    return windows


@pytest.mark.parametrize("band_mode", [1, 2, 3])
def test_cli_consensus_equals_oracle(tmp_path, band_mode, device_fdiv):
    windows = sample_windows(24, 8)
    wf = tmp_path / "windows.txt"
    wf.write_text("".join("%d\n%s\n" % (len(w), "\n".join(w)) for w in windows))
    r = subprocess.run([CLI, "-i", str(wf), "-b", str(band_mode), "-w", "256"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    got = [l for l in r.stdout.split("\n") if l]
    assert len(got) == len(windows)
    assert "Processed groups" in r.stderr
    cfg = ol.batch_config(1024, 8, 256, band_mode)
    res = ol.poa_run(windows, cfg)
    assert (res["status"] == 0).all()
    # the tool prints group by group in the order of its size bins: compare as multisets and, when a single bin was used, in order
    assert sorted(got) == sorted(res["consensus"])
    if r.stderr.count("(batch 0)") == r.stderr.count("Processed groups"):
        assert got == list(res["consensus"])


def test_cli_msa_and_graph_outputs(tmp_path):
    windows = sample_windows(4, 6)
    wf = tmp_path / "windows.txt"
    wf.write_text("".join("%d\n%s\n" % (len(w), "\n".join(w)) for w in windows))
    dot = tmp_path / "graphs.dot"
    r = subprocess.run([CLI, "-i", str(wf), "-a", "-d", str(dot)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    rows = [l for l in r.stdout.split("\n") if l]
    assert len(rows) == sum(len(w) for w in windows)
    # removing the gaps from an MSA row gives back a read of the input (Test_CudapoaGenerateMSA2.cu:85-129)
    reads = {s for w in windows for s in w}
    assert all(row.replace("-", "") in reads for row in rows)
    assert "digraph" in dot.read_text()
    # FASTA input: one group per file
    fa = tmp_path / "g0.fa"
    fa.write_text("".join(">r%d\n%s\n" % (k, s) for k, s in enumerate(windows[0])))
    r2 = subprocess.run([CLI, "-i", str(fa), "-b", "0"], capture_output=True, text=True, timeout=600)
    assert r2.returncode == 0, r2.stderr
    assert len([l for l in r2.stdout.split("\n") if l]) == 1
