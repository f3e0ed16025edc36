"""Pins the CPU oracle (oracle/poa_oracle.cpp) against the reference's own known-answer tests.

Vectors transcribed from /root/reference/cudapoa/tests:
  Test_CudapoaTopSort.cu:48-58, Test_CudapoaAddAlignment.cu:111-231, Test_CudapoaNW.cu:83-187 and :444-508,
  Test_CudapoaGenerateConsensus.cu:83-165, Test_CudapoaBatchEnd2End.cu:39-91 (+ cudapoa/data sample windows / golden),
  Test_CudapoaBatch.cu:155-203, pygenomeworks/test/test_cudapoa_bindings.py (complex batch).
"""
import gzip
import os
import random

import numpy as np
import pytest

import oracle_lib as ol

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


# ---- Test_CudapoaTopSort.cu:48-58 ------------------------------------------------------------------
@pytest.mark.parametrize("edges,answer", [
    ([[], [], [3], [1], [0, 1], [0, 2]], "4-5-0-2-3-1"),
    ([[1, 3], [2, 3], [3, 4, 5], [4, 5], [5], []], "0-1-2-3-4-5"),
    ([[], [], [3], [1], [0, 1, 7], [0, 2], [4], [5]], "6-4-7-5-0-2-3-1"),
])
def test_topsort_kat(edges, answer):
    assert "-".join(str(x) for x in ol.topsort(edges)) == answer


# ---- Test_CudapoaAddAlignment.cu:111-231 -----------------------------------------------------------
ADD_CASES = [
    # answer out-edges, nodes, out_edges, coverage, read, base weights, alignment_graph, alignment_read
    ([[], [0], [1], [2, 4], [1]], "AAAA", [[], [0], [1], [2]], [1, 1, 1, 1], "AATA", [0, 0, 1, 2], [0, 1, 2, 3], [0, 1, 2, 3]),
    ([[], [0], [1], [2], [3]], "ATCG", [[], [0], [1], [2]], [1, 1, 1, 1], "ATCGA", [0, 1, 2, 3, 4], [0, 1, 2, 3, -1], [0, 1, 2, 3, 4]),
    ([[], [0], [1, 4, 5], [2], [0], [0]], "AACGC", [[], [0], [1, 4], [2], [0]], [2, 1, 2, 2, 1], "ATCG", [0, 1, 1, 5], [0, 4, 2, 3],
     [0, 1, 2, 3]),
    ([[], [0], [1], [2], [3, 0]], "ATTGA", [[], [0], [1], [2], [3]], [1, 1, 1, 1, 1], "AA", [5, 1], [0, 1, 2, 3, 4], [0, -1, -1, -1, 1]),
    ([[], [0], [1], [2, 6, 7], [3], [0], [5], [5]], "ATGTACA", [[], [0], [1], [2, 6], [3], [0], [5]], [2, 1, 1, 2, 2, 1, 1], "ACTTA",
     [10, 9, 8, 7, 6], [0, 5, 6, 3, 4], [0, 1, 2, 3, 4]),
]


@pytest.mark.parametrize("case", ADD_CASES)
def test_add_alignment_kat(case):
    ans, nodes, out_edges, cov, read, bw, ag, ar = case
    g = ol.OGraph(nodes, out_edges, coverage=cov)
    assert g.add_alignment(ag, ar, read, bw) == 0
    assert g.out_edges() == ans


# ---- Test_CudapoaNW.cu:83-187 (full band) ----------------------------------------------------------
NW_CASES = [
    ("3,2,1,0", "3,2,1,0", "AAAA", [0, 1, 2, 3], [[1], [2], [3], []], "AATA"),
    ("-1,3,2,1,0", "4,3,2,1,0", "ATCG", [0, 1, 2, 3], [[1], [2], [3], []], "ATCGA"),
    ("3,2,1,0", "3,2,1,0", "AACGC", [0, 4, 1, 2, 3], [[1, 4], [2], [3], [], [2]], "ATCG"),
    ("4,3,2,1,0", "1,-1,-1,-1,0", "ATTGA", [0, 1, 2, 3, 4], [[1], [2], [3], [4], []], "AA"),
    ("4,3,6,5,0", "4,3,2,1,0", "ATGTACA", [0, 5, 1, 6, 2, 3, 4], [[1, 5], [2], [3], [4], [], [6], [3]], "ACTTA"),
]


@pytest.mark.parametrize("case", NW_CASES)
def test_nw_full_kat(case):
    ans_g, ans_r, nodes, sorted_graph, out_edges, read = case
    g = ol.OGraph(nodes, out_edges, sorted_graph=sorted_graph)
    r, ag, ar = g.nw(read, mode=0)
    assert r > 0
    assert ",".join(map(str, ag)) == ans_g
    assert ",".join(map(str, ar)) == ans_r


# ---- Test_CudapoaNW.cu:444-508: banded == full on a 500-node chain vs a 520-base read ---------------
NODES_STR = "TTTAACCTAATAAATCAGTGAAGATTTAAAATATGATAATTATTGATTTTGGTGAGAGTGCAAAGAAATTTGTTACCCTCATAAGCTGAGCAGACAGATAAGATAGAAAAACAGAAGATAGAATATTAAAACCATGATAGGTACAGACTGAAAAATTCTTGGATAAATATTAAAATTTAGGCTTTAGTAGTAGATTGATGACTGTGAGGAAAAAGGATGTCCAATTGTTGAGTGACATGTAGAATGCCTTAAAATAATTTTACACGTCACTGAAAGCTATATTTATATTCAGGAAGGATATATCCCAGTCATGATTTTCTTAATAAGTTGCCCCATTTTCCAAGTTTAGCTAATTAACATTTATGTCTTCTATAATCAGGAATAGTCATTAACTGACACAGAAACAATTGGAAGCATATGTAGCCAAAAACATAAAAATTATTGCATCCAAATAATGATAAAGTAAAATATTAAAAAATATAGTCTTCTAAAT"
READ_STR = "TTTCACCTAGAAAATCAGTGAAGATTTAACAAAAAAAAAAAAAAAAAAAAAAAAATATTGATAATTATTGATTTTGGTGAGAGTGCAAAGCAATTGGCTACCCTCATAAGCTGAGCAGAAGATAAGATAGACAACAGAAGATAGAATAGTTAAACCATGATAGGTACAGACTGCAAAAAAATTCGATAAATATTAAAATTTAGGGCTTTAGTATATATTGATGACTGAGAAAAATCGTGATGTGCAATTGTGCGTGACATGTAGAATTGCCTTAAATAAAATTTAATCTGTCACTGAAGCTATATTTATATTCAGGAAGGATATATCCCAGTCATTGCTTTTCTTAATAAGTGCCCATGTTCCAAGTTTAGCCTAATTAAAAACTTTATGTCTTCTATATCAGAATAGTCATTAATGCACAGAAACAATTTGCGAAGGCATTATGTAGCAAAAACATAAAAAATTATTGCAGCCAAATAATGAATAAAAGTAACACAATCATTTAAAAAAATTATTATGTACTTCTAAAC"


# modes 3 / 4 = static / adaptive band with traceback matrix: NWStaticBandTracebackvsFull, NWAdaptiveBandTracebackvsFull (:491-511)
@pytest.mark.parametrize("mode", [1, 2, 3, 4])
def test_nw_banded_equals_full(mode):
    n = len(NODES_STR)
    edges = [[i + 1] for i in range(n - 1)] + [[]]
    g = ol.OGraph(NODES_STR, edges, sorted_graph=list(range(n)))
    rf, agf, arf = g.nw(READ_STR, mode=0)
    # BatchConfig(1024, 2, 128, static/adaptive) -> matrix_sequence_dimension 136 / 272 (Test_CudapoaNW.cu:326-327)
    msd = 136 if mode in (1, 3) else 272
    rb, agb, arb = g.nw(READ_STR, mode=mode, band_width=128, max_nodes=3072, matrix_seq_dim=msd)
    assert rf > 0 and rb == rf
    assert agb == agf and arb == arf


# ---- Test_CudapoaGenerateConsensus.cu:83-165 --------------------------------------------------------
def _in_w_raw(out_edges, out_w, n, max_e=50):
    # the reference harness stores the weight at slot index == FROM-node id (get_incoming_edge_w); other slots stay 0
    raw = np.zeros(n * max_e, dtype=np.uint16)
    for i, (es, ws) in enumerate(zip(out_edges, out_w)):
        for to, w in zip(es, ws):
            raw[to * max_e + i] = w
    return raw


CONS_CASES = [
    ("ATAA", "AAAAT", [0, 1, 2, 4, 3], [[], [], [4], [], [2]], [[1], [2, 4], [3], [], [3]], [2, 2, 1, 2, 1], [[5], [4, 3], [2], [], [1]]),
    ("AGCTA", "ATCGA", [0, 1, 2, 3, 4], [[], [], [], [], []], [[1], [2], [3], [4], []], [1, 1, 1, 1, 1], [[4], [3], [2], [1], []]),
    ("GCCA", "AACGCT", [0, 1, 4, 5, 2, 3], [[], [4, 5], [], [], [1, 5], [1, 4]], [[1, 4, 5], [2], [3], [], [2], [2]], [3, 1, 3, 3, 1, 1],
     [[7, 6, 5], [4], [3], [], [2], [1]]),
    ("AGTTA", "ATTGA", [0, 1, 2, 3, 4], [[], [], [], [], []], [[1, 4], [2], [3], [4], []], [2, 1, 1, 1, 2], [[5, 4], [3], [2], [1], []]),
    ("ATTCA", "ATGTACAT", [0, 1, 5, 2, 6, 7, 3, 4], [[], [5], [6, 7], [], [], [1], [2, 7], [2, 6]],
     [[1, 5], [2], [3], [4], [], [6, 7], [3], [3]], [3, 1, 1, 3, 3, 2, 1, 1], [[9, 8], [7], [6], [5], [], [4, 3], [2], [1]]),
]


@pytest.mark.parametrize("case", CONS_CASES)
def test_consensus_kat(case):
    ans, nodes, sg, aln, out_edges, cov, out_w = case
    g = ol.OGraph(nodes, out_edges, sorted_graph=sg, in_w_raw=_in_w_raw(out_edges, out_w, len(nodes)), coverage=cov, alignments=aln)
    st, cons_rev, _ = g.consensus()
    assert st == 0
    assert cons_rev == ans  # the reference test compares the device (reversed) string


# ---- End-to-end golden: Test_CudapoaBatchEnd2End.cu:39-91 ------------------------------------------
def load_sample_windows():
    with gzip.open(os.path.join(GOLDEN, "sample-windows.txt.gz"), "rt") as f:
        lines = f.read().split("\n")
    windows, i = [], 0
    while i < len(lines) and lines[i].strip():
        n = int(lines[i])
        windows.append(lines[i + 1:i + 1 + n])
        i += 1 + n
    return windows


def assembly(consensus, coverage):
    # MultiBatch::assembly(), cudapoa/benchmarks/multi_batch.hpp:182-213
    genome = ""
    for c, cov in zip(consensus, coverage):
        cov = [int(x) for x in cov]
        avg = sum(cov) // len(cov)
        begin, end = 0, len(c) - 1
        while begin < len(cov) and cov[begin] < avg:
            begin += 1
        while end >= 0 and cov[end] < avg:
            end -= 1
        if begin < end:
            genome += c[begin:end + 1]
    return genome


def test_end2end_golden_full_band():
    windows = load_sample_windows()
    assert len(windows) == 67
    cfg = ol.batch_config(1024, 200)  # BatchConfig(1024, 200) => full_band (multi_batch.hpp:49)
    res = ol.poa_run(windows, cfg)
    assert (res["status"] == 0).all()
    golden = open(os.path.join(GOLDEN, "sample-golden-value.txt")).read().strip()
    assert len(golden) == 33457
    assert assembly(res["consensus"], res["coverage"]) == golden


# ---- Test_CudapoaBatch.cu:155-203: 3 x 'A'*1023 -> consensus == input --------------------------------
@pytest.mark.parametrize("band_mode", [0, 1, 2, 3, 4])
def test_identity_consensus(band_mode):
    seq = "A" * 1023
    cfg = ol.batch_config(1024, 10, 256, band_mode)
    res = ol.poa_run([[seq, seq, seq]], cfg)
    assert res["status"][0] == 0
    assert res["consensus"][0] == seq
    assert (res["coverage"][0] == 3).all()


# ---- pygenomeworks/test/test_cudapoa_bindings.py::test_cudapoa_complex_batch ------------------------
@pytest.mark.parametrize("band_mode", [0, 1, 2, 3, 4])
def test_complex_batch_consensus_equals_reference(band_mode):
    random.seed(2)
    read_len = 500
    ref = "".join(random.choice("ACTG") for _ in range(read_len))
    num_reads = 100
    mutation_rate = 0.02
    reads = []
    for _ in range(num_reads):
        new_read = "".join(r if random.random() > mutation_rate else random.choice("ACTG") for r in ref)
        reads.append(new_read)
    cfg = ol.batch_config(1024, 100, 256, band_mode)
    res = ol.poa_run([reads], cfg)
    assert res["status"][0] == 0
    assert res["consensus"][0] == ref


def test_msa_rows_strip_to_inputs():
    # Test_CudapoaGenerateMSA2.cu:85-129 weak invariant: removing '-' from each MSA row gives back the read
    rng = random.Random(7)
    backbone = "".join(rng.choice("ACGT") for _ in range(50))
    reads = [backbone]
    for _ in range(40):
        r = list(backbone)
        for _ in range(3):
            p = rng.randrange(len(r))
            op = rng.randrange(3)
            if op == 0:
                r[p] = rng.choice("ACGT")
            elif op == 1:
                r.insert(p, rng.choice("ACGT"))
            else:
                del r[p]
        reads.append("".join(r))
    cfg = ol.batch_config(1024, 100, 256, 1)
    res = ol.poa_run([reads], cfg, msa=True)
    assert res["status"][0] == 0
    rows = res["msa"][0]
    assert len(rows) == len(reads)
    assert len(set(len(r) for r in rows)) == 1
    for row, rd in zip(rows, reads):
        assert row.replace("-", "") == rd


# ---- outputs of the unmodified reference kernels (tests/golden/make_reference_fixture.py, run on a B200) ---------------
@pytest.mark.parametrize("case_idx", range(5))
def test_oracle_matches_reference_kernel_fixture(case_idx):
    """Pins the CPU restatement to the reference ITSELF: consensus, coverage and status that libgwref.so (the reference's
    own kernels rebuilt for sm_100a) produced for 12 synthetic windows in every band mode, incl. both traceback modes."""
    import gzip
    import json
    from genomeworks_b200 import synth
    fx = json.load(gzip.open(os.path.join(GOLDEN, "reference_kernel_outputs.json.gz"), "rt"))
    gen = fx["generator"]
    win_nseq, seq_len, data = synth.poa_windows(gen["n_windows"], gen["backbone"], gen["reads"], gen["mut"], gen["ins"], gen["dele"],
                                                seed0=gen["seed0"])
    case = fx["cases"][case_idx]
    assert case["oracle_with_ieee_division_identical"]  # recorded at generation time: div.approx vs IEEE does not matter for these windows
    cfg = ol.batch_config(fx["max_sequence_size"], fx["max_sequences_per_poa"], case["band_width"], case["band_mode"],
                          max_pred_dist=case["max_pred"])
    res = ol.poa_run(synth.split_windows(win_nseq, seq_len, data), cfg)
    assert [int(x) for x in res["status"]] == case["status"]
    assert list(res["consensus"]) == case["consensus"]
    assert [[int(v) for v in cv] for cv in res["coverage"]] == case["coverage"]
