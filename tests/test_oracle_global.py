"""CPU: the oracle of the fixed-size global aligners (oracle/global_oracle.cpp) against the reference's own known answers:
Test_AlignerGlobal.cpp:78-147 (Default / HirschbergMyers / Myers rows) and test_cudaaligner_bindings.py:27-45."""
import random

import numpy as np
import pytest

import oracle_lib as ol

KAT = [("AAAA", "TTAT", "4M", 3), ("ATAAAAAAAA", "AAAAAAAAA", "1M1D8M", 1), ("AAAAAAAAA", "ATAAAAAAAA", "1M1I8M", 1),
       ("ACTGA", "GCTAG", "3M1D1M1I", 3), ("ACTG", "ACTG", "4M", 0), ("A", "T", "1M", 1),
       ("", "GACTCTCCCCCTCCCCTTTAAATATATAAAAATGGGGTGTAGCTAG", "46I", 46), ("GACTCTCCCCCTCCCCTTTAAATATATAAAAATGGGGTGTAGCTAG", "", "46D", 46),
       ("", "", "", 0)]
PY_KAT = [("AAAAAAA", "TTTTTTT", "7M"), ("AAATC", "TACGTTTT", "3M1I2M2I"), ("TACGTA", "ACATAC", "1D5M1I"), ("TGCA", "ATACGCT", "1I1M2I3M")]


def edit_distance(states):
    return int(np.count_nonzero(np.asarray(states) != 0))


def plain_edit_distance(q, t):
    prev = list(range(len(t) + 1))
    for i in range(1, len(q) + 1):
        cur = [i] + [0] * len(t)
        for j in range(1, len(t) + 1):
            cur[j] = min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (q[i - 1] != t[j - 1]))
        prev = cur
    return prev[len(t)]


def consumes(states, q, t):
    s = np.asarray(states)
    return int(np.count_nonzero(s != 2)) == len(q) and int(np.count_nonzero(s != 3)) == len(t)


@pytest.mark.parametrize("q,t,cigar,dist", KAT)
def test_reference_table_hirschberg_and_myers(q, t, cigar, dist):
    st, failed = ol.hirschberg_myers_align(q, t, max_query_length=max(len(q), 1))
    assert not failed
    assert ol.states_to_cigar(st) == cigar and edit_distance(st) == dist
    st2 = ol.myers_full_align(q, t)
    assert ol.states_to_cigar(st2) == cigar and edit_distance(st2) == dist


@pytest.mark.parametrize("q,t,cigar", PY_KAT)
def test_python_binding_table(q, t, cigar):
    st, failed = ol.hirschberg_myers_align(q, t, max_query_length=len(q))
    assert not failed and ol.states_to_cigar(st) == cigar


def test_random_pairs_are_optimal_alignments():
    rng = random.Random(3)
    for n in (1, 2, 31, 32, 33, 62, 63, 64, 65, 127, 200, 700):
        ref = "".join(rng.choice("ACGT") for _ in range(n))
        q = "".join(c for c in ref if rng.random() > 0.05)
        t = "".join(c if rng.random() > 0.08 else rng.choice("ACGT") for c in ref) + "".join(rng.choice("ACGT") for _ in range(rng.randrange(4)))
        d = plain_edit_distance(q, t)
        for mq in (len(q), max(len(q), 1) * 8):
            st, failed = ol.hirschberg_myers_align(q, t, max_query_length=max(mq, 1))
            assert not failed and consumes(st, q, t) and edit_distance(st) == d, (n, mq)
        st2 = ol.myers_full_align(q, t)
        assert consumes(st2, q, t) and edit_distance(st2) == d


@pytest.mark.parametrize("q,t,cigar,dist", [k for k in KAT if k[0] and k[1]])
def test_reference_table_ukkonen(q, t, cigar, dist):
    # Test_AlignerGlobal.cpp:138: the Ukkonen rows share the Default table (empty sequences excluded, :151-152)
    st = ol.ukkonen_align(q, t)
    assert ol.states_to_cigar(st) == cigar and edit_distance(st) == dist


def test_ukkonen_random_pairs_are_optimal_within_the_band():
    rng = random.Random(4)
    for n in (5, 40, 100, 300, 900):
        ref = "".join(rng.choice("ACGT") for _ in range(n))
        q = "".join(c for c in ref if rng.random() > 0.03)
        t = "".join(c if rng.random() > 0.05 else rng.choice("ACGT") for c in ref)
        st = ol.ukkonen_align(q, t)
        assert consumes(st, q, t) and edit_distance(st) == plain_edit_distance(q, t)
        st = ol.ukkonen_align(t, q)  # query longer than target: the swapped orientation
        assert consumes(st, t, q) and edit_distance(st) == plain_edit_distance(q, t)
