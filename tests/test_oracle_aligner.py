"""Pins the CPU oracle of the banded Myers aligner (oracle/myers_oracle.cpp) against the reference's known-answer tests:
cudaaligner/tests/Test_AlignerGlobal.cpp:73-157 (MyersBanded rows, max_bandwidth 1024), Test_ApproximateBandedMyers.cpp:48-172,
pygenomeworks/test/test_cudaaligner_bindings.py (CIGAR list)."""
import random

import pytest

import oracle_lib as ol

KAT = [
    ("AAAA", "TTAT", "4M", 3),
    ("ATAAAAAAAA", "AAAAAAAAA", "1M1D8M", 1),
    ("AAAAAAAAA", "ATAAAAAAAA", "1M1I8M", 1),
    ("ACTGA", "GCTAG", "3M1D1M1I", 3),
    ("ACTG", "ACTG", "4M", 0),
    ("A", "T", "1M", 1),
    ("", "GACTCTCCCCCTCCCCTTTAAATATATAAAAATGGGGTGTAGCTAG", "46I", 46),
    ("GACTCTCCCCCTCCCCTTTAAATATATAAAAATGGGGTGTAGCTAG", "", "46D", 46),
    ("", "", "", 0),
    # pygenomeworks/test/test_cudaaligner_bindings.py
    ("AAAAAAA", "TTTTTTT", "7M", 7),
    ("AAATC", "TACGTTTT", "3M1I2M2I", None),
    ("TACGTA", "ACATAC", "1D5M1I", None),
    ("TGCA", "ATACGCT", "1I1M2I3M", None),
]


@pytest.mark.parametrize("q,t,cigar,ed", KAT)
def test_aligner_global_kat(q, t, cigar, ed):
    r = ol.myers_align(q, t, 1024)
    assert r["status"] == 0
    assert r["is_optimal"] == 1
    assert r["cigar"] == cigar
    if ed is not None:
        assert r["edit_distance"] == ed


@pytest.mark.parametrize("q,t,cigar", [
    ("AACCGGTTAACCGGTTAACCGGTTTT", "AACCGGTTAAAACCCCGGGGGTTAAACGGTT", "10M2I2M2I7M3I5M2D"),
    ("AACCGGTTAACCGGTTAACCGGTTT", "AACCGGTTAAAACCCCGGGGGTTAACCGGTT", "10M2I2M2I3M2I3M1I6M1D"),
])
def test_implicit_new_entries_bw7(q, t, cigar):
    # Test_ApproximateBandedMyers.cpp:95-119
    r = ol.myers_align(q, t, 7)
    assert r["status"] == 0
    assert r["is_optimal"] == 0
    assert r["cigar"] == cigar


def test_edit_distance_monotone_with_bandwidth():
    # Test_ApproximateBandedMyers.cpp:122-172
    q = "AGGGCGAATATCGCCTCCCGCATTAAGCTGTACCTTCCAGCCCCGCCGGTAATTCCAGCCGGTTGAAGCCACGTCTGCCACGGCACAATGTTTTCGCTTTGCCCGGTGACGGATTTAATCCACCACAG"
    t = "AGGGCGAATATCGCCTCCGCATTAAACTGTACTTCCCAGCCCCGCCAGTATTCCAGCGGGTTGAAGCCGCGTCTGCCACAGCGCAATGTTTTCTTTGCCCACGGTGACCGGTTTAGTCACTACAGTTGC"
    true_ed = 23
    last = 1 << 30
    bws = [2, 4, 16, 31, 32, 34, 63, 64, 66, 255, 256, 258, 1023, 1024, 1026, 2048]
    for bw in bws:
        if bw % 32 == 1:
            continue
        r = ol.myers_align(q, t, bw)
        if r["status"] == 0:
            assert r["edit_distance"] <= last
            if r["edit_distance"] > true_ed:
                assert r["is_optimal"] == 0
            if bw == bws[-1]:
                assert r["is_optimal"] == 1 and r["edit_distance"] == true_ed
            last = r["edit_distance"]


def _naive_edit_distance(a, b):
    prev = list(range(len(b) + 1))
    for i in range(1, len(a) + 1):
        cur = [i] + [0] * len(b)
        for j in range(1, len(b) + 1):
            cur[j] = min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (a[i - 1] != b[j - 1]))
        prev = cur
    return prev[-1]


def _apply(q, t, actions, runs):
    """Checks that the RLE path is a valid global alignment of q against t and returns its cost."""
    i = j = cost = 0
    for a, r in zip(actions, runs):
        a, r = int(a), int(r)
        if a in (0, 1):
            for _ in range(r):
                assert (q[i] == t[j]) == (a == 0)
                i += 1
                j += 1
            cost += r if a == 1 else 0
        elif a == 2:
            j += r
            cost += r
        else:
            i += r
            cost += r
    assert i == len(q) and j == len(t)
    return cost


@pytest.mark.parametrize("seed", range(6))
def test_random_pairs_optimal_when_band_is_wide(seed):
    rng = random.Random(seed)
    L = rng.choice([40, 97, 150, 300, 700, 1300])
    q = "".join(rng.choice("ACGT") for _ in range(L))
    t = list(q)
    for _ in range(L // 12):
        p = rng.randrange(len(t))
        op = rng.randrange(3)
        if op == 0:
            t[p] = rng.choice("ACGT")
        elif op == 1:
            t.insert(p, rng.choice("ACGT"))
        else:
            del t[p]
    t = "".join(t)
    r = ol.myers_align(q, t, 2048)
    assert r["status"] == 0 and r["is_optimal"] == 1
    assert _apply(q, t, r["actions"], r["runs"]) == r["edit_distance"]
    assert r["edit_distance"] == _naive_edit_distance(q, t)
    # narrow band: still a valid alignment, cost >= optimum
    r2 = ol.myers_align(q, t, 34)
    if r2["status"] == 0:
        assert _apply(q, t, r2["actions"], r2["runs"]) == r2["edit_distance"] >= r["edit_distance"]
