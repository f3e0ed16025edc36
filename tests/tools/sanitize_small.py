"""Development aid (GPU box): a small pass over the hot paths for compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool memcheck python tests/tools/sanitize_small.py"""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from genomeworks_b200 import cudaaligner, cudapoa, synth

rng = random.Random(3)


def seq(n):
    return "".join(rng.choice("ACGT") for _ in range(n))


def mutate(q, n):
    t = list(q)
    for _ in range(n):
        p = rng.randrange(len(t))
        op = rng.randrange(3)
        if op == 0:
            t[p] = rng.choice("ACGT")
        elif op == 1:
            t.insert(p, rng.choice("ACGT"))
        elif len(t) > 1:
            del t[p]
    return "".join(t)


# banded Myers: skewed passes (bands >= 128 rows), classic passes (narrow bands), both backtrace stages
pairs = []
for L in (100, 700, 1500, 2500):
    a = seq(L)
    pairs.append((a, mutate(a, L // 12)))
    pairs.append((mutate(a, L // 25), a))
pairs.append((seq(900), seq(850)))
for bw in (64, 512, 1024):
    al = cudaaligner.FixedBandAligner(bw)
    for q, t in pairs:
        al.add_alignment(q, t)
    al.align_all()
    al.sync_alignments()
    res = al.get_alignments()
    print("aligner bw", bw, [r.status for r in res][:6], flush=True)
    al.close()

# POA: static band int16 (C2 shape, few windows) and adaptive band int32 (4 kb windows)
win_nseq, seq_len, data = synth.poa_windows(24, 1000, 16, 20, 10, 10, seed0=1000, max_read_len=1024)
cfg = cudapoa.make_config(1024, 16, 256, "static_band")
b = cudapoa.CudaPoaBatch(16, 1024, 1 << 30, output_type="consensus", config=cfg)
b.add_poa_groups_flat(win_nseq, seq_len, data)
b.generate_poa()
c, cov, lens, st = b.get_consensus_arrays()
print("poa c2-shape statuses", np.bincount(st), flush=True)
L = 4000
win_nseq, seq_len, data = synth.poa_windows(6, L, 12, L // 50, L // 100, L // 100, seed0=7, max_read_len=4096)
cfg = cudapoa.make_config(4096, 12, 256, "adaptive_band", adaptive_storage_factor=3.0)
b = cudapoa.CudaPoaBatch(12, 4096, 3 << 30, output_type="consensus", config=cfg)
b.add_poa_groups_flat(win_nseq, seq_len, data)
b.generate_poa()
c, cov, lens, st = b.get_consensus_arrays()
print("poa adaptive 4 kb statuses", np.bincount(st), flush=True)
print("SANITIZE_RUN_DONE")
