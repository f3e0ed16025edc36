"""Quick C2 timing: ours vs the reference kernels on the same GPU (development aid; bench.py is the contract)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from genomeworks_b200 import cudapoa, synth
import ref_lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
win_nseq, seq_len, data = synth.poa_windows(n, 1000, 16, 20, 10, 10, seed0=1000, max_read_len=1024)
cfg = cudapoa.make_config(1024, 16, 256, "static_band")
b = cudapoa.CudaPoaBatch(16, 1024, 16 << 30, config=cfg)
for it in range(3):
    b.reset()
    t0 = time.time()
    b.add_poa_groups_flat(win_nseq, seq_len, data)
    t1 = time.time()
    b.generate_poa()
    c, cov, lens, st = b.get_consensus_arrays()
    t2 = time.time()
    print("ours: add %.1f ms, generate+get %.1f ms, kernel %.2f ms, cells %.3e, ok=%d" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, b.last_kernel_ms(),
                                                                                      b.last_cells(), int((st == 0).sum())), flush=True)
if ref_lib.have_gwref():
    for it in range(2):
        r = ref_lib.ref_poa_run(win_nseq, seq_len, data, 1024, 16, 256, 1)
        print("ref : total %.1f ms, generate+get %.1f ms, batches %d" % (r["timings"][0], r["timings"][1], r["timings"][2]), flush=True)
