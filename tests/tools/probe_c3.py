"""Development probe: config C3 (10 kb x 32 reads, adaptive band) statuses and timing, ours vs reference."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from genomeworks_b200 import cudapoa, synth
import ref_lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
for factor in (2.0, 6.0):
    win_nseq, seq_len, data = synth.poa_windows(n, 10000, 32, 200, 100, 100, seed0=1000, max_read_len=10240)
    cfg = cudapoa.make_config(10240, 32, 256, "adaptive_band", adaptive_storage_factor=factor)
    b = cudapoa.CudaPoaBatch(32, 10240, 60 << 30, config=cfg)
    print("factor", factor, "max_poas", b.max_poas, "score_bytes", b.score_bytes, flush=True)
    for it in range(2):
        b.reset()
        b.add_poa_groups_flat(win_nseq, seq_len, data)
        t1 = time.time()
        b.generate_poa()
        c, cov, lens, st = b.get_consensus_arrays()
        t2 = time.time()
        print("ours: generate+get %.1f ms kernel %.2f ms cells %.3e status %s" % ((t2 - t1) * 1e3, b.last_kernel_ms(), b.last_cells(),
                                                                             np.bincount(st, minlength=13).tolist()), flush=True)
    ours_c = [bytes(c[i, :lens[i]]).decode() for i in range(n)]
    b.close()
    if ref_lib.have_gwref():
        r = ref_lib.ref_poa_run(win_nseq, seq_len, data, 10240, 32, 256, 2, adaptive_storage_factor=factor, mem_fraction=0.3,
                                max_windows_per_batch=n)
        print("ref : total %.1f ms generate+get %.1f ms status %s same_consensus=%s same_status=%s" % (
            r["timings"][0], r["timings"][1], np.bincount(r["status"], minlength=13).tolist(), r["consensus"] == ours_c,
            list(r["status"]) == list(st)), flush=True)
