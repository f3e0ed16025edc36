"""BASELINE config C5: POA throughput sweep, window length x coverage, adaptive band.
Errors per read (L/50 substitutions, L/100 insertions, L/100 deletions candidates), adaptive_storage_factor 6.0.
Prints one JSON line per cell: windows/s (kernel time, CUDA events), executed DP cells/s, fraction of the measured HBM
roofline, statuses. Cells whose windows do not all succeed are reported with their status histogram (SURVEY: "skip cells
where the oracle reports non-success").
usage: sweep_poa.py [--lengths 1,2,4,8,16,32] [--coverages 8,16,32,64] [--max-seconds 240] [--ref]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from genomeworks_b200 import cudapoa, synth

ap = argparse.ArgumentParser()
ap.add_argument("--lengths", default="1,2,4,8,16,32")
ap.add_argument("--coverages", default="8,16,32,64")
ap.add_argument("--max-seconds", type=float, default=240.0)
ap.add_argument("--ref", action="store_true", help="also run the unmodified reference kernels and compare outputs")
args = ap.parse_args()
peak = 6592.9
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
t_start = time.time()
import torch
free_b, _ = torch.cuda.mem_get_info()
for lk in [int(x) for x in args.lengths.split(",")]:
    for cov in [int(x) for x in args.coverages.split(",")]:
        if time.time() - t_start > args.max_seconds:
            print(json.dumps({"skipped": "time budget", "length_kb": lk, "coverage": cov}), flush=True)
            continue
        L = lk * 1000
        max_seq = lk * 1024
        cfg = cudapoa.make_config(max_seq, cov, 256, "adaptive_band", adaptive_storage_factor=6.0)
        b = cudapoa.CudaPoaBatch(cov, max_seq, int(free_b * 0.90), config=cfg)
        n = min(b.max_poas, max(b.resident_windows, 1))
        n = max(1, min(n, 1480))
        win_nseq, seq_len, data = synth.poa_windows(n, L, cov, L // 50, L // 100, L // 100, seed0=1000, max_read_len=max_seq)
        rc, added = b.add_poa_groups_flat(win_nseq, seq_len, data)
        assert rc == 0 and added == n
        b.generate_poa()
        c, cv, lens, st = b.get_consensus_arrays()
        b.launch()
        b.sync()
        ms = b.last_kernel_ms()
        cells = b.last_cells()
        line = {"length_kb": lk, "coverage": cov, "windows": n, "kernel_ms": ms, "windows_per_s": n / (ms / 1e3), "dp_cells": cells,
                "dp_cells_per_s": cells / (ms / 1e3), "score_bytes": b.score_bytes,
                "roofline_frac": cells * b.score_bytes / (ms / 1e3) / 1e9 / peak, "status_hist": np.bincount(st, minlength=13).tolist()}
        ours = [bytes(c[i, :lens[i]]).decode() for i in range(n)]
        b.close()
        if args.ref:
            import ref_lib
            if ref_lib.have_gwref():
                k = min(n, 64)
                sub_nseq = win_nseq[:k]
                sub_len = seq_len[:k * cov]
                sub_data = data[:int(sub_len.sum()) + 1]
                r = ref_lib.ref_poa_run(sub_nseq, sub_len, sub_data, max_seq, cov, 256, 2, adaptive_storage_factor=6.0, mem_fraction=0.5,
                                        max_windows_per_batch=k)
                line["identical_to_reference_first_%d" % k] = bool(r["consensus"] == ours[:k] and list(r["status"]) == list(st[:k]))
        print(json.dumps(line), flush=True)
