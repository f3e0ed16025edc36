"""Test-side measurement aid: run a BASELINE workload through this engine and through the unmodified reference kernels
(oracle/_ref/libgwref.so, rebuilt for sm_100a) on the same GPU and inputs; print one JSON line with both rates and whether the
outputs are identical. The reference is the checker here, never the thing shipped; bench.py does not use it.
usage: gpu_reference.py {c2|c3|c4} [--windows N] [--factor F]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import ref_lib

ap = argparse.ArgumentParser()
ap.add_argument("workload", choices=["c2", "c3", "c4"])
ap.add_argument("--windows", type=int, default=0)
ap.add_argument("--factor", type=float, default=3.0)
args = ap.parse_args()
if not ref_lib.have_gwref():
    raise SystemExit("oracle/_ref/libgwref.so is not built (python -c 'import __graft_entry__ as g; g.build()' where /root/reference exists)")

if args.workload == "c4":
    from genomeworks_b200 import cudaaligner, synth
    n = args.windows or 512
    ql, qd, tl, td = synth.aligner_pairs(n, 10000, seed=1)
    al = cudaaligner.FixedBandAligner(1024)
    best = None
    for _ in range(3):
        al.reset()
        qo = to = 0
        for i in range(n):
            al.add_alignment(bytes(qd[qo:qo + ql[i]]), bytes(td[to:to + tl[i]]))
            qo += int(ql[i])
            to += int(tl[i])
        t0 = time.perf_counter()
        al.align_all()
        al.sync_alignments()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    res = al.get_alignments()
    kernel_ms = al.last_kernel_ms()
    al.close()
    ref_lib.ref_aligner_run(ql, qd, tl, td, 1024, max_device_memory=32 << 30)
    rr = ref_lib.ref_aligner_run(ql, qd, tl, td, 1024, max_device_memory=32 << 30)
    same = all((r.convert_to_cigar(True) == rr["cigar_extended"][i]) and (int(r.is_optimal) == rr["is_optimal"][i]) for i, r in enumerate(res))
    print(json.dumps({"workload": "C4 aligner 10k x 10k, band 1024", "pairs": n, "ours_pairs_per_s_align_all_sync": n / best,
                      "ours_kernel_ms": kernel_ms, "reference_pairs_per_s_align_all_sync": n / (rr["timings"][1] / 1e3),
                      "identical_outputs": bool(same)}))
    sys.exit(0)

from genomeworks_b200 import cudapoa, synth
if args.workload == "c2":
    n = args.windows or 1024
    win_nseq, seq_len, data = synth.poa_windows(n, 1000, 16, 20, 10, 10, seed0=1000, max_read_len=1024)
    cfg = cudapoa.make_config(1024, 16, 256, "static_band")
    ref_args, factor, mem = (1024, 16, 256, 1), 2.0, 16 << 30
    name = "C2 POA 1 kb x 16, static band 256"
else:
    n = args.windows or 296
    factor = args.factor
    win_nseq, seq_len, data = synth.poa_windows(n, 10000, 32, 200, 100, 100, seed0=1000, max_read_len=10240)
    cfg = cudapoa.make_config(10240, 32, 256, "adaptive_band", adaptive_storage_factor=factor)
    ref_args, mem = (10240, 32, 256, 2), int(n * (factor * 32.5e6 + 20e6)) + (2 << 30)
    name = "C3 POA 10 kb x 32, adaptive band 256, adaptive_storage_factor %g" % factor
b = cudapoa.CudaPoaBatch(cfg.max_sequences_per_poa, cfg.max_sequence_size, mem, config=cfg)
best = None
for _ in range(3):
    b.reset()
    b.add_poa_groups_flat(win_nseq, seq_len, data)
    t0 = time.perf_counter()
    b.generate_poa()
    c, cov, lens, st = b.get_consensus_arrays()
    dt = time.perf_counter() - t0
    best = dt if best is None else min(best, dt)
kernel_ms = b.last_kernel_ms()
ours = [bytes(c[i, :lens[i]]).decode() for i in range(n)]
b.close()
ref_lib.ref_poa_run(win_nseq, seq_len, data, *ref_args, adaptive_storage_factor=factor, mem_fraction=0.5, max_windows_per_batch=n)
rr = ref_lib.ref_poa_run(win_nseq, seq_len, data, *ref_args, adaptive_storage_factor=factor, mem_fraction=0.5, max_windows_per_batch=n)
same = rr["consensus"] == ours and list(rr["status"]) == list(st) and all(list(a) == list(cov[i, :lens[i]]) for i, a in enumerate(rr["coverage"]))
print(json.dumps({"workload": name, "windows": n, "ours_windows_per_s_generate_get": n / best, "ours_kernel_ms": kernel_ms,
                  "reference_windows_per_s_generate_get": n / (rr["timings"][1] / 1e3), "reference_batches": int(rr["timings"][2]),
                  "windows_ok": int((st == 0).sum()), "identical_outputs": bool(same)}))
