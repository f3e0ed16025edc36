"""bench.py leg for the aligner workload (config C4): pairs/s of the banded Myers global aligner."""
import json
import os
import time

import numpy as np


def run_aligner_bench(args, wp, rank, world, local_rank, barrier, max_over_ranks, sum_over_ranks):
    import torch
    from genomeworks_b200 import _lib, cudaaligner, synth
    from bench import ClockSampler, measured_peaks
    L = _lib.lib()
    n = wp["windows"]
    ql, qd, tl, td = synth.aligner_pairs(n, wp["genome"], seed=1 + rank)
    qb, tb = bytes(qd), bytes(td)
    pairs, qo, to = [], 0, 0
    for i in range(n):
        pairs.append((qb[qo:qo + ql[i]], tb[to:to + tl[i]]))
        qo += int(ql[i])
        to += int(tl[i])
    stream = torch.cuda.Stream()
    al = cudaaligner.FixedBandAligner(wp["max_bw"], stream=stream, device_id=local_rank)

    def step():
        for q, t in pairs:
            al.add_alignment(q, t)
        al.align_all()
        al.sync_alignments(want_strings=False)

    launches0 = L.gwb200_kernel_launch_count()
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    kms, t0 = 0.0, time.perf_counter()
    l0 = L.gwb200_kernel_launch_count()
    for _ in range(args.steps):
        step()
        kms += al.last_kernel_ms()
    torch.cuda.synchronize()
    wall = max_over_ranks(time.perf_counter() - t0)
    kms = max_over_ranks(kms)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    timed_launches = L.gwb200_kernel_launch_count() - l0
    cells = al.last_cells()
    res = al.get_alignments()
    n_ok = sum(1 for r in res if r.status == 0)
    n_opt = sum(1 for r in res if r.is_optimal)
    total = sum_over_ranks(float(n))
    if rank == 0:
        peak, peak_src = measured_peaks()
        k_ms = kms / args.steps
        achieved = cells * 0.375 / (k_ms / 1e3) / 1e9
        seq_bytes = int(ql.sum() + tl.sum())
        line = {
            "metric": "cudaaligner_pairs_per_s", "value": total * args.steps / (kms / 1e3), "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": k_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 bit-vectors + int32 scores", "data": "synthetic",
            "config": {"workload": wp["name"], "pairs_per_gpu_per_step": n, "max_bandwidth": wp["max_bw"],
                       "l2": "band matrices written per step (%.2f GB) exceed the 126 MB L2" % (cells * 0.375 / 1e9),
                       "pairs_ok_last_step": n_ok, "pairs_optimal_last_step": n_opt,
                       "note": "`value` is device time of align_all's kernels (CUDA events); inputs are host buffers in both"},
            "dp_cells_per_step_per_gpu": cells, "dp_cells_per_s": cells * world / (k_ms / 1e3),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "peak_source": peak_src, "kernel": "myers_banded_kernel", "algorithmic_bytes_per_cell": 0.375,
                         "kernel_ms_per_launch": k_ms},
            "e2e": {"value": total * args.steps / wall, "unit": "pairs/s", "h2d_bytes_per_step": seq_bytes + 20 * n,
                    "d2h_bytes_per_step": int(sum(len(r.actions) for r in res)) * 5 + 8 * n},
            "gpu_launches": int(timed_launches), "gpu_launches_total": int(L.gwb200_kernel_launch_count() - launches0), "clocks": clocks,
            "cpu_baseline": {"value": None, "unit": "pairs/s", "cores": 0, "kind": "reference",
                             "sample": "no CPU implementation of this path is named by the reference (SURVEY.md 8d); the reference's GPU kernel is compared by tests/tools/gpu_reference.py"},
        }
        print(json.dumps(line))
