#!/usr/bin/env python
"""bench.py -- the measurement contract.

Metric (BASELINE.json): POA consensus windows/sec on synthetic 10 kb x 32-read windows (config C3: adaptive band 256,
int32 scores; adaptive_storage_factor 3.0 -- the reference's default 2.0 yields exceeded_adaptive_banded_matrix_size for
every window of this workload on both this engine and the reference (SURVEY.md fact 3 / DESIGN.md), 3.0 is the smallest
integer factor at which all windows succeed; the factor only sizes the per-window score slab, results are identical). One "step" = one pass of the hot path over one batch
of `--windows` windows per GPU. Windows shard embarrassingly: every rank owns a Batch and its own windows (weak scaling),
NCCL is used only for the barrier / max-over-ranks timing and the result gather after the timed region.

  value : windows/s, inputs packed and resident in HBM before the timed region (K x [launch, sync], CUDA events on the
          batch stream, max over ranks)
  e2e   : windows/s through the public API with HOST buffers every step: add_poa_group packing, H2D, kernel, D2H of
          consensus + coverage (wall clock around barrier + synchronize, max over ranks)
  roofline: algorithmic bytes = executed DP cells x sizeof(ScoreT) / kernel time, vs MEASURED_PEAKS hbm_gbs
  cpu_baseline / --impl reference: 3rdparty/spoa (oracle/_ref/libspoa_ref.so, unmodified) on the host cores.

Other workloads for development: --workload c2 (1 kb x 16, static band, 1024 windows), c4 (aligner 10k x 10k, bw 1024).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


# dram bytes (read + write) per window of the dominant kernel, from the committed ncu --set full captures (profiles/)
NCU_DRAM_BYTES_PER_WINDOW = {"c3": (1.810265e9 + 15.109566e9) / 16, "c2": (4.804323e9 + 9.375816e9) / 1024}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi SM clocks + throttle reasons during the timed region."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split("\n")[0]
                f = [x.strip() for x in out.split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def workload_params(name, windows):
    if name == "c3":
        return dict(kind="poa", name="C3: cudapoa long-read consensus, 10 kb x 32 reads/window, adaptive band 256, int32 scores, "
                    "adaptive_storage_factor 3.0", backbone=10000, reads=32, mut=200, ins=100, dele=100, max_seq=10240, band=256,
                    band_mode="adaptive_band", factor=3.0, windows=windows or 296)
    if name == "c2":
        return dict(kind="poa", name="C2: cudapoa short-read consensus, 1 kb x 16 reads/window, static band 256, int16 scores",
                    backbone=1000, reads=16, mut=20, ins=10, dele=10, max_seq=1024, band=256, band_mode="static_band", factor=2.0,
                    windows=windows or 1024)
    if name == "c4":
        return dict(kind="aligner", name="C4: cudaaligner global, 10000 x 10000 bp, Myers banded (max_bandwidth 1024)", genome=10000,
                    max_bw=1024, windows=windows or 512)
    raise SystemExit("unknown workload " + name)


def spoa_sample(wp, n_windows, seed0, threads=0, reads=None):
    """Times unmodified spoa on n_windows windows of the workload. For long-read workloads a full 10 kb x 32-read window costs
    spoa ~3e9 DP cells (minutes per window per core), so the bounded sample uses the first `reads` reads of each window and
    the windows/s figure is extrapolated with spoa's own DP-cell count: cells(full window) is estimated from the sample's
    graph growth (nodes grow linearly with the number of reads fused)."""
    import ref_lib
    from genomeworks_b200 import synth
    full_reads = wp["reads"]
    reads = full_reads if reads is None else min(reads, full_reads)
    win_nseq, seq_len, data = synth.poa_windows(n_windows, wp["backbone"], full_reads, wp["mut"], wp["ins"], wp["dele"], seed0=seed0,
                                                max_read_len=wp["max_seq"])
    if reads < full_reads:
        # keep the first `reads` reads of every window
        sl = seq_len.reshape(n_windows, full_reads)
        offs = np.concatenate([[0], np.cumsum(seq_len)]).astype(np.int64)
        keep = []
        for w in range(n_windows):
            b = int(offs[w * full_reads])
            e = int(offs[w * full_reads + reads])
            keep.append(data[b:e])
        data = np.concatenate(keep + [np.zeros(1, np.uint8)])
        seq_len = np.ascontiguousarray(sl[:, :reads]).reshape(-1)
        win_nseq = np.full(n_windows, reads, dtype=np.int32)
    r = ref_lib.spoa_consensus(win_nseq, seq_len, data, n_threads=threads, want_strings=False)
    r["windows"] = n_windows
    r["reads_used"] = reads
    L = float(wp["backbone"])
    if reads < full_reads:
        # spoa cells = sum_r nodes_r * len_r; nodes_r ~ L * (1 + g * (r - 1)) with growth g fitted on the sample
        a = reads - 1
        cells_pw = r["cells"] / n_windows
        g = max(0.0, (cells_pw / (L * L) - a) / max(a * (a - 1) / 2.0, 1e-9))
        A = full_reads - 1
        full_cells = L * L * (A + g * A * (A - 1) / 2.0)
        r["scale"] = cells_pw / full_cells
    else:
        r["scale"] = 1.0
    r["windows_per_s"] = (n_windows / r["seconds"]) * r["scale"]
    return r


def run_reference_arm(args, wp, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (3rdparty/spoa, unmodified, oracle/_ref) on
    the box's host cores. Rank 0 only."""
    if rank != 0:
        return
    import ref_lib
    if wp["kind"] != "poa":
        print(json.dumps({"impl": "reference", "unavailable": "no CPU reference implementation is named for the aligner path"}))
        return
    if not ref_lib.have_spoa():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libspoa_ref.so is not built"}))
        return
    cores = os.cpu_count() or 1
    long_reads = wp["backbone"] >= 5000
    # bounded sample per step: one window per host thread; long-read windows are cut to their first 3 reads (see spoa_sample)
    per_step = max(1, min(cores, 256)) if long_reads else max(8, 8 * cores)
    sample_reads = 3 if long_reads else None
    times, cells, wps = [], 0.0, []
    for it in range(args.warmup + args.steps):
        r = spoa_sample(wp, per_step, 1000 + it * per_step, threads=cores, reads=sample_reads)
        if it >= args.warmup:
            times.append(r["seconds"])
            cells += r["cells"]
            wps.append(r["windows_per_s"])
    total = sum(times)
    value = float(np.mean(wps))
    line = {
        "impl": "reference", "metric": "poa_consensus_windows_per_s", "value": value, "unit": "windows/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(len(times), 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int16 (spoa AVX2 SIMD)", "data": "synthetic",
        "config": {"workload": wp["name"], "windows_per_step": per_step, "engine": "3rdparty/spoa kNW linear gaps, full DP, all host threads"},
        "cpu_baseline": {"value": value, "unit": "windows/s", "cores": cores, "kind": "reference",
                         "sample": "%d windows per step x %d steps, %s, spoa DP cells/s %.3e" % (
                             per_step, len(times), ("first %d of %d reads per window, windows/s extrapolated by spoa DP cells" % (sample_reads, wp["reads"]))
                             if sample_reads else "full windows", cells / total)},
        "e2e": {"value": value, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c3", "c2", "c4"])
    ap.add_argument("--windows", type=int, default=0, help="windows (or pairs) per GPU per step")
    ap.add_argument("--factor", type=float, default=0.0, help="override adaptive_storage_factor of the workload (C3 default 3.0)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wp = workload_params(args.workload, args.windows)

    if args.impl == "reference":
        run_reference_arm(args, wp, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: this engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    from genomeworks_b200 import _lib
    L = _lib.lib()
    if wp["kind"] == "aligner":
        from bench_aligner import run_aligner_bench
        run_aligner_bench(args, wp, rank, world, local_rank, barrier, max_over_ranks, sum_over_ranks)
        if world > 1:
            dist.destroy_process_group()
        return

    from genomeworks_b200 import cudapoa, synth
    if args.factor > 0:
        wp["factor"] = args.factor
        wp["name"] = wp["name"].replace("adaptive_storage_factor 3.0", "adaptive_storage_factor %g" % args.factor)
    cfg = cudapoa.make_config(wp["max_seq"], wp["reads"], wp["band"], wp["band_mode"], adaptive_storage_factor=wp["factor"])
    stream = torch.cuda.Stream()
    free_b, _ = torch.cuda.mem_get_info()
    batch = cudapoa.CudaPoaBatch(wp["reads"], wp["max_seq"], int(free_b * 0.92), config=cfg, device_id=local_rank, stream=stream)
    n_win = args.windows or wp["windows"]
    if args.workload == "c3" and not args.windows:
        # one full batch per GPU: as many windows as the batch holds, rounded to a multiple of the SM count
        # and capped at what the device keeps resident at once (one wave)
        n_win = min(batch.max_poas, max(batch.resident_windows, 148))
        n_win = n_win // 148 * 148 if n_win >= 148 else n_win
    if batch.max_poas < n_win:
        raise SystemExit("batch capacity %d < requested windows %d" % (batch.max_poas, n_win))
    # every rank owns its own windows (seeds disjoint across ranks): weak scaling, no data-path collective
    win_nseq, seq_len, data = synth.poa_windows(n_win, wp["backbone"], wp["reads"], wp["mut"], wp["ins"], wp["dele"],
                                                seed0=1000 + rank * n_win, max_read_len=wp["max_seq"])

    launches0 = L.gwb200_kernel_launch_count()

    # ---------------- device-resident timing (`value`) ----------------
    rc, added = batch.add_poa_groups_flat(win_nseq, seq_len, data)
    assert rc == 0 and added == n_win
    batch.upload()
    batch.sync()
    for _ in range(args.warmup):
        batch.launch()
        batch.sync()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    launches_before_timed = L.gwb200_kernel_launch_count()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(args.steps):
            batch.launch()
        ev1.record(stream)
    stream.synchronize()
    barrier()
    timed_launches = L.gwb200_kernel_launch_count() - launches_before_timed
    dev_ms = ev0.elapsed_time(ev1)
    dev_ms = max_over_ranks(dev_ms)
    clocks = sampler.stop() if rank == 0 else None
    cells = batch.last_cells()
    last_kernel_ms = batch.last_kernel_ms()
    c, cov, lens, st = batch.get_consensus_arrays()
    n_ok = int((st == 0).sum())
    total_windows = sum_over_ranks(float(n_win))
    value = total_windows * args.steps / (dev_ms / 1e3)

    # ---------------- end-to-end through the public API with host buffers (`e2e`) ----------------
    # sequences (padded to 4 B per read) + window descriptors + read lengths; base weights are unit weights here and are
    # set on the device (no host traffic)
    h2d = int(((seq_len + 3) // 4 * 4).sum()) + 24 * n_win + 4 * len(seq_len)
    d2h = n_win * cfg.max_consensus_size * 3 + n_win * 20
    for _ in range(max(1, min(args.warmup, 2))):
        batch.reset()
        batch.add_poa_groups_flat(win_nseq, seq_len, data)
        batch.generate_poa()
        batch.get_consensus_arrays()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        batch.reset()
        batch.add_poa_groups_flat(win_nseq, seq_len, data)
        batch.generate_poa()
        c2, cov2, lens2, st2 = batch.get_consensus_arrays()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = total_windows * args.steps / e2e_s
    assert (lens2 == lens).all() and (st2 == st).all()

    # result gather over NCCL (outside the timed regions): per-window status + consensus length to rank 0
    if world > 1:
        mine = torch.from_numpy(np.stack([st.astype(np.int32), lens.astype(np.int32)], 1)).cuda()
        gathered = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
        dist.gather(mine, gathered, dst=0)
        if rank == 0:
            n_ok = int(sum(int((g[:, 0] == 0).sum().item()) for g in gathered))

    launches = L.gwb200_kernel_launch_count() - launches0

    if rank == 0:
        peak, peak_src = measured_peaks()
        sb = batch.score_bytes
        k_ms = dev_ms / args.steps
        achieved = cells * sb / (k_ms / 1e3) / 1e9
        line = {
            "metric": "poa_consensus_windows_per_s", "value": value, "unit": "windows/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32" if sb == 4 else "int16", "data": "synthetic",
            "config": {"workload": wp["name"], "windows_per_gpu_per_step": n_win, "reads_per_window": wp["reads"],
                       "band_mode": wp["band_mode"], "band_width": wp["band"], "scores": "gap -8 mismatch -6 match 8",
                       "parallelism": "windows sharded over %d GPU(s), one Batch per rank" % world,
                       "l2": "score matrices written per step (%.1f GB) exceed the 126 MB L2; no explicit flush" % (cells * sb / 1e9),
                       "windows_ok_last_step": n_ok},
            "dp_cells_per_step_per_gpu": cells, "dp_cells_per_s": cells * world / (k_ms / 1e3),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_DRAM_BYTES_PER_WINDOW.get(args.workload, 0) * n_win or None,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum per window from the ncu --set full captures in "
                                           "profiles/r01_poa_v2_{c3,c2}_ncu.md (16 / 1024 windows), scaled to this launch's window count",
                         "algorithmic_bytes": cells * sb,
                         "peak_source": peak_src, "kernel": "poa_window_kernel_v2", "algorithmic_bytes_per_cell": sb,
                         "kernel_ms_per_launch": k_ms},
            "e2e": {"value": e2e_value, "unit": "windows/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(timed_launches), "gpu_launches_total": int(launches), "clocks": clocks,
        }
        if not args.no_cpu_baseline:
            try:
                import ref_lib
                if ref_lib.have_spoa():
                    cores = os.cpu_count() or 1
                    long_reads = wp["backbone"] >= 5000
                    ns = max(1, min(cores, 256)) if long_reads else 8 * cores
                    r = spoa_sample(wp, ns, 1000, threads=cores, reads=3 if long_reads else None)
                    line["cpu_baseline"] = {"value": r["windows_per_s"], "unit": "windows/s", "cores": cores, "kind": "reference",
                                            "sample": "%d windows of the same workload%s, unmodified 3rdparty/spoa (AVX2), %.1f s, %.3e spoa DP cells/s"
                                                      % (ns, (" cut to their first %d reads, windows/s extrapolated by spoa DP-cell count (x%.4f)"
                                                              % (r["reads_used"], r["scale"])) if long_reads else "", r["seconds"],
                                                         r["cells"] / r["seconds"])}
                else:
                    line["cpu_baseline"] = {"value": None, "unit": "windows/s", "cores": 0, "kind": "reference",
                                            "sample": "oracle/_ref/libspoa_ref.so not built"}
            except Exception as e:  # pragma: no cover
                line["cpu_baseline"] = {"value": None, "unit": "windows/s", "cores": 0, "kind": "reference", "sample": "failed: %r" % (e,)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
