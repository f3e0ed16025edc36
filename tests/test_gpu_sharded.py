"""Multi-rank result parity: one fixed window list on rank 0 is scattered over 2 ranks (sharding.scatter_windows), every rank runs
its shard through its own Batch, the fixed-stride consensus rows are gathered back (sharding.gather_fixed_stride) and must equal
the 1-GPU output in input order, bit for bit. Backend: NCCL when the box has >= 2 GPUs (one rank per GPU), else gloo with both
ranks computing on cuda:0 (the collectives then run on CPU tensors; the sharding logic and the engine are the same)."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, backend, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from genomeworks_b200 import cudapoa, sharding, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = rank if backend == "nccl" else 0
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
        device = torch.device("cuda", dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
        device = None
    cfg = cudapoa.make_config(2048, 12, 256, "adaptive_band")
    # windows of different sizes so that the cost-balanced partition is not the identity
    parts = [synth.poa_windows(9, 1500, 12, 30, 15, 15, seed0=11, max_read_len=2048), synth.poa_windows(14, 600, 7, 12, 6, 6, seed0=77, max_read_len=2048),
             synth.poa_windows(8, 1900, 10, 38, 19, 19, seed0=5, max_read_len=2048)]
    win_nseq = np.concatenate([p[0] for p in parts])
    seq_len = np.concatenate([p[1] for p in parts])
    seq_data = np.concatenate([p[2][:-1] for p in parts] + [np.zeros(1, np.uint8)])

    def make_batch():
        # a deliberately small batch: every rank needs several generate_poa rounds for its shard
        return cudapoa.CudaPoaBatch(12, 2048, 600 << 20, config=cfg, device_id=dev)

    if rank == 0:
        out = sharding.sharded_consensus(win_nseq, seq_len, seq_data, make_batch, dist, device=device)
        b = cudapoa.CudaPoaBatch(12, 2048, 8 << 30, config=cfg, device_id=dev)
        rc, added = b.add_poa_groups_flat(win_nseq, seq_len, seq_data)
        assert rc == 0 and added == len(win_nseq)
        b.generate_poa()
        c, cov, lens, st = b.get_consensus_arrays()
        b.close()
        ok = (out["status"] == st).all() and (out["lengths"] == lens).all() and (st == 0).all()
        for w in range(len(win_nseq)):
            ok = ok and (out["consensus"][w, :lens[w]] == c[w, :lens[w]]).all() and (out["coverage"][w, :lens[w]] == cov[w, :lens[w]]).all()
        q.put(bool(ok))
    else:
        sharding.sharded_consensus(None, None, None, make_batch, dist, device=device)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_consensus_equals_single_gpu():
    import torch.multiprocessing as mp
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30300 + (os.getpid() % 300)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, backend, q)) for r in range(2)]
    for p in procs:
        p.start()
    import queue as _queue
    ok = None
    for _ in range(300):
        try:
            ok = q.get(timeout=1)
            break
        except _queue.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):
                break  # a rank died: fail now instead of waiting for the timeout
    for p in procs:
        p.join(timeout=60)
        if p.is_alive():
            p.terminate()
    assert ok is not None, "a rank failed before producing a result (backend %s)" % backend
    assert ok, "sharded consensus differs from the single-GPU result (backend %s)" % backend
