"""Host-side sharding of independent units (POA windows / alignment pairs) over the ranks of one node.

The path has no exchange step: every rank owns a Batch / Aligner and its own units. This module holds the
only multi-rank logic there is: a cost-balanced static partition (largest units first, dealt in snake order, like the
reference's per-batch scheduling by size, aligner_global_myers_banded.cpp:306-309) and the gather of per-unit results back
into input order over torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np


def partition_units(costs, world_size):
    """Returns a list of index arrays, one per rank: units sorted by cost descending and dealt in snake order."""
    costs = np.asarray(costs)
    order = np.argsort(-costs, kind="stable")
    shards = [[] for _ in range(world_size)]
    for k, idx in enumerate(order):
        rnd, pos = divmod(k, world_size)
        r = pos if rnd % 2 == 0 else world_size - 1 - pos
        shards[r].append(int(idx))
    return [np.array(sorted(s), dtype=np.int64) for s in shards]


def window_costs(win_nseq, seq_len, band_width=256):
    """Cost model of a POA window: sum of read lengths x band (DP cells up to a constant)."""
    win_nseq = np.asarray(win_nseq)
    ends = np.cumsum(win_nseq)
    starts = ends - win_nseq
    csum = np.concatenate([[0], np.cumsum(np.asarray(seq_len, dtype=np.int64))])
    return (csum[ends] - csum[starts]) * band_width


def gather_fixed_stride(local_rows, local_index, n_total, dist, device=None, dst=0):
    """Gathers per-unit fixed-stride rows (e.g. consensus buffers) to rank `dst` and restores input order.

    local_rows: torch tensor [n_local, stride]; local_index: the global unit index of each local row.
    Returns the [n_total, stride] tensor on rank dst, None elsewhere."""
    import torch
    world = dist.get_world_size()
    rank = dist.get_rank()
    n_local = torch.tensor([local_rows.shape[0]], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local)
    counts = [int(c.item()) for c in counts]
    mx = max(counts)
    stride = local_rows.shape[1]
    pad_rows = torch.zeros((mx, stride), dtype=local_rows.dtype, device=device)
    pad_rows[:local_rows.shape[0]] = local_rows
    pad_idx = torch.full((mx,), -1, dtype=torch.int64, device=device)
    pad_idx[:local_rows.shape[0]] = torch.as_tensor(local_index, dtype=torch.int64, device=device)
    rows = [torch.zeros_like(pad_rows) for _ in range(world)] if rank == dst else None
    idxs = [torch.zeros_like(pad_idx) for _ in range(world)] if rank == dst else None
    dist.gather(pad_rows, rows, dst=dst)
    dist.gather(pad_idx, idxs, dst=dst)
    if rank != dst:
        return None
    out = torch.zeros((n_total, stride), dtype=local_rows.dtype, device=device)
    for r in range(world):
        k = counts[r]
        out[idxs[r][:k]] = rows[r][:k]
    return out
