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


def scatter_windows(win_nseq, seq_len, seq_data, dist, device=None, src=0, band_width=256):
    """Input scatter: rank `src` holds a window list in the flat format (win_nseq, seq_len, seq_data as
    cudapoa.add_poa_groups_flat takes them; the other ranks pass None); every rank receives the windows of its shard of the
    cost-balanced partition. One header broadcast, then three padded `dist.scatter` calls (NCCL over NVLink when `device` is a
    CUDA device, gloo on CPU tensors in the CPU tests).

    Returns (global index of each local window, win_nseq, seq_len, seq_data) as numpy arrays; n_total is len of the list."""
    import torch
    world = dist.get_world_size()
    rank = dist.get_rank()
    header = torch.zeros((world, 3), dtype=torch.int64, device=device)
    shards = packs = None
    n_total = torch.zeros(1, dtype=torch.int64, device=device)
    if rank == src:
        win_nseq = np.ascontiguousarray(win_nseq, dtype=np.int32)
        seq_len = np.ascontiguousarray(seq_len, dtype=np.int32)
        seq_data = np.ascontiguousarray(seq_data, dtype=np.uint8)
        shards = partition_units(window_costs(win_nseq, seq_len, band_width), world)
        r_end = np.cumsum(win_nseq)
        r_start = r_end - win_nseq
        b_csum = np.concatenate([[0], np.cumsum(seq_len.astype(np.int64))])
        packs = []
        for r in range(world):
            idx = shards[r]
            ns = win_nseq[idx]
            sl = np.concatenate([seq_len[r_start[w]:r_end[w]] for w in idx]) if len(idx) else np.zeros(0, np.int32)
            sd = np.concatenate([seq_data[b_csum[r_start[w]]:b_csum[r_end[w]]] for w in idx]) if len(idx) else np.zeros(0, np.uint8)
            packs.append((idx.astype(np.int64), ns, sl, sd))
            header[r, 0], header[r, 1], header[r, 2] = len(idx), len(sl), len(sd)
        n_total[0] = len(win_nseq)
    dist.broadcast(header, src=src)
    dist.broadcast(n_total, src=src)
    hdr = header.cpu().numpy()
    mx = hdr.max(axis=0)

    def scat(dtype, width, field):
        mine = torch.zeros(max(int(width), 1), dtype=dtype, device=device)
        lst = None
        if rank == src:
            lst = []
            for r in range(world):
                t = torch.zeros(max(int(width), 1), dtype=dtype)
                a = packs[r][field]
                t[:len(a)] = torch.from_numpy(np.ascontiguousarray(a))
                lst.append(t.to(device) if device is not None else t)
        dist.scatter(mine, lst, src=src)
        return mine.cpu().numpy()

    idx = scat(torch.int64, mx[0], 0)[:hdr[rank, 0]]
    ns = scat(torch.int32, mx[0], 1)[:hdr[rank, 0]]
    sl = scat(torch.int32, mx[1], 2)[:hdr[rank, 1]]
    sd = scat(torch.uint8, mx[2], 3)[:hdr[rank, 2]]
    return idx, ns.astype(np.int32), sl.astype(np.int32), np.concatenate([sd, np.zeros(1, np.uint8)]), int(n_total.item())


def sharded_consensus(win_nseq, seq_len, seq_data, make_batch, dist, device=None, src=0, band_width=256):
    """One fixed window list on rank `src` -> scatter -> every rank runs its shard through its own Batch (as many
    generate_poa rounds as its capacity needs) -> fixed-stride gather of consensus / coverage / length / status back to `src`
    in input order. make_batch() returns a cudapoa.CudaPoaBatch for this rank. Returns a dict on `src`, None elsewhere."""
    import torch
    idx, ns, sl, sd, n_total = scatter_windows(win_nseq, seq_len, seq_data, dist, device=device, src=src, band_width=band_width)
    batch = make_batch()
    mc = batch.config.max_consensus_size
    n_local = len(ns)
    cons = np.zeros((n_local, mc), dtype=np.uint8)
    cov = np.zeros((n_local, mc), dtype=np.uint16)
    meta = np.zeros((n_local, 2), dtype=np.int32)  # length, status
    done, r_off, b_off = 0, 0, 0
    while done < n_local:
        k = min(batch.max_poas, n_local - done)
        nr = int(ns[done:done + k].sum())
        nb = int(sl[r_off:r_off + nr].astype(np.int64).sum())
        batch.reset()
        rc, added = batch.add_poa_groups_flat(ns[done:done + k], sl[r_off:r_off + nr], sd[b_off:b_off + nb + 1])
        if added == 0:
            raise RuntimeError("a window does not fit the batch (status %d)" % rc)
        if added < k:  # capacity reached earlier than estimated: take what fits
            k = added
            nr = int(ns[done:done + k].sum())
            nb = int(sl[r_off:r_off + nr].astype(np.int64).sum())
        batch.generate_poa()
        c, cv, lens, st = batch.get_consensus_arrays()
        cons[done:done + k], cov[done:done + k] = c[:k], cv[:k]
        meta[done:done + k, 0], meta[done:done + k, 1] = lens[:k], st[:k]
        done += k
        r_off += nr
        b_off += nb
    batch.close()
    to = (lambda a: torch.from_numpy(a).to(device)) if device is not None else torch.from_numpy
    g_cons = gather_fixed_stride(to(cons), idx, n_total, dist, device=device, dst=src)
    # coverage travels as bytes: 16-bit integers are not a collective dtype of every backend (gloo rejects them)
    g_cov = gather_fixed_stride(to(cov.view(np.uint8)), idx, n_total, dist, device=device, dst=src)
    g_meta = gather_fixed_stride(to(meta), idx, n_total, dist, device=device, dst=src)
    if dist.get_rank() != src:
        return None
    g_meta = g_meta.cpu().numpy()
    return dict(consensus=g_cons.cpu().numpy(), coverage=np.ascontiguousarray(g_cov.cpu().numpy()).view(np.uint16), lengths=g_meta[:, 0], status=g_meta[:, 1])
