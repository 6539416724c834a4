"""One process per GPU: shard bookkeeping and the single collective each path needs.

sandbag shards gene-pair tiles (tile index mod world size, inside the library), cyclone shards
contiguous cell ranges.  Counts/thresholds/decisions are per pair and scores are per cell, so no
data-path collective exists; the only exchange is the final all-gather of the compacted marker
lists / score columns (NCCL when the process group is NCCL, gloo in the CPU tests).
"""
import numpy as np


def world():
    """(rank, world_size) of the default process group, (0, 1) when not initialised."""
    try:
        import torch.distributed as dist
    except ImportError:  # pragma: no cover
        return 0, 1
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def cell_shard(n_cells, rank, world_size):
    """Contiguous, balanced [begin, end) of `rank` (first n % world shards get one extra cell)."""
    base, extra = divmod(n_cells, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def _device():
    import torch
    import torch.distributed as dist
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def allgather_varlen(arrays):
    """All-gather a tuple of equally long 1-D numpy arrays whose length differs per rank.
    Returns, per input array, the concatenation over ranks in rank order."""
    import torch
    import torch.distributed as dist
    rank, ws = world()
    if ws == 1:
        return [np.asarray(a) for a in arrays]
    dev = _device()
    n = torch.tensor([len(arrays[0])], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(ws)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    mx = max(max(counts), 1)
    out = []
    for a in arrays:
        a = np.ascontiguousarray(a)
        view = a.view(np.int64) if a.dtype == np.uint64 else a  # torch has no uint64 collectives
        t = torch.zeros(mx, dtype=torch.from_numpy(view[:0]).dtype, device=dev)
        t[:len(view)] = torch.from_numpy(view).to(dev)
        parts = [torch.empty_like(t) for _ in range(ws)]
        dist.all_gather(parts, t)
        cat = np.concatenate([p[:c].cpu().numpy() for p, c in zip(parts, counts)])
        out.append(cat.view(np.uint64) if a.dtype == np.uint64 else cat)
    return out


def allgather_columns(local, n_total):
    """local: float64[C, n_local] of this rank's contiguous cell shard -> float64[C, n_total]."""
    import torch
    import torch.distributed as dist
    rank, ws = world()
    if ws == 1:
        return local
    dev = _device()
    base, extra = divmod(n_total, ws)
    mx = base + (1 if extra else 0)
    t = torch.zeros((local.shape[0], max(mx, 1)), dtype=torch.float64, device=dev)
    t[:, :local.shape[1]] = torch.from_numpy(np.ascontiguousarray(local)).to(dev)
    parts = [torch.empty_like(t) for _ in range(ws)]
    dist.all_gather(parts, t)
    cols = []
    for r, p in enumerate(parts):
        b, e = cell_shard(n_total, r, ws)
        cols.append(p[:, :e - b].cpu().numpy())
    return np.concatenate(cols, axis=1)


def allgather_names(names):
    """Concatenate every rank's list of sample names in rank order."""
    import torch.distributed as dist
    rank, ws = world()
    if ws == 1:
        return list(names)
    parts = [None] * ws
    dist.all_gather_object(parts, list(names))
    return [n for p in parts for n in p]
