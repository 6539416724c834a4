"""One process per GPU: shard bookkeeping and the bootstrap of the library's communicator.

sandbag shards gene-pair tiles (and the rank / bit-slice prep over cells), cyclone shards contiguous cell
ranges.  Counts / thresholds / decisions are per pair and scores are per cell, so no data-path collective
exists; what crosses GPUs (bit planes, compacted marker keys, score columns) is exchanged by the CUDA
library itself with NCCL on its own stream (pp_sandbag_dist / pp_cyclone_dist, csrc/pp_comm.cu).
torch.distributed is used for exactly two host-side things: carrying the 128-byte NCCL unique id from rank 0
to the other ranks, and (local_shard only) exchanging sample names as fixed-width bytes.
"""
import numpy as np

_comm_ready = {}  # device index -> (rank, world) the library communicator was created for


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


def shard_sizes(n_cells, world_size):
    return np.array([cell_shard(n_cells, r, world_size)[1] - cell_shard(n_cells, r, world_size)[0]
                     for r in range(world_size)], dtype=np.int64)


def broadcast_object(obj, src=0):
    """Small host object from `src` to every rank (unique id, seed)."""
    import torch.distributed as dist
    box = [obj]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def ensure_comm(device=None):
    """Make sure this process's library context holds an NCCL communicator spanning the default process
    group (collective on first use; one communicator per process for its lifetime).  -> (rank, world)."""
    from . import _ffi
    rank, ws = world()
    if ws == 1:
        return rank, ws
    if device is None:
        device = _ffi._current_device()
    have = _comm_ready.get(device)
    if have == (rank, ws):
        return rank, ws
    if have is not None:
        _ffi.comm_destroy(device)
    uid = _ffi.comm_unique_id() if rank == 0 else None
    uid = broadcast_object(uid, 0)
    _ffi.comm_init(uid, rank, ws, device)
    _comm_ready[device] = (rank, ws)
    return rank, ws


def drop_comm(device=None):
    """Destroy the library communicator (before the process group is torn down)."""
    from . import _ffi
    if device is None:
        device = _ffi._current_device()
    if _comm_ready.pop(device, None) is not None:
        _ffi.comm_destroy(device)


def allgather_ints(values):
    """A few ints from every rank -> int64[world, len(values)]."""
    import torch
    import torch.distributed as dist
    rank, ws = world()
    vals = np.asarray(values, dtype=np.int64).reshape(-1)
    if ws == 1:
        return vals[None, :]
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.from_numpy(vals).to(dev)
    parts = [torch.zeros_like(t) for _ in range(ws)]
    dist.all_gather(parts, t)
    return np.stack([p.cpu().numpy() for p in parts])


def allgather_counts(n):
    """int from every rank -> int64[world]."""
    return allgather_ints([int(n)])[:, 0]


def allgather_names(names):
    """Concatenate every rank's sample names in rank order.  The names travel as one fixed-width byte matrix per
    rank (UTF-8, padded to the longest name of any rank) through a single tensor all-gather -- no pickling of
    Python lists; the result is a numpy str array."""
    import torch
    import torch.distributed as dist
    rank, ws = world()
    if ws == 1:
        return names
    enc = np.char.encode(np.asarray(names, dtype=str), "utf-8") if len(names) else np.empty(0, dtype="S1")
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    meta = torch.tensor([len(enc), enc.dtype.itemsize], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(ws)]
    dist.all_gather(metas, meta)
    counts = [int(m[0].item()) for m in metas]
    width = max(max(int(m[1].item()) for m in metas), 1)
    rows = max(max(counts), 1)
    buf = np.zeros((rows, width), dtype=np.uint8)
    if len(enc):
        buf[:len(enc), :enc.dtype.itemsize] = np.ascontiguousarray(enc).view(np.uint8).reshape(len(enc), -1)
    t = torch.from_numpy(buf).to(dev)
    parts = [torch.empty_like(t) for _ in range(ws)]
    dist.all_gather(parts, t)
    cat = np.concatenate([p[:c].cpu().numpy() for p, c in zip(parts, counts)])
    return np.char.decode(np.ascontiguousarray(cat).view("S{}".format(width)).ravel(), "utf-8")
