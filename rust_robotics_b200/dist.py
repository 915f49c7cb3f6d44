"""Host-side plumbing for the sharded engine: one process per GPU, torch.distributed (any backend) as the control plane.
The data path (exact global sums, CDF/ancestry allgathers, map exchange) runs on NCCL inside libpfgpu."""
import ctypes as C


def shard_bounds(n_global, rank, world):
    """contiguous block [lo, hi) of the particle index space owned by `rank`"""
    if n_global % world != 0:
        raise ValueError("particle count must divide evenly over the ranks")
    nl = n_global // world
    return rank * nl, (rank + 1) * nl


def broadcast_unique_id(dist, make_id, rank, src=0):
    """rank `src` calls make_id() -> 128 bytes (an ncclUniqueId); every rank gets the same bytes back."""
    import torch
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == src:
        raw = bytes(make_id())
        if len(raw) != 128:
            raise ValueError("ncclUniqueId must be 128 bytes")
        buf = torch.tensor(list(raw), dtype=torch.uint8)
    dist.broadcast(buf, src)
    return bytes(buf.tolist())


def nccl_unique_id():
    from .api import _check, load_library
    L = load_library()
    buf = C.create_string_buffer(128)
    _check(L, L.pfgpu_nccl_unique_id(buf))
    return buf.raw


def max_over_ranks(dist, x):
    import torch
    t = torch.tensor([float(x)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
