"""Sharding of independent units (frame pairs / windows) over ranks -- the reference's only parallel
axis (`#pragma omp parallel for` over windows, slow_flow.cpp:706; over folders, adaptiveFR.cpp:245).
No data-path collective exists: ranks only agree on the timing (max) and the unit count (sum)."""


def shard_range(n_units, rank, world):
    """Contiguous block of units for `rank`: sizes differ by at most one, earlier ranks get the extras.
    Contiguity matters: consecutive frame pairs share a frame, which a rank uploads once."""
    base, extra = divmod(int(n_units), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _reduce(value, op_name):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=getattr(dist.ReduceOp, op_name))
    return float(t.item())


def max_over_ranks(value):
    return _reduce(value, "MAX")


def sum_over_ranks(value):
    return int(round(_reduce(value, "SUM")))
