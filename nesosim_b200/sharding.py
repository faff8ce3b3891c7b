"""Partitioning of the naturally parallel axes over the GPUs of one box: calibration-ensemble
members and seasons are independent, so every rank simply takes a contiguous block -- there is no collective on the
data path.  The only communication is an optional gather of small per-member results at the end."""
import numpy as np


def member_range(n_members, rank, world):
    """Members [lo, hi) of rank `rank`: contiguous blocks, sizes differing by at most one, larger blocks first."""
    base, extra = divmod(int(n_members), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def season_assignment(years, rank, world):
    """Seasons (start years) of rank `rank`: round-robin, as in run_multiseason.py's independent yearly runs."""
    return [y for i, y in enumerate(years) if i % world == rank]


def gather_member_results(local, n_members, rank, world, group=None):
    """All-gather small per-member results (e.g. a misfit per member, shape (m_local, ...)) into member order on
    every rank.  Works with any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist
    local = torch.as_tensor(local)
    sizes = [member_range(n_members, r, world)[1] - member_range(n_members, r, world)[0] for r in range(world)]
    pad = max(sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[:local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)


def shard_params(params, rank, world):
    """This rank's rows of an (M, 4) parameter table."""
    p = np.asarray(params, dtype=np.float64).reshape(-1, 4)
    lo, hi = member_range(len(p), rank, world)
    return p[lo:hi]
