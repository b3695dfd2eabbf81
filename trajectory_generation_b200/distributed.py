"""Trajectory-parallel sharding over the GPUs of one box.

Trajectories never interact (in the reference they are iterations of a Python loop with per-id
seeds: generation_type1.py:267,296; generation_type2.py:169,177,191), so rank r of W owns the
contiguous id block [r*B/W, (r+1)*B/W) and there is NO collective on the solve path.  The only
communication is the final gather of the finished rows -- the id-offset merge that
generation_traj/merge_datasets.py:41-47 does on CPU files.  Results are independent of W because the
noise seed is keyed by the global trajectory id.
"""
import numpy as np


def shard_range(num_traj, rank, world_size):
    """contiguous block of trajectory ids of ``rank``: sizes differ by at most one."""
    base, rem = divmod(int(num_traj), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def generate_sharded(generate_fn, x0, u0, scenarios, T, rank=0, world_size=1):
    """Run ``generate_fn(x0_local, u0_local, scenarios_local, T, traj_id0)`` on this rank's id block.
    Returns (lo, hi, local_result)."""
    lo, hi = shard_range(len(x0), rank, world_size)
    sc = scenarios.slice(lo, hi) if hasattr(scenarios, "slice") else scenarios[lo:hi]
    return lo, hi, generate_fn(x0[lo:hi], u0[lo:hi], sc, T, lo)


def generate_openloop_sharded(generator, x0, T, rank=0, world_size=1):
    """The open-loop generator modes (OpenLoopGenerator.generate) on this rank's id block; control and noise streams are
    keyed by the global trajectory id, so the union over ranks equals the single-GPU result.  Returns (lo, hi, local_result)."""
    lo, hi = shard_range(len(x0), rank, world_size)
    return lo, hi, generator.generate(x0[lo:hi], T, traj_id0=lo)


def gather_results(local, lo, hi, num_traj, dist=None, dst=0, device=None):
    """Final gather of per-rank result dicts (arrays with leading dimension = local trajectories) onto
    rank ``dst`` in global id order.  ``dist`` = torch.distributed (initialised) or None for one process.
    Uses all_gather of padded blocks (NCCL over NVLink on GPUs, gloo on CPU)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    import torch
    W, rank = dist.get_world_size(), dist.get_rank()
    max_local = max(shard_range(num_traj, r, W)[1] - shard_range(num_traj, r, W)[0] for r in range(W))
    out = {}
    for key in sorted(local):
        a = np.ascontiguousarray(local[key])
        pad = np.zeros((max_local,) + a.shape[1:], dtype=a.dtype)
        pad[: hi - lo] = a
        t = torch.from_numpy(pad)
        if device is not None:
            t = t.to(device)
        parts = [torch.empty_like(t) for _ in range(W)]
        dist.all_gather(parts, t)
        if rank == dst:
            chunks = []
            for r in range(W):
                l, h = shard_range(num_traj, r, W)
                chunks.append(parts[r][: h - l].cpu().numpy())
            out[key] = np.concatenate(chunks, axis=0)
    return out if rank == dst else None
