"""Multi-GPU plumbing for closed-loop batches: the batch of independent solves shards across ranks with NO collective on the
hot path; torch.distributed (NCCL over NVLink on GPUs, gloo in CPU tests) is used only to broadcast the
shared inputs and to gather solutions / merits for best-trajectory selection (mpc_controller.py:240-242, 294-309)."""
import numpy as np


def shard_range(batch, rank, world):
    """Contiguous row block of `batch` solves owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_inputs(dist, tensors, src=0):
    """Broadcast x0 / reference window tensors from `src` to every rank (in place)."""
    for t in tensors:
        dist.broadcast(t, src=src)
    return tensors


def gather_rows(dist, local, batch, world):
    """All-gather per-rank row blocks (possibly ragged) into the full [batch, ...] tensor on every rank."""
    import torch

    sizes = [shard_range(batch, r, world)[1] - shard_range(batch, r, world)[0] for r in range(world)]
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    return torch.cat([o[:n] for o, n in zip(outs, sizes)], dim=0)


def select_best(merits):
    """Index of the best (lowest final merit, NaN-safe) solve — the selection the MPC harness makes per control step."""
    m = np.where(np.isnan(merits), np.inf, merits)
    return int(np.argmin(m))
