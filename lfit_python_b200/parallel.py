"""Walkers sharded over the GPUs of one box: one process per GPU.

The reference spreads walkers over a multiprocessing.Pool (mcmcfit.py:273-288).  Here every
rank holds the whole ensemble and the same random stream, evaluates a contiguous slice of each
half-step's proposals on its own GPU, and an all-gather (NCCL over NVLink on device tensors;
gloo on host tensors in the CPU tests) hands every rank all log-probabilities.  Walkers are
independent, so an N-GPU run returns bit-for-bit what one GPU returns.
"""
import numpy as np


def shard_bounds(n, rank, world):
    """Contiguous, balanced slice [lo, hi) of n rows for `rank` of `world`."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class ShardedLogProb:
    """log_prob_fn(theta[n, ndim]) -> (n,) evaluated cooperatively by all ranks of a
    torch.distributed group.  `local_fn(theta_rows)` evaluates rows on this rank."""

    def __init__(self, local_fn, group=None, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.local_fn, self.group = local_fn, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = device  # None: host tensors (gloo); else a cuda device (nccl)

    def __call__(self, theta):
        torch, dist = self.torch, self.dist
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        n = theta.shape[0]
        lo, hi = shard_bounds(n, self.rank, self.world)
        local = np.asarray(self.local_fn(theta[lo:hi]), dtype=np.float64) if hi > lo else np.empty(0)
        # equal-sized slots so that one all_gather_into_tensor serves ragged shards
        slot = -(-n // self.world)
        buf = torch.full((slot,), float("nan"), dtype=torch.float64)
        buf[: hi - lo] = torch.from_numpy(local)
        if self.device is not None:
            buf = buf.to(self.device)
        out = torch.empty(self.world * slot, dtype=torch.float64, device=buf.device)
        dist.all_gather_into_tensor(out, buf, group=self.group)
        out = out.cpu().numpy().reshape(self.world, slot)
        res = np.empty(n)
        for r in range(self.world):
            a, b = shard_bounds(n, r, self.world)
            res[a:b] = out[r, : b - a]
        return res
