"""Multi-GPU sharding of the sample batch.

Samples are independent: every rank holds the full (< 1 MB) epoch tables and
evaluates a contiguous slice of the N parameter vectors.  The only communication is
one gather of the (N,) log-L values (NCCL all_gather over NVLink on GPUs; gloo on
CPU for the host-logic tests).  One process per GPU, torch.distributed plumbing.
"""
import numpy as np


def shard_bounds(n, world_size, rank):
    """[lo, hi) of rank's contiguous slice; the first n % world_size ranks get one extra row."""
    base, extra = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n, world_size):
    return [shard_bounds(n, world_size, r)[1] - shard_bounds(n, world_size, r)[0] for r in range(world_size)]


class ShardedLikelihood(object):
    """Evaluates a global batch across the ranks of a torch.distributed process group.

    `likelihood` is a sterne_b200.Gaussianlikelihood built on this rank's GPU (or, in
    CPU tests, any object with log_likelihood_batch(theta, plan=...) -> (n,)).
    The execution plan (lanes per sample, CTAs per tile; it fixes the summation order) is
    chosen from the GLOBAL batch size, so a sample's log-L has the same bits whatever the
    number of GPUs.
    """

    def __init__(self, likelihood, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.like = likelihood
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def global_plan(self, n_global):
        ctx = getattr(self.like, '_context', None)
        if ctx is None:
            return 0
        return ctx(None).auto_plan(int(n_global))

    def local_slice(self, n_global):
        return shard_bounds(n_global, self.world, self.rank)

    def evaluate_local(self, theta_local, n_global, **kw):
        return self.like.log_likelihood_batch(theta_local, plan=self.global_plan(n_global), **kw)

    def all_gather(self, local_out, n_global):
        """(n_global,) log-L on every rank from the per-rank slices (one collective)."""
        import torch
        sizes = shard_sizes(n_global, self.world)
        is_np = isinstance(local_out, np.ndarray)
        t = torch.from_numpy(local_out) if is_np else local_out
        if len(set(sizes)) == 1:
            out = torch.empty(n_global, dtype=t.dtype, device=t.device)
            self.dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
        else:
            # collectives need equal counts: pad every slice to the largest, strip afterwards
            m = max(sizes)
            padded = torch.zeros(m, dtype=t.dtype, device=t.device)
            padded[:t.shape[0]] = t
            buf = torch.empty(m * self.world, dtype=t.dtype, device=t.device)
            self.dist.all_gather_into_tensor(buf, padded, group=self.group)
            out = torch.cat([buf[r * m:r * m + sizes[r]] for r in range(self.world)])
        return out.numpy() if is_np else out

    def log_likelihood_batch(self, theta_global, **kw):
        """Every rank passes the same global (N, P) batch (e.g. a sampler replicated on all
        ranks); each evaluates its slice and all ranks get the full (N,) result."""
        n = int(theta_global.shape[0])
        lo, hi = self.local_slice(n)
        local = self.evaluate_local(theta_global[lo:hi], n, **kw)
        return self.all_gather(local, n)
