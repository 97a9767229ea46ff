"""world_size-2 gloo test (CPU) of the multi-GPU host logic: slice bounds, the plan
choice from the GLOBAL batch, and the all_gather of per-rank log-L slices.  The
per-slice evaluator is the oracle here (no GPU in this container); on a GPU box the
same ShardedLikelihood wraps sterne_b200.Gaussianlikelihood."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sterne_b200.sharding import ShardedLikelihood, shard_bounds, shard_sizes


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 8, 9, 100003, 1 << 22):
        for w in (1, 2, 3, 4, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = shard_sizes(n, w)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1


class _OracleLike(object):
    """Stand-in evaluator with Gaussianlikelihood's batched signature."""

    def __init__(self, case):
        self.case = case
        self.seen_plans = []

    def log_likelihood_batch(self, theta, plan=0, **kw):
        from oracle import sterne_oracle as O
        self.seen_plans.append(plan)
        c = self.case
        return O.batch_log_likelihood(c.refepoch, c.LoD_timing, c.LoD_VLBI, c.shares, c.earth_xyz, np.asarray(theta))


def _worker(rank, world, port, n, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from sterne_b200 import synthetic
        case = synthetic.config4()
        theta = case.sample_theta(n, seed=3)                 # same global batch on every rank
        sharded = ShardedLikelihood(_OracleLike(case))
        got = sharded.log_likelihood_batch(theta)
        lo, hi = sharded.local_slice(n)
        ret[rank] = (got, lo, hi)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n', [64, 65])
def test_two_rank_gather_matches_single_process(n):
    from sterne_b200 import synthetic
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, n, ret), nprocs=2, join=True)
    case = synthetic.config4()
    want = _OracleLike(case).log_likelihood_batch(case.sample_theta(n, seed=3))
    assert ret[0][1:] == shard_bounds(n, 2, 0) and ret[1][1:] == shard_bounds(n, 2, 1)
    for r in (0, 1):
        assert np.array_equal(ret[r][0], want)
