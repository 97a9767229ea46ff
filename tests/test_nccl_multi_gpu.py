"""Hardware test of the sharded path: two ranks on two GPUs over NCCL give the single-GPU bits.
Skipped on boxes with fewer than two GPUs (the CPU/gloo twin is tests/test_sharding_gloo.py)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, n, ret):
    import torch.distributed as dist
    from sterne_b200 import synthetic
    from sterne_b200.sharding import ShardedLikelihood
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        case = synthetic.config5('C')
        like = case.likelihood(device=rank)
        theta = synthetic.device_theta(case, n, seed=5, device=torch.device('cuda', rank))   # same batch on each rank
        got = ShardedLikelihood(like).log_likelihood_batch(theta)
        ret[rank] = got.cpu().numpy()
        like.close()
    finally:
        dist.destroy_process_group()


def test_two_gpu_nccl_gather_is_bit_identical_to_one_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    from sterne_b200 import synthetic
    n = 50001
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, n, ret), nprocs=2, join=True)
    case = synthetic.config5('C')
    like = case.likelihood(device=0)
    theta = synthetic.device_theta(case, n, seed=5, device=torch.device('cuda', 0))
    want = like.log_likelihood_batch(theta, plan=like._context(None).auto_plan(n)).cpu().numpy()
    like.close()
    assert np.array_equal(ret[0], want) and np.array_equal(ret[1], want)
