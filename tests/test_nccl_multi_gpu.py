"""Hardware test of the sharded path: ranks on separate GPUs give the single-GPU bits, both through
the fused kernel-stores-to-peers gather (CUDA IPC + NVLink) and through the NCCL all_gather.
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
        sh = ShardedLikelihood(like)
        fused = sh.log_likelihood_batch(theta)                        # kernel epilogue stores into every rank's buffer
        again = sh.log_likelihood_batch(theta * 1.0)                  # second call: the other ping-pong buffer
        nccl = sh.log_likelihood_batch(theta, fused=False)            # NCCL all_gather of the slices
        lo, hi = sh.local_slice(n)
        root_only = sh.evaluate_scatter(theta[lo:hi], n, roots=[0])   # gather to rank 0 only
        torch.cuda.synchronize()
        ret[rank] = (fused.cpu().numpy(), again.cpu().numpy(), nccl.cpu().numpy(),
                     root_only.cpu().numpy() if rank == 0 else None)
        sh.close()
        like.close()
    finally:
        dist.destroy_process_group()


def test_multi_gpu_gathers_are_bit_identical_to_one_gpu():
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    from sterne_b200 import synthetic
    world = min(ngpu, 8)
    n = 50001
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, n, ret), nprocs=world, join=True)
    case = synthetic.config5('C')
    like = case.likelihood(device=0)
    theta = synthetic.device_theta(case, n, seed=5, device=torch.device('cuda', 0))
    want = like.log_likelihood_batch(theta, plan=like._context(None).auto_plan(n)).cpu().numpy()
    like.close()
    for r in range(world):
        fused, again, nccl, root_only = ret[r]
        assert np.array_equal(fused, want) and np.array_equal(again, want) and np.array_equal(nccl, want), r
    assert np.array_equal(ret[0][3], want)


def _worker_no_ipc(rank, world, port, n, ret):
    import torch.distributed as dist
    from sterne_b200 import synthetic
    from sterne_b200.sharding import ShardedLikelihood
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    os.environ['STERNE_B200_NO_IPC'] = '1'
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        case = synthetic.config5('C')
        like = case.likelihood(device=rank)
        theta = synthetic.device_theta(case, n, seed=5, device=torch.device('cuda', rank))
        sh = ShardedLikelihood(like)
        got = sh.log_likelihood_batch(theta)                  # fused unavailable on every rank -> NCCL all_gather
        ret[rank] = (got.cpu().numpy(), sh._fused_ok, sh.fused_error)
        sh.close()
        like.close()
    finally:
        dist.destroy_process_group()


def test_falls_back_to_nccl_when_buffers_cannot_be_mapped():
    """Where CUDA IPC is not possible all ranks give up on the fused gather TOGETHER and use NCCL."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    from sterne_b200 import synthetic
    n = 20001
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ret = mp.Manager().dict()
    mp.spawn(_worker_no_ipc, args=(2, port, n, ret), nprocs=2, join=True)
    case = synthetic.config5('C')
    like = case.likelihood(device=0)
    theta = synthetic.device_theta(case, n, seed=5, device=torch.device('cuda', 0))
    want = like.log_likelihood_batch(theta, plan=like._context(None).auto_plan(n)).cpu().numpy()
    like.close()
    for r in range(2):
        got, ok, err = ret[r]
        assert ok is False and 'STERNE_B200_NO_IPC' in err
        assert np.array_equal(got, want)
