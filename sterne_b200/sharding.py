"""Multi-GPU sharding of the sample batch, one process per GPU (torchrun).

Samples are independent: every rank holds the full (< 1 MB) epoch tables and
evaluates a contiguous slice of the N parameter vectors.  The only communication is
the gather of the (N,) log-L values, done in one of two ways:

* `PeerGather` (GPUs of one box, the product path): every rank's (N,) result buffer is
  mapped into every other rank through CUDA IPC, and the likelihood kernel's own epilogue
  stores each log L -- 8 bytes per vector -- straight into the slice's place in EVERY
  rank's buffer over NVLink (`stn_loglike_scatter`).  Compute and gather are one kernel;
  the ranks only meet at a barrier before reading.
* one NCCL `all_gather_into_tensor` (any topology; gloo on CPU for the host-logic tests).

(A single process driving several GPUs does not need this module: see
`Gaussianlikelihood(device=[...])` / `stn_multi_*`.)
"""
import ctypes

import numpy as np


def shard_bounds(n, world_size, rank):
    """[lo, hi) of rank's contiguous slice; the first n % world_size ranks get one extra row."""
    base, extra = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n, world_size):
    return [shard_bounds(n, world_size, r)[1] - shard_bounds(n, world_size, r)[0] for r in range(world_size)]


class PeerGatherUnavailable(RuntimeError):
    """The ranks cannot map each other's buffers (raised on every rank together)."""


class _DeviceArray(object):
    """Library-owned device memory seen by torch through __cuda_array_interface__."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {'shape': (int(n),), 'typestr': '<f8', 'data': (int(ptr), False),
                                         'version': 2, 'strides': None}


class PeerGather(object):
    """Two (n_max,) float64 result buffers per rank (ping-pong), mapped into every rank of `group`.

    `outs(lo)` gives the list of pointers a rank passes to `stn_loglike_scatter` so that its
    slice [lo, lo + n) lands at the same offset of every rank's current buffer; `view()` is this
    rank's current buffer as a torch tensor; `flip()` switches to the other buffer.  With one
    barrier per evaluation, alternating between two buffers is what makes it safe for a fast rank
    to start writing evaluation k+1 while a slow rank still reads evaluation k.  All ranks must be
    GPUs of one box (CUDA IPC + NVLink peer access) and at most 8 of them (STN_MAX_OUTS)."""

    def __init__(self, n_max, device, group=None):
        import torch
        import torch.distributed as dist
        from . import _lib
        self.lib = _lib.load()
        self.check = _lib.check
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = int(device)
        self.n_max = int(n_max)
        self.ptr = None
        self.peers = []
        if self.world > _lib.MAX_OUTS:
            raise PeerGatherUnavailable('PeerGather supports at most %d ranks' % _lib.MAX_OUTS)
        # Every stage is followed by a collective exchange of its outcome, so that a rank where CUDA IPC is not
        # possible (different nodes, no peer access, a sandbox without IPC) makes ALL ranks give up together
        # instead of leaving the others inside a collective.
        handle, err = None, None
        try:
            import os
            if os.environ.get('STERNE_B200_NO_IPC'):              # lets the NCCL fallback be exercised where IPC works
                raise RuntimeError('disabled by STERNE_B200_NO_IPC')
            ptr = ctypes.c_void_p()
            self.check(self.lib.stn_dev_alloc(self.device, ctypes.byref(ptr), 2 * self.n_max * 8))
            self.ptr = ptr.value
            buf = ctypes.create_string_buffer(_lib.IPC_HANDLE_BYTES)
            self.check(self.lib.stn_ipc_export(self.ptr, buf))
            handle = bytes(buf.raw)
        except Exception as e:                      # noqa: BLE001 - reported collectively below
            err = str(e)
        handles = [None] * self.world
        dist.all_gather_object(handles, handle, group=group)
        if any(h is None for h in handles):
            self._release()
            raise PeerGatherUnavailable('CUDA IPC export failed on rank(s) %s%s' % (
                [r for r, h in enumerate(handles) if h is None], ': ' + err if err else ''))
        opened = [None] * self.world
        err = None
        try:
            for r, h in enumerate(handles):
                if r == self.rank:
                    opened[r] = self.ptr
                    continue
                p = ctypes.c_void_p()
                hb = ctypes.create_string_buffer(h, _lib.IPC_HANDLE_BYTES)
                self.check(self.lib.stn_ipc_open(self.device, hb, ctypes.byref(p)))
                opened[r] = p.value
        except Exception as e:                      # noqa: BLE001
            err = str(e)
        self.peers = opened
        oks = [None] * self.world
        dist.all_gather_object(oks, err is None, group=group)
        if not all(oks):
            self._release()
            raise PeerGatherUnavailable('CUDA IPC open failed on rank(s) %s%s' % (
                [r for r, ok in enumerate(oks) if not ok], ': ' + err if err else ''))
        self.local = torch.as_tensor(_DeviceArray(self.ptr, 2 * self.n_max), device=torch.device('cuda', self.device))
        self.parity = 0
        dist.barrier(group=group)

    def _release(self):
        for r, p in enumerate(self.peers):
            if p and r != self.rank:
                self.lib.stn_ipc_close(self.device, p)
        self.peers = []
        if self.ptr:
            self.lib.stn_dev_free(self.device, self.ptr)
            self.ptr = None

    def outs(self, lo, roots=None):
        """Pointers to element `lo` of the current buffers of `roots` (default: every rank), own buffer first."""
        ranks = list(range(self.world)) if roots is None else list(roots)
        order = [self.rank] + [r for r in ranks if r != self.rank] if self.rank in ranks else ranks
        return [self.peers[r] + 8 * (self.parity * self.n_max + int(lo)) for r in order]

    def view(self, n=None):
        a = self.parity * self.n_max
        return self.local[a:a + (self.n_max if n is None else int(n))]

    def flip(self):
        self.parity ^= 1

    def close(self):
        import torch.distributed as dist
        if getattr(self, 'ptr', None):
            try:
                dist.barrier(group=self.group)            # nobody may still be writing into a buffer that goes away
            except Exception:
                pass
            self.local = None
            self._release()

    def __del__(self):
        try:
            if getattr(self, 'ptr', None):
                self.lib.stn_dev_free(self.device, self.ptr)
                self.ptr = None
        except Exception:
            pass


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
        self._peer = None
        self._flag = None
        self._fused_ok = None
        self.fused_error = None
        ctx = getattr(likelihood, '_context', None)
        if ctx is not None and self.world > 1:
            ctx(None).host_hint(self.world)          # the ranks share this host's PCIe / memory bandwidth

    def global_plan(self, n_global):
        ctx = getattr(self.like, '_context', None)
        if ctx is None:
            return 0
        return ctx(None).auto_plan(int(n_global))

    def local_slice(self, n_global):
        return shard_bounds(n_global, self.world, self.rank)

    def evaluate_local(self, theta_local, n_global, **kw):
        return self.like.log_likelihood_batch(theta_local, plan=self.global_plan(n_global), **kw)

    # ------------------------------------------------------------------ fused compute + gather
    def peer_gather(self, n_global):
        """The PeerGather of this object, (re)built when a larger batch comes along (collective).
        Raises PeerGatherUnavailable on every rank when the buffers cannot be mapped."""
        if self._peer is None or self._peer.n_max < n_global:
            if self._peer is not None:
                self._peer.close()
                self._peer = None
            self._peer = PeerGather(n_global, self.like.device, self.group)
        return self._peer

    def fused_available(self, n_global):
        """True when the kernel-stores-to-peers gather can be used for batches of n_global (collective; tried once)."""
        if self._fused_ok is None or (self._fused_ok and (self._peer is None or self._peer.n_max < n_global)):
            try:
                self.peer_gather(n_global)
                self._fused_ok = True
            except PeerGatherUnavailable as e:
                self._fused_ok = False
                self.fused_error = str(e)
        return self._fused_ok

    def evaluate_scatter(self, theta_local, n_global, keys=None, mode=None, plan=None, roots=None, barrier=True):
        """Fused path: this rank's slice (a CUDA tensor whose rows are [lo, hi) of the global batch) is
        evaluated and every log L is stored by the kernel itself into all ranks' result buffers
        (or only those of `roots`).  Returns this rank's (n_global,) buffer -- a view that stays valid
        until the evaluation after the next one; with barrier=True it is complete once the stream-ordered
        barrier (a one-element all_reduce queued on the current stream) has run."""
        import torch
        from . import _lib
        like = self.like
        peer = self.peer_gather(n_global)
        lo, hi = self.local_slice(n_global)
        if int(theta_local.shape[0]) != hi - lo:
            raise ValueError('rank %d holds rows [%d, %d) of the batch, got %d rows' % (self.rank, lo, hi, theta_local.shape[0]))
        ctx = like._context(keys)
        if theta_local.dtype != torch.float64 or theta_local.dim() != 2 or theta_local.shape[1] != ctx.problem.n_params:
            raise TypeError('theta_local must be (n, %d) float64' % ctx.problem.n_params)
        if theta_local.stride(1) != 1:
            theta_local = theta_local.contiguous()
        mode = like.mode if mode is None else _lib.MODES[mode]
        plan = self.global_plan(n_global) if plan is None else int(plan)
        stream = torch.cuda.current_stream(theta_local.device).cuda_stream
        if hi > lo:
            ctx.loglike_scatter_ptr(theta_local.data_ptr(), hi - lo, theta_local.stride(0) if hi - lo > 1 else
                                    max(theta_local.stride(0), ctx.problem.n_params), _lib.LAYOUT_AOS, mode, plan,
                                    peer.outs(lo, roots), stream)
        out = peer.view(n_global)
        peer.flip()                                   # the next evaluation writes the other buffer
        if barrier:
            self.barrier()
        return out

    def barrier(self):
        """Stream-ordered barrier: a one-element all_reduce queued behind the kernels of this stream."""
        import torch
        if self._flag is None:
            self._flag = torch.zeros(1, dtype=torch.float32, device=torch.device('cuda', self.like.device))
        self.dist.all_reduce(self._flag, group=self.group)

    # ------------------------------------------------------------------ NCCL / gloo gather
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

    def log_likelihood_batch(self, theta_global, fused=None, **kw):
        """Every rank passes the same global (N, P) batch (e.g. a sampler replicated on all
        ranks); each evaluates its slice and all ranks get the full (N,) result.  CUDA batches
        take the fused kernel-stores-to-peers path (fused=None: when the batch is a CUDA tensor)."""
        n = int(theta_global.shape[0])
        lo, hi = self.local_slice(n)
        is_cuda = hasattr(theta_global, 'is_cuda') and theta_global.is_cuda
        if fused is None:
            fused = is_cuda and self.world <= 8 and hasattr(self.like, '_context') and self.fused_available(n)
        if fused:
            import torch
            return self.evaluate_scatter(theta_global[lo:hi], n, **kw).clone()
        local = self.evaluate_local(theta_global[lo:hi], n, **kw)
        return self.all_gather(local, n)

    def close(self):
        if self._peer is not None:
            self._peer.close()
            self._peer = None
